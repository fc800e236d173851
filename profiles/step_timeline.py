"""Two-stream timeline of the graph-replayed c2 step: timing events recorded at the phase boundaries of
GANInstructor.adv_step (captured into the graph as event-record nodes), averaged over replays.
    python profiles/step_timeline.py [--mode bf16] [--e2e]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="bf16"); ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--e2e", action="store_true", help="no uniforms / masks supplied: drawn inside the step")
a = ap.parse_args()
import gic_b200
from gic_b200.args import default_args
from gic_b200.training import GANInstructor
gic_b200.set_gemm_mode(bench.MODES[a.mode])
cfg = bench.WORKLOADS["c2"]; B, L, V = cfg["B"], cfg["L"], cfg["V"]
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:                 # under torchrun: the data-parallel step, one timeline per rank
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
args = default_args(vocab_size=V, gen_embed_dim=cfg["E"], gen_hidden_dim=cfg["H"], gen_num_layers=1,
                    disc_num_filters=list(cfg["filters"]), conditional_gan=1, feature_dim=cfg["feat"], device="cuda")
torch.manual_seed(1008)
inst = GANInstructor(args, device=dev); inst.gen.train(); inst.disc.train(); inst.gen.decoder.temperature = 1.0
g = torch.Generator(device=dev).manual_seed(1 + rank)
caps = torch.randint(4, V, (B, L), generator=g, device=dev); pooled = torch.randn(B, cfg["feat"], generator=g, device=dev)
u = None if a.e2e else torch.rand(L, B, V, generator=g, device=dev)
keep = None if a.e2e else (torch.rand(3, B * 64, 900, generator=g, device=dev) >= 0.2).to(torch.uint8)
inst.timeline = []
inst.adv_step(caps, pooled=pooled, u=u, keep=keep, graph="static")          # eager step + capture (markers captured too)
marks = inst.timeline[len(inst.timeline) // 2:]                              # the second half belongs to the captured pass
inst.timeline = None
acc = {}
for _ in range(a.reps):
    inst.adv_step(caps, pooled=pooled, u=u, keep=keep, graph="static")
    torch.cuda.synchronize()
    t0 = marks[0][1]
    for name, ev in marks:
        acc.setdefault(name, []).append(t0.elapsed_time(ev) * 1e3)
lines = [f"rank {rank}/{world}: c2 step timeline, mode {a.mode}, {'library-drawn' if a.e2e else 'supplied'} randomness (us after the first marker, mean of {a.reps} replays)"]
for name, ev in marks:
    v = acc[name]; lines.append(f"  {sum(v) / len(v):8.1f}  {name}")
if world > 1:
    import time
    dist.barrier(); time.sleep(0.2 * rank)
print("\n".join(lines), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
