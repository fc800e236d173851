"""Where the end-to-end step (host buffers in, losses out) spends its time beyond the device-resident step:
device time of the replayed graph (which draws u and the dropout masks itself), the H2D copies, and the host-side
latency between two synchronising steps.   python profiles/e2e_breakdown.py [--mode bf16]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="bf16"); ap.add_argument("--steps", type=int, default=30)
a = ap.parse_args()
import gic_b200
from gic_b200.args import default_args
from gic_b200.training import GANInstructor
gic_b200.set_gemm_mode(bench.MODES[a.mode])
cfg = bench.WORKLOADS["c2"]; B, L, V = cfg["B"], cfg["L"], cfg["V"]
dev = torch.device("cuda:0")
args = default_args(vocab_size=V, gen_embed_dim=cfg["E"], gen_hidden_dim=cfg["H"], gen_num_layers=1,
                    disc_num_filters=list(cfg["filters"]), conditional_gan=1, feature_dim=cfg["feat"], device="cuda")
torch.manual_seed(1008)
inst = GANInstructor(args, device=dev); inst.gen.train(); inst.disc.train(); inst.gen.decoder.temperature = 1.0
g = torch.Generator(device=dev).manual_seed(1)
caps = torch.randint(4, V, (B, L), generator=g, device=dev); pooled = torch.randn(B, cfg["feat"], generator=g, device=dev)
h_caps, h_pool = caps.cpu().pin_memory(), pooled.cpu().pin_memory()

def step():
    r = inst.adv_step(h_caps, pooled=h_pool, graph=True)
    return r

for _ in range(5):
    r = step(); torch.stack([r["g_loss"], r["d_loss"]]).cpu()
torch.cuda.synchronize()
# (1) full e2e loop with the per-step D2H read
t0 = time.perf_counter()
for _ in range(a.steps):
    r = step(); torch.stack([r["g_loss"], r["d_loss"]]).cpu()
t_e2e = (time.perf_counter() - t0) / a.steps * 1e3
# (2) back-to-back replays without the per-step read (device time of graph + copies, host runs ahead)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    r = step()
e1.record(); torch.cuda.synchronize()
t_dev = e0.elapsed_time(e1) / a.steps
# (3) host time of one adv_step(graph=True) call (enqueue only)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(a.steps):
    r = step()
t_host = (time.perf_counter() - t0) / a.steps * 1e3
torch.cuda.synchronize()
# (4) the RNG draws alone
e0.record()
for _ in range(a.steps):
    u = torch.rand(L, B, V, device=dev)
e1.record(); torch.cuda.synchronize(); t_u = e0.elapsed_time(e1) / a.steps
e0.record()
for _ in range(a.steps):
    k = (torch.rand(3, B * 64, 900, device=dev) >= 0.2).to(torch.uint8)
e1.record(); torch.cuda.synchronize(); t_k = e0.elapsed_time(e1) / a.steps
print(f"e2e loop {t_e2e:.3f} ms/step | device-paced (no per-step read) {t_dev:.3f} ms | host enqueue {t_host:.3f} ms | "
      f"torch.rand u {t_u*1e3:.0f} us | dropout masks {t_k*1e3:.0f} us")
