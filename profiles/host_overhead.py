"""Host-side enqueue time vs device time of Decoder.sample and the full step (is the GPU starved by launches?)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import gic_b200  # noqa: E402
from gic_b200.args import default_args  # noqa: E402
from gic_b200.training import GANInstructor  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
gic_b200.set_gemm_mode(bench.MODES[mode])
cfg = bench.WORKLOADS["c2"]
B, L, V = cfg["B"], cfg["L"], cfg["V"]
dev = torch.device("cuda:0")
args = default_args(vocab_size=V, gen_embed_dim=cfg["E"], gen_hidden_dim=cfg["H"], disc_num_filters=list(cfg["filters"]),
                    conditional_gan=1, feature_dim=cfg["feat"], device="cuda")
torch.manual_seed(1008)
inst = GANInstructor(args, device=dev)
inst.gen.train(); inst.disc.train(); inst.gen.decoder.temperature = 1.0
g = torch.Generator(device=dev).manual_seed(1)
caps = torch.randint(4, V, (B, L), generator=g, device=dev)
pooled = torch.randn(B, cfg["feat"], generator=g, device=dev)
u = torch.rand(L, B, V, generator=g, device=dev)
keep = (torch.rand(3, B * 64, 900, generator=g, device=dev) >= 0.2).to(torch.uint8)


def decode():
    with torch.no_grad():
        f = inst.gen.encoder(pooled)
        inst.gen.decoder.sample(f, max_caption_len=L, u=u)


def step():
    inst.adv_step(caps, pooled=pooled, u=u, keep=keep)


for name, fn in (("decode", decode), ("step", step)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/iter, total {1e3 * (t2 - t0) / n:.3f} ms/iter", flush=True)
