"""One eager c2a step (c2 with additive attention over the [256, 49, 2048] grid) between cudaProfilerStart/Stop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, gic_b200
from gic_b200.args import default_args
from gic_b200.training import GANInstructor
gic_b200.set_gemm_mode(gic_b200.GEMM_BF16)
B, L, V = 256, 20, 10000
a = default_args(vocab_size=V, gen_embed_dim=512, gen_hidden_dim=512, gen_num_layers=1, conditional_gan=1, feature_dim=2048, device="cuda", gen_attention=1)
torch.manual_seed(1008)
inst = GANInstructor(a, device="cuda:0"); inst.gen.train(); inst.disc.train(); inst.gen.decoder.temperature = 1.0
g = torch.Generator(device="cuda:0").manual_seed(1)
caps = torch.randint(4, V, (B, L), generator=g, device="cuda:0")
grid = torch.randn(B, 49, 2048, generator=g, device="cuda:0"); pooled = grid.mean(1)
u = torch.rand(L, B, V, generator=g, device="cuda:0"); keep = (torch.rand(3, B * 64, 900, generator=g, device="cuda:0") >= 0.2).to(torch.uint8)
for _ in range(3):
    inst.adv_step(caps, pooled=pooled, u=u, keep=keep, grid=grid)
torch.cuda.synchronize()
torch.cuda.profiler.start()
inst.adv_step(caps, pooled=pooled, u=u, keep=keep, grid=grid)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one c2a step")
