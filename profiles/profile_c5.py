"""One eager c5 step (discriminator-only, 4096 real + 4096 fake captions of length 32) between cudaProfilerStart/Stop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, gic_b200
from gic_b200.args import default_args
from gic_b200.training import GANInstructor
gic_b200.set_gemm_mode(gic_b200.GEMM_BF16)
B, L, V = 4096, 32, 10000
a = default_args(vocab_size=V, gen_embed_dim=512, gen_hidden_dim=512, gen_num_layers=1, conditional_gan=0, device="cuda")
torch.manual_seed(1008)
inst = GANInstructor(a, device="cuda:0"); inst.gen.train(); inst.disc.train()
g = torch.Generator(device="cuda:0").manual_seed(1)
caps = torch.randint(4, V, (B, L), generator=g, device="cuda:0"); fake = torch.randint(4, V, (B, L), generator=g, device="cuda:0")
keep = (torch.rand(2, B * 64, 900, generator=g, device="cuda:0") >= 0.2).to(torch.uint8)
for _ in range(2):
    inst.disc_step(caps, fake, keep=keep)
torch.cuda.synchronize()
torch.cuda.profiler.start()
inst.disc_step(caps, fake, keep=keep)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one c5 step")
