"""Two eager Decoder.sample calls at the c2 shape in TF32 mode (no CUDA graph): the target of ncu captures of the fused
decode-step kernel (`-k regex:vocab_sample_kernel -s 25 -c 2`: skip the first decode, capture two mid steps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, gic_b200
import gic_b200.generator as G
from gic_b200.args import default_args
B, L, V, E, H = 256, 20, 10000, 512, 512
a = default_args(vocab_size=V, gen_embed_dim=E, gen_hidden_dim=H, gen_num_layers=1, conditional_gan=0, device="cuda")
torch.manual_seed(0)
gen = G.Generator(a).to("cuda:0"); gen.train(); gen.decoder.temperature = 1.0
u = torch.rand(L, B, V, device="cuda:0")
feats = torch.randn(B, E, device="cuda:0") * 0.05
gic_b200.set_gemm_mode(gic_b200.GEMM_TF32)
with torch.no_grad():
    for _ in range(2):
        p, ids = gen.decoder.sample(feats, max_caption_len=L, u=u)
torch.cuda.synchronize()
print("ok", float(p.sum()), int(ids.sum()))
