"""Per-CTA timeline (globaltimer stamps, GIC_VS_STAMPS=1) of the fused vocab projection + Gumbel-softmax + sample kernel at the c2 shape."""
import os, sys
os.environ["GIC_VS_STAMPS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, gic_b200
from gic_b200 import _lib
import gic_b200.generator as G
from gic_b200.args import default_args
a = default_args(vocab_size=10000, gen_embed_dim=512, gen_hidden_dim=512, gen_num_layers=1, conditional_gan=0, device="cuda")
torch.manual_seed(0)
gen = G.Generator(a).to("cuda:0"); gen.train(); gen.decoder.temperature = 1.0
u = torch.rand(20, 256, 10000, device="cuda:0"); feats = torch.randn(256, 512, device="cuda:0") * 0.05
gic_b200.set_gemm_mode(gic_b200.GEMM_TF32)
with torch.no_grad():
    for _ in range(3):
        gen.decoder.sample(feats, max_caption_len=20, u=u)
torch.cuda.synchronize()
L = _lib.lib()
L.gic_vs_stamps_table(126)
