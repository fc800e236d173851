"""Long-run check of the replayed c2 step: N training steps (the generator collapses onto a few tokens within ~300 steps on
random data, which is the regime that exposed the data-dependent decode tail), variants chosen on the command line:
  sets=1|2      one resident input set or two alternating (two captured graphs)
  e2e=0|1       interleave steps that draw their noise in-kernel (u = None, keep = None; graph=True)
  decg=0|1      also capture / replay stand-alone Decoder.sample graphs first
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gic_b200
from gic_b200.args import default_args
from gic_b200.training import GANInstructor
opt = dict(x.split("=") for x in sys.argv[1:])
nsets, e2e, decg, steps = int(opt.get("sets", 2)), int(opt.get("e2e", 0)), int(opt.get("decg", 0)), int(opt.get("steps", 450))
gic_b200.set_gemm_mode(3)
B, L, V = 256, 20, 10000
a = default_args(vocab_size=V, gen_embed_dim=512, gen_hidden_dim=512, gen_num_layers=1, conditional_gan=1, feature_dim=2048, device="cuda")
torch.manual_seed(1008)
inst = GANInstructor(a, device="cuda:0"); inst.gen.train(); inst.disc.train(); inst.gen.decoder.temperature = 1.0
g = torch.Generator(device="cuda:0").manual_seed(1)
sets = []
for i in range(nsets):
    caps = torch.randint(4, V, (B, L), generator=g, device="cuda:0"); caps[:, 0] = 1; caps[:, -1] = 2
    sets.append(dict(caps=caps, pooled=torch.randn(B, 2048, generator=g, device="cuda:0"), u=torch.rand(L, B, V, generator=g, device="cuda:0"),
                     keep=(torch.rand(3, B * 64, 900, generator=g, device="cuda:0") >= 0.2).to(torch.uint8)))
if decg:
    gs = []
    with torch.no_grad():
        for s in sets:
            f = inst.gen.encoder(s["pooled"]); inst.gen.decoder.sample(f, max_caption_len=L, u=s["u"]); torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                f = inst.gen.encoder(s["pooled"]); inst.gen.decoder.sample(f, max_caption_len=L, u=s["u"])
            gs.append(gr)
    for i in range(20):
        gs[i % nsets].replay()
    torch.cuda.synchronize(); print("decode graphs ok", flush=True)
h_caps = [s["caps"].cpu().pin_memory() for s in sets]; h_pool = [s["pooled"].cpu().pin_memory() for s in sets]
for i in range(steps):
    s = sets[i % nsets]
    if e2e and i % 3 == 2:
        r = inst.adv_step(h_caps[i % nsets], pooled=h_pool[i % nsets], graph=True)
    else:
        r = inst.adv_step(s["caps"], pooled=s["pooled"], u=s["u"], keep=s["keep"], graph="static")
    if i % 50 == 0:
        torch.cuda.synchronize()
        ids = r["ids"]
        print(i, "losses %.4f %.4f" % (float(r["g_loss"]), float(r["d_loss"])), "max rows with the same token at a step",
              int(max(torch.bincount(ids[:, t]).max() for t in range(L))), flush=True)
torch.cuda.synchronize(); print("done", opt)
