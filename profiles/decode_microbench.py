"""Decoder.sample at the c2 shape (B 256, L 20, V 10 000, E = H = 512), TF32 mode, replayed from a CUDA graph:
the fused decode step (one kernel per step) against the round-1 path (GIC_DECODE_STEP=0: LSTM-step kernel + projection /
sample kernel), and -- with GIC_VS_STAMPS=1 -- the per-CTA timeline of one fused step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, gic_b200
from gic_b200 import _lib
import gic_b200.generator as G
from gic_b200.args import default_args

B, L, V, E, H = [int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (256, 20, 10000, 512, 512))]
a = default_args(vocab_size=V, gen_embed_dim=E, gen_hidden_dim=H, gen_num_layers=1, conditional_gan=0, device="cuda")
torch.manual_seed(0)
gen = G.Generator(a).to("cuda:0"); gen.train(); gen.decoder.temperature = 1.0
us = [torch.rand(L, B, V, device="cuda:0") for _ in range(2)]
feats = torch.randn(B, E, device="cuda:0") * 0.05
gic_b200.set_gemm_mode(gic_b200.GEMM_TF32)


def timed(tag, env):
    for k, v in env.items():                 # switches of the current library context (gic_ctx_set_option)
        _lib.set_option(k, int(v))
    try:
        graphs = []
        with torch.no_grad():
            for u in us:
                gen.decoder.sample(feats, max_caption_len=L, u=u); torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    gen.decoder.sample(feats, max_caption_len=L, u=u)
                graphs.append(g)
        for i in range(5):
            graphs[i % 2].replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 40
        e0.record()
        for i in range(n):
            graphs[i % 2].replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{tag:28s} {ms * 1e3:8.1f} us per decode  ({ms * 1e3 / L:6.2f} us per step)   kernels: "
              f"{ {k: v for k, v in _lib.kernel_counts().items() if 'step' in k or 'vocab' in k} }", flush=True)
    finally:
        for k in env:
            _lib.clear_option(k)


if os.environ.get("GIC_VS_STAMPS") == "1":
    with torch.no_grad():
        for _ in range(3):
            gen.decoder.sample(feats, max_caption_len=L, u=us[0])
    torch.cuda.synchronize()
    _lib.lib().gic_vs_stamps_table(148)
else:
    timed("fused decode step", {})
    timed("LSTM step + vocab/sample", {"GIC_DECODE_STEP": "0"})
    timed("fused decode step (again)", {})
