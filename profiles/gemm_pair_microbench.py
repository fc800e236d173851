"""CTA-pair (cta_group::2) bf16 GEMM vs one CTA per tile on the discriminator's shapes (GIC_GEMM_2CTA=1 / 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gic_b200
from gic_b200 import _lib as L
L.require_cuda(); lib = L.lib(); dev = torch.device("cuda:0")


def bench(tB, M, N, K, iters=100, ld=None, bias=False):
    rb, cb = ((N, K) if tB else (K, N))
    lda = ld or K; ldb = ld or cb
    A = torch.randn(M, lda, device=dev).to(torch.bfloat16); B = torch.randn(rb, ldb, device=dev).to(torch.bfloat16)
    C = torch.zeros(M, N, device=dev); s = L.stream()
    bv = torch.randn(N, device=dev) if bias else None
    out = []
    for flag in ("0", "1"):
        L.set_option("GIC_GEMM_2CTA", int(flag))
        def run():
            L.check(lib.gic_gemm_bf16(0, tB, M, N, K, 1.0, L.ptr(A), lda, L.ptr(B), ldb, 0.0, L.ptr(C), N, L.ptr(bv) if bias else None, s), "g")
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        out.append(us)
    L.clear_option("GIC_GEMM_2CTA")
    f = 2.0 * M * N * K / 1e6
    print(f"tB{tB} {M}x{N}x{K} ld{lda}{' +bias' if bias else ''}: one CTA per tile {out[0]:7.1f} us {f/out[0]:6.0f} TF/s | CTA pair {out[1]:7.1f} us {f/out[1]:6.0f} TF/s", flush=True)


bench(1, 16384, 900, 900, ld=960)        # highway
bench(1, 16384, 900, 900, ld=960, bias=True)
bench(0, 16384, 900, 900, ld=960)        # dx
bench(1, 32768, 900, 900, ld=960)
bench(1, 262144, 900, 900, ld=960, iters=10)   # c5
bench(1, 8192, 8192, 8192, iters=10)
bench(1, 18944, 1024, 1024)
