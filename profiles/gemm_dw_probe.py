"""dW_h-shaped bf16 GEMM (900 x 900 x 16384, stream-K) in the four operand layouts: does the MN-major feed cost tensor rate?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gic_b200
from gic_b200 import _lib as L
L.require_cuda(); lib = L.lib(); dev = torch.device("cuda:0")


def bench(tA, tB, M, N, K, iters=100):
    pad = lambda x: (x + 63) // 64 * 64
    ra, ca = ((K, M) if tA else (M, K)); rb, cb = ((N, K) if tB else (K, N))
    lda, ldb = pad(ca), pad(cb)
    A = torch.randn(ra, lda, device=dev).to(torch.bfloat16); B = torch.randn(rb, ldb, device=dev).to(torch.bfloat16)
    C = torch.zeros(M, N, device=dev); s = L.stream()
    def run():
        L.check(lib.gic_gemm_bf16(tA, tB, M, N, K, 1.0, L.ptr(A), lda, L.ptr(B), ldb, 0.0, L.ptr(C), N, None, s), "g")
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"bf16 tA{tA} tB{tB} {M}x{N}x{K}: {us:7.1f} us  {2.0*M*N*K/us/1e6:6.0f} TF/s   kernels {[k for k, v in L.kernel_counts().items() if 'gemm' in k]}", flush=True)


for (M, N, K) in ((900, 900, 16384), (1024, 960, 16384), (10000, 512, 5120), (2048, 512, 5120)):
    for tA, tB in ((1, 0), (0, 0), (1, 1), (0, 1)):
        bench(tA, tB, M, N, K)
