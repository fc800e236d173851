"""bf16-operand GEMM (tcgen05 kind::f16) vs the TF32 kernel on the discriminator's shapes, back-to-back launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gic_b200
from gic_b200 import _lib as L
L.require_cuda(); lib = L.lib(); dev = torch.device("cuda:0")

def bench(bf, tA, tB, M, N, K, beta=0.0, iters=100, ld=None):
    ra, ca = ((K, M) if tA else (M, K)); rb, cb = ((N, K) if tB else (K, N))
    lda = ld or ca; ldb = ld or cb
    dt = torch.bfloat16 if bf else torch.float32
    A = torch.randn(ra, lda, device=dev).to(dt); B = torch.randn(rb, ldb, device=dev).to(dt)
    C = torch.zeros(M, N, device=dev); s = L.stream()
    def run():
        if bf: L.check(lib.gic_gemm_bf16(tA, tB, M, N, K, 1.0, L.ptr(A), lda, L.ptr(B), ldb, beta, L.ptr(C), N, None, s), "g")
        else: L.check(lib.gic_gemm(1, tA, tB, M, N, K, 1.0, L.ptr(A), lda, L.ptr(B), ldb, beta, L.ptr(C), N, None, s), "g")
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"{'bf16' if bf else 'tf32'} tA{tA} tB{tB} {M}x{N}x{K} ld{lda} beta{beta}: {us:7.1f} us  {2.0*M*N*K/us/1e6:6.0f} TF/s", flush=True)

for bf in (0, 1):
    ld = 960 if bf else None
    bench(bf, 0, 1, 16384, 900, 900, ld=ld)              # highway
    bench(bf, 0, 0, 16384, 900, 900, beta=1.0, ld=ld)    # dx
    bench(bf, 1, 0, 900, 900, 16384, ld=ld)              # dW_h
    bench(bf, 0, 1, 8192, 8192, 8192, iters=10)
    bench(bf, 0, 1, 18944, 1024, 1024)
