"""Back-to-back launches of one GEMM shape, CUDA-event timed: separates the per-launch fixed cost of the
tcgen05 kernel from its main loop.   python profiles/gemm_microbench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gic_b200  # noqa: E402
from gic_b200 import _lib as L  # noqa: E402

L.require_cuda()
lib = L.lib()
dev = torch.device("cuda:0")


def bench(mode, tA, tB, M, N, K, iters=200, beta=0.0):
    A = torch.randn((K, M) if tA else (M, K), device=dev)
    B = torch.randn((N, K) if tB else (K, N), device=dev)
    C = torch.zeros(M, N, device=dev)
    s = L.stream()

    def run():
        L.check(lib.gic_gemm(mode, tA, tB, M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], beta, L.ptr(C), N,
                             None, s), "gemm")
    for _ in range(10):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
    print(f"mode {mode} tA{tA} tB{tB} {M:6d}x{N:6d}x{K:6d} beta {beta}: {us:8.2f} us/launch  {tf:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for K in (32, 128, 512, 2048, 8192):
        bench(1, 0, 1, 128, 128, K)
    for K in (32, 512, 2048):
        bench(1, 0, 1, 256, 2048, K)
    bench(1, 0, 1, 256, 2048, 512, beta=1.0)
    bench(1, 0, 1, 256, 10000, 512)
    bench(1, 0, 1, 16384, 900, 900)
    bench(1, 0, 0, 16384, 900, 900)
    bench(1, 1, 0, 900, 900, 16384)
    bench(1, 0, 1, 8192, 8192, 8192, iters=10)
    bench(1, 0, 1, 5120, 64, 10000)
    bench(0, 0, 1, 256, 2048, 512)
    bench(0, 0, 1, 16384, 900, 900, iters=20)
