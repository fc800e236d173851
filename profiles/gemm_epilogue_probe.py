"""What bounds the persistent tcgen05 GEMM on the step's large shapes: main loop or epilogue?  Each shape is timed with the
normal epilogue (TMA stores), with the global stores skipped (GIC_GEMM_DBG=1: main loop + TMEM drain only) and with plain
per-row stores (GIC_GEMM_DBG=5).   python profiles/gemm_epilogue_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gic_b200
from gic_b200 import _lib as L
L.require_cuda(); lib = L.lib(); dev = torch.device("cuda:0")


def bench(mode, tA, tB, M, N, K, iters=100, bias=False):
    A = torch.randn((K, M) if tA else (M, K), device=dev)
    B = torch.randn((N, K) if tB else (K, N), device=dev)
    C = torch.zeros(M, N, device=dev)
    bv = torch.randn(N, device=dev) if bias else None
    s = L.stream()
    out = []
    for dbg in (0, 1, 5):
        L.set_option("GIC_GEMM_DBG", dbg)
        def run():
            L.check(lib.gic_gemm(mode, tA, tB, M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], 0.0, L.ptr(C), N,
                                 L.ptr(bv) if bias else None, s), "gemm")
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): run()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / iters)
    L.clear_option("GIC_GEMM_DBG")
    gf = 2.0 * M * N * K / 1e6
    print(f"mode {mode} tA{tA} tB{tB} {M:6d}x{N:6d}x{K:6d}: TMA-store epilogue {out[0]:7.1f} us ({gf/out[0]:5.0f} TF/s, C {M*N*4/out[0]/1e6:5.2f} TB/s)"
          f" | no stores {out[1]:7.1f} us | per-row stores {out[2]:7.1f} us", flush=True)


bench(1, 0, 1, 10000, 2048, 512, bias=True)     # EW = embed W_ih^T + b   (once per step)
bench(1, 0, 1, 5120, 10000, 512)                # a projection-like wide output
bench(1, 0, 0, 5120, 512, 10000)                # dhtop
bench(1, 1, 0, 10000, 512, 5120)                # dW_out
bench(1, 0, 1, 16384, 900, 900, bias=True)      # highway, tf32
bench(1, 0, 1, 8192, 8192, 512)
