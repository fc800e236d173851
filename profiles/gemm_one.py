"""Run one GEMM shape a few times (for ncu).  usage: gemm_one.py tA tB M N K [beta]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gic_b200
from gic_b200 import _lib as L
L.require_cuda(); lib = L.lib(); dev = torch.device("cuda:0")
tA, tB, M, N, K = [int(x) for x in sys.argv[1:6]]
beta = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
A = torch.randn((K, M) if tA else (M, K), device=dev)
B = torch.randn((N, K) if tB else (K, N), device=dev)
C = torch.zeros(M, N, device=dev)
for _ in range(3):
    L.check(lib.gic_gemm(1, tA, tB, M, N, K, 1.0, L.ptr(A), A.shape[1], L.ptr(B), B.shape[1], beta, L.ptr(C), N, None, L.stream()), "gemm")
torch.cuda.synchronize()
print("ok")
