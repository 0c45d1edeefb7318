"""Developer diagnostic: duration of the tcgen05 contraction at small row counts (run under ncu --metrics gpu__time_duration.sum)."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
K = 256
for M in (256, 2048, 12288, 49152):
    A = torch.randn(M, K, device="cuda")
    for N in (256, 768):
        W = torch.randn(K, N, device="cuda") / 16
        b = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        for _ in range(3):
            nat.gemm(A, W, b, 1, out=out, impl=1)
torch.cuda.synchronize()
print("ok")
