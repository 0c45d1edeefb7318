"""Developer tool: role wait-time breakdown of the tcgen05 GEMM at the c3 local-energy shape (needs DH_GEMM_PROF=1)."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
M, K = 1024 * 12 * 32, 256
A = torch.randn(M, K, device="cuda")
for N in [int(a) for a in sys.argv[1:]] or [256]:
    W = torch.randn(K, N, device="cuda") / 16
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    for _ in range(2):
        nat.gemm(A, W, b, 32, out=out, impl=1)
    torch.cuda.synchronize()
