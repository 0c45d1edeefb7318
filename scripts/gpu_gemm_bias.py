"""Developer tool: error statistics (max, rms, signed mean = bias) of the tcgen05 GEMM vs fp64."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
torch.manual_seed(0)
M, K, N = 4096, 256, 256
for kind in ("gauss", "positive"):
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(K, N, device="cuda") / 16
    if kind == "positive":
        A, W = A.abs(), W.abs()
    ref = A.double() @ W.double()
    for impl in (0, 1, 2):
        out = nat.gemm(A, W, None, 1, impl=impl).double()
        rel = (out - ref) / ref.abs().clamp(min=ref.abs().mean())
        print(f"{kind} impl{impl}: max {rel.abs().max():.2e} rms {rel.pow(2).mean().sqrt():.2e} signed mean {rel.mean():+.2e} (sign-weighted {(rel * ref.sign()).mean():+.2e})")
