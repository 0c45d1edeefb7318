"""Quick GPU check after a GEMM change: dh_gemm (tcgen05) vs fp64 on odd shapes, then c3 timings."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from deephall_b200 import _native as nat

dev = "cuda"
torch.manual_seed(0)
bad = 0
for (M, N, K, rpg) in [(128, 256, 256, 1), (300, 256, 256, 1), (1000, 768, 256, 4), (77, 48, 256, 3), (129, 816, 256, 32),
                       (64, 18, 32, 1), (5000, 96, 64, 1), (148 * 128 * 3 + 17, 256, 256, 32), (40000, 816, 256, 32)]:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(K, N, device=dev) / 16
    b = torch.randn(N, device=dev)
    ref = A.double() @ W.double()
    ref[::rpg] += b.double()
    out = nat.gemm(A, W, b, rpg, impl=1)
    torch.cuda.synchronize()
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    bad += err > 2e-6
    print(f"tc M{M} N{N} K{K} rpg{rpg}: max rel err {err:.2e}", flush=True)
print("GEMM", "FAIL" if bad else "ok", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "time":
    import gpu_check
    gpu_check.timing(dict(nspins=(12, 0), flux=33), 8192)
