"""tcgen05 3xTF32 GEMM: correctness vs fp64 and throughput vs the SIMT kernel."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

dev = "cuda"
torch.manual_seed(0)
for (M, N, K, rpg) in [(128, 256, 256, 1), (300, 256, 256, 1), (1000, 768, 256, 4), (77, 48, 256, 3), (129, 816, 256, 32), (64, 18, 32, 1), (5000, 96, 64, 1)]:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(K, N, device=dev) / 16
    b = torch.randn(N, device=dev)
    ref = A.double() @ W.double()
    ref[::rpg] += b.double()
    for impl in (0, 1):
        out = nat.gemm(A, W, b, rpg, impl=impl)
        torch.cuda.synchronize()
        err = (out.double() - ref).abs().max() / ref.abs().max()
        print(f"impl{impl} M{M} N{N} K{K} rpg{rpg}: max rel err {err:.2e}", flush=True)
# throughput at the c3 local-energy shape
M, K = 1024 * 12 * 32, 256
A = torch.randn(M, K, device=dev)
for N in (256, 768, 816):
    W = torch.randn(K, N, device=dev) / 16
    out = torch.empty(M, N, device=dev)
    for impl in (0, 1):
        nat.gemm(A, W, None, 1, out=out, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        # note: impl=1 through dh_gemm includes the weight split + a stream sync per call
        e0.record()
        for _ in range(5):
            nat.gemm(A, W, None, 1, out=out, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"impl{impl} M{M} N{N}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
