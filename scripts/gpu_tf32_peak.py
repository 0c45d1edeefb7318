"""Measured dense TF32 throughput of this board (torch.matmul with TF32 allowed, 8192^3, best of 10 and sustained over 3 s):
the peak the TF32-piece contraction plans (dh_config.contraction = 1, the range-guard fallback) stand against."""
import time
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.time(); k = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(10):
        a @ b
    k += 10
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / k
fl = 2.0 * n ** 3
print(f"tf32 dense: burst {fl / best / 1e9:.1f} TFLOP/s, sustained {fl / sus / 1e9:.1f} TFLOP/s")
