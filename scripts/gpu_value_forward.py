"""Developer diagnostic: one value-only forward (log psi) at c3 -- run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import sys

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
plan = nat.Plan(nspins=(12, 0), flux=33)
torch.manual_seed(0)
params = torch.randn(plan.num_params, device="cuda") * 0.05
x = plan.init_walkers(B, seed=1)
for _ in range(3):
    out = plan.logpsi(params, x)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(20):
    out = plan.logpsi(params, x)
ev[1].record()
torch.cuda.synchronize()
print(f"B={B}: log psi {ev[0].elapsed_time(ev[1]) / 20:.4f} ms")
