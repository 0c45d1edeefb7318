"""Profiling target: three local-energy passes (c3, one chunk of 1024 walkers)."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

plan = nat.Plan(nspins=(12, 0), flux=33)
torch.manual_seed(0)
params = (torch.randn(plan.num_params, device="cuda") * 0.05)
x = plan.init_walkers(1024, seed=1)
for _ in range(6):
    out = plan.local_energy(params, x)
torch.cuda.synchronize()
print("ok", float(out["potential"].mean()))
