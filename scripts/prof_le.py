"""Profiling target: one local-energy pass (c3, one chunk of 1024 walkers) after a warm-up pass."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
plan = nat.Plan(nspins=(12, 0), flux=33)
torch.manual_seed(0)
params = (torch.randn(plan.num_params, device="cuda") * 0.05)
x = plan.init_walkers(B, seed=1)
for _ in range(2):
    out = plan.local_energy(params, x)
torch.cuda.synchronize()
print("ok", float(out["potential"].mean()))
