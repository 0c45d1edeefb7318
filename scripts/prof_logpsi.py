"""Profiling target: value-only log psi passes (c3, 8192 walkers), as used by the Metropolis sweep."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
plan = nat.Plan(nspins=(12, 0), flux=33)
torch.manual_seed(0)
params = (torch.randn(plan.num_params, device="cuda") * 0.05)
x = plan.init_walkers(B, seed=1)
for _ in range(2):
    out = plan.logpsi(params, x)
torch.cuda.synchronize()
print("ok", float(out.real.mean()))
