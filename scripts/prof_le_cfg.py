"""Profiling target: local-energy passes of a BASELINE configuration (c2 | c4 | c5k4) on B walkers."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

CFG = {"c2": dict(nspins=(6, 0), flux=15), "c3": dict(nspins=(12, 0), flux=33), "c4": dict(nspins=(10, 0), flux=21),
       "c5k4": dict(nspins=(16, 0), flux=45, ndets=4)}
name = sys.argv[1] if len(sys.argv) > 1 else "c5k4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
plan = nat.Plan(**CFG[name])
torch.manual_seed(0)
params = torch.randn(plan.num_params, device="cuda") * 0.05
x = plan.init_walkers(B, seed=1)
for _ in range(2):
    out = plan.local_energy(params, x)
torch.cuda.synchronize()
print("ok", float(out["potential"].mean()))
