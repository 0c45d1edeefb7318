"""Developer diagnostic: local-energy pass time (c3, 8192 walkers) with the per-category breakdown."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

plan = nat.Plan(nspins=(12, 0), flux=33)
torch.manual_seed(0)
params = torch.randn(plan.num_params, device="cuda") * 0.05
x = plan.init_walkers(8192, seed=1)
for _ in range(3):
    out = plan.local_energy(params, x)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(10):
    out = plan.local_energy(params, x)
e[1].record()
torch.cuda.synchronize()
plan.profile_begin()
plan.local_energy(params, x)
prof = plan.profile_end()
print(f"local energy 8192 walkers: {e[0].elapsed_time(e[1]) / 10:.3f} ms", prof if not isinstance(prof, dict) else {k: round(v, 3) if isinstance(v, float) else v for k, v in prof.items()})
