import sys, torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
plan = nat.Plan(nspins=(12, 0), flux=33)
x = plan.init_walkers(8192, seed=1)
for _ in range(3): phi = nat.lll_orbitals(x, 33)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): phi = nat.lll_orbitals(x, 33)
e1.record(); torch.cuda.synchronize()
print(f"lll_orbitals 98304 points x 34: {e0.elapsed_time(e1)/20*1000:.1f} us")
