"""Profiling target for the per-kernel ncu table: one local-energy chunk, one Metropolis move, the potential and the
public slogdet at the c3 configuration (everything run twice; profile the second round)."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
from deephall_b200 import networks

plan = nat.Plan(nspins=(12, 0), flux=33)
params = networks.Psiformer((12, 0), 16.5).init(0)
x = plan.init_walkers(1024, seed=1)
xb = plan.init_walkers(8192, seed=2)
m = torch.randn(8192, 1, 12, 12, dtype=torch.complex64, device="cuda")
for rnd in range(2):
    out = plan.local_energy(params, x)
    plan.mcmc_sweep(params, xb, 1, 0.1, seed=5 + rnd)
    pot = plan.potential(xb)
    nat.slogdet(m)
torch.cuda.synchronize()
print("ok", float(out["potential"].mean()), float(pot.mean()))
