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
cot = torch.randn(8192, 2, device="cuda")
stats = None
for rnd in range(2):
    out = plan.local_energy(params, x)
    plan.mcmc_sweep(params, xb, 1, 0.1, seed=5 + rnd)
    pot = plan.potential(xb)
    nat.slogdet(m)
    # round 2: reverse pass, KFAC factor pass + update, energy statistics
    plan.logpsi_vjp(params, xb, cot)
    raw = plan.kfac_factors(params, xb, reuse_forward=True)
    stats = raw / (8192 * 12) if stats is None else stats
    coef, ms, ml = plan.kfac_damped_factors(stats, torch.eye(4, device="cuda"), 1.0, 1e-3)
    nat.spd_inverse(ms, inplace=True)
    nat.spd_inverse(ml, inplace=True)
    plan.kfac_update(ms, ml, coef, stats, 1.0, 1e-3, torch.randn_like(params))
    big = plan.local_energy(params, xb[:1024].contiguous())
    red = nat.energy_stats(big["energy"], big)
    nat.energy_diff(big["energy"], big, red)
torch.cuda.synchronize()
print("ok", float(out["potential"].mean()), float(pot.mean()))
