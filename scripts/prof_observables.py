"""Profiling target: the estimator kernels at the c3 walker batch (8192 x 12), after a warm-up call each."""
import sys

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat

B, N, flux = 8192, 12, 33
plan = nat.Plan(nspins=(N, 0), flux=flux)
x = plan.init_walkers(B, seed=1)
rp = plan.init_walkers(B, seed=2)[:, 0, :].contiguous()
state = torch.zeros(200, device="cuda")
counts = torch.zeros(50, dtype=torch.int64, device="cuda")
lp = torch.randn(B, dtype=torch.complex64, device="cuda")
lpp = torch.randn(B, N, dtype=torch.complex64, device="cuda")
for _ in range(2):
    nat.pair_correlation(x, state)
    nat.density_histogram(x, counts)
    s = nat.overlap_sum(lp, lp * 0.5)
    nat.overlap_ratio(lp, lp * 0.5, s / B)
    xp = nat.one_rdm_scatter(x, rp)
    phi = nat.lll_orbitals(x, flux)
    phip = nat.lll_orbitals(rp, flux)
    nat.one_rdm_product(lp, lpp, phi, phip)
torch.cuda.synchronize()
print("ok", float(state.sum()), int(counts.sum()))
