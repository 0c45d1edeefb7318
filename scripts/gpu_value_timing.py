"""Developer diagnostic: value-only path at c3, B = 8192: log psi pass, 10-move Metropolis sweep, per-category times."""
import sys

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402

plan = nat.Plan(nspins=(12, 0), flux=33)
flat = torch.randn(plan.num_params, device="cuda") * 0.05
x = plan.init_walkers(8192, seed=1)
plan.mcmc_sweep(flat, x, 20, 0.1, seed=0)


def t(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"logpsi        {t(lambda: plan.logpsi(flat, x)):7.3f} ms")
print(f"mcmc 10 moves {t(lambda: plan.mcmc_sweep(flat, x, 10, 0.1, seed=1), 5):7.3f} ms")
plan.profile_begin()
plan.logpsi(flat, x)
prof = plan.profile_end()
print({k: (round(v["ms"], 3), v["count"]) for k, v in prof.items() if v["count"]})
cot = torch.randn(8192, 2, device="cuda") / 8192
print(f"logpsi_vjp    {t(lambda: plan.logpsi_vjp(flat, x, cot), 5):7.3f} ms")
plan.profile_begin()
plan.logpsi_vjp(flat, x, cot)
prof = plan.profile_end()
print({k: (round(v["ms"], 3), v["count"], round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)) for k, v in prof.items() if v["count"]})
print(f"kfac_factors  {t(lambda: plan.kfac_factors(flat, x), 5):7.3f} ms")
