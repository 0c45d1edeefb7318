"""Developer diagnostic: Metropolis sweep time per walkers-per-GPU count (c3), determinism of the graph-replayed sweep."""
import sys
import time

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402

plan = nat.Plan(nspins=(12, 0), flux=33)
torch.manual_seed(0)
params = torch.randn(plan.num_params, device="cuda") * 0.05
for B in (8192, 2048, 1024):
    x = plan.init_walkers(B, seed=1)
    for _ in range(3):
        plan.mcmc_sweep(params, x, 10, 0.1, seed=3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for i in range(n):
        nacc, _ = plan.mcmc_sweep(params, x, 10, 0.1, seed=3, offset=100 + 10 * i)
    torch.cuda.synchronize()
    print(f"B={B}: sweep of 10 moves {(time.perf_counter() - t0) / n * 1e3:.3f} ms, acceptance {int(nacc) / (10 * B):.3f}")
print("ok")
