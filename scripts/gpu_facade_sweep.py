"""Developer diagnostic: Metropolis step through the facade (mcmc.make_mcmc_step) vs the plan call, c3, 8192 walkers."""
import sys

import torch

sys.path.insert(0, ".")
from deephall_b200 import mcmc, networks  # noqa: E402
from deephall_b200.config import Network, System  # noqa: E402

system = System(flux=33, nspins=(12, 0))
model = networks.make_network(system, Network())
params = model.init(0)
B = 8192
data = mcmc.init_guess(1, B, 12, model)
step = mcmc.make_mcmc_step(model.apply, B, steps=10)
plan = model.plan()


def timed(fn, n, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("facade, 1 warm-up + 3:", timed(lambda: step(params, data, mcmc.PhiloxKey(7), 0.1), 3, 1))
print("facade, 3 warm-up + 10:", timed(lambda: step(params, data, mcmc.PhiloxKey(7), 0.1), 10, 3))
print("plan,   3 warm-up + 10:", timed(lambda: plan.mcmc_sweep(params, data, 10, 0.1, seed=7), 10, 3))
l0 = plan.launch_count
step(params, data, mcmc.PhiloxKey(7), 0.1)
print("launches per sweep:", plan.launch_count - l0)
