"""Developer diagnostic: time of each part of a KFAC iteration (c3), per walkers-per-GPU count."""
import sys
import time

import torch

sys.path.insert(0, ".")
from deephall_b200 import kfac as K  # noqa: E402
from deephall_b200 import loss, mcmc, networks  # noqa: E402
from deephall_b200.config import Network, Optim, System  # noqa: E402


def timed(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


system = System(flux=33, nspins=(12, 0))
model = networks.make_network(system, Network())
params = model.init(0)
plan = model.plan(system)
for B in (int(a) for a in (sys.argv[1:] or ["8192", "1024"])):
    data = mcmc.init_guess(1, B, 12, model)
    mcmc.make_mcmc_step(model.apply, B, steps=10)(params, data, mcmc.PhiloxKey(2), 0.1)
    lg = loss.make_loss_fn(model.apply, system)
    init, step = K.make_kfac_training_step(Optim().kfac, lg, model.apply, system)
    t_lg, (stats, grads) = timed(lambda: lg(params, data))
    t_fac, raw = timed(lambda: plan.kfac_factors(params, data))
    from deephall_b200.optimizers import CheckpointState
    st = CheckpointState(params, data, init(params, None, data), 0.1)
    t_step, _ = timed(lambda: step(st, None))
    t_sweep, _ = timed(lambda: plan.mcmc_sweep(params, data, 10, 0.1, seed=3))
    print(f"B={B}: loss_and_grad {t_lg:.2f} ms | kfac_factors {t_fac:.2f} ms | whole kfac step {t_step:.2f} ms | "
          f"=> stats-normalise + precondition + update {t_step - t_lg - t_fac:.2f} ms | sweep {t_sweep:.2f} ms")
    for n, b in ((257, 13), (256, 14), (408, 2), (410, 29)):
        m = torch.randn(b, n, n, device="cuda")
        m = m @ m.transpose(1, 2) / n + 0.1 * torch.eye(n, device="cuda")
        t_inv, _ = timed(lambda: K._native.spd_inverse(m))
        print(f"   dh_spd_inverse {b} x {n}^2: {t_inv:.2f} ms")
