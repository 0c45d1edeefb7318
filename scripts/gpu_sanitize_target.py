"""Small target for compute-sanitizer: every round-2 kernel once (c3-shaped network, few walkers)."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat, kfac as K, loss, mcmc, networks
from deephall_b200.config import Network, Optim, System
from deephall_b200.optimizers import CheckpointState

system = System(flux=33, nspins=(12, 0))
model = networks.make_network(system, Network())
params = model.init(0)
B = 8
data = mcmc.init_guess(1, B, 12, model)
mcmc.make_mcmc_step(model.apply, B, steps=3)(params, data, mcmc.PhiloxKey(2), 0.1)
lg = loss.make_loss_fn(model.apply, system)
stats, grads = lg(params, data)
init, step = K.make_kfac_training_step(Optim().kfac, lg, model.apply, system)
st = CheckpointState(params, data, init(params, None, data), 0.1)
st, stats = step(st, None)
torch.cuda.synchronize()
print("ok", complex(stats["energy"]), float(grads.norm()), model.plan(system).status())
