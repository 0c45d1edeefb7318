"""Metropolis-Hastings sampling with the reference's factory signature (deephall/mcmc.py).

`make_mcmc_step(batch_network, batch_per_device, steps)` returns
`mcmc_step(params, data, key, width) -> (data, pmove)` as mcmc.py:105-150 does.  The walker
buffer is updated in place (the reference donates it, train.py:75).  `key` is a `PhiloxKey`
(the reference's jax threefry keys cannot be generated here); the in-kernel generator is
Philox4x32-10 keyed by (seed, offset) with one subsequence per walker.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import constants
from .networks import B200Network as Psiformer  # any network of this engine


@dataclass(frozen=True)
class PhiloxKey:
    """Counter-based RNG state: (seed, offset).  `split()` mirrors jax.random.split usage in
    train.py:109,127: returns (next_key, subkey); a subkey owns 2^20 consecutive offsets."""

    seed: int
    offset: int = 0

    def split(self):
        return PhiloxKey(self.seed, self.offset + (1 << 20)), PhiloxKey(self.seed, self.offset)


def init_guess(key, batch: int, nelec: int, network: Psiformer, subsequence0: int = 0):
    """train.py:40-54: uniform points on the sphere, (batch, nelec, 2)."""
    seed = int(getattr(key, "seed", key))
    return network.plan().init_walkers(batch, seed=seed, subsequence0=subsequence0)


def make_mcmc_step(batch_network, batch_per_device: int, steps: int = 10):
    net = getattr(batch_network, "__self__", batch_network)
    if not isinstance(net, Psiformer):
        raise TypeError("batch_network must be `model.apply` of a deephall_b200 network")

    def mcmc_step(params: torch.Tensor, data: torch.Tensor, key: PhiloxKey, width):
        assert data.shape[0] == batch_per_device
        w = float(width)
        nacc, _ = net.plan().mcmc_sweep(params, data, steps, w, seed=key.seed, offset=key.offset,
                                  subsequence0=constants.rank() * batch_per_device)
        pmove = nacc.to(torch.float32) / (steps * batch_per_device)  # mcmc.py:146
        pmove = constants.pmean(pmove)  # mcmc.py:147
        return data, pmove

    return mcmc_step


def update_mcmc_width(t, width, adapt_frequency, pmove, pmoves, pmove_max=0.55, pmove_min=0.5):
    """mcmc.py:153-186 (host side)."""
    t_since = t % adapt_frequency
    pmoves[t_since] = float(pmove.reshape(-1)[0].item()) if hasattr(pmove, "reshape") else float(pmove)
    if t > 0 and t_since == 0:
        if np.mean(pmoves) > pmove_max:
            width *= 1.1
        elif np.mean(pmoves) < pmove_min:
            width /= 1.1
    return width, pmoves
