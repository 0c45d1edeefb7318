"""The reference's NetObs adaptor on the B200 engine (deephall/netobs_bridge/adaptor.py:35-121).

`DeepHallAdaptor` keeps the method names, arguments and return structure NetObs drives: `restore(ckpt_file)`,
`call_signed_network`, `call_network`, `make_walking_step`, `call_local_kinetic_energy`,
`call_local_potential_energy`.  NetObs, OmegaConf and universal_pathlib are not dependencies here: `config.yml` next
to the checkpoint is read with PyYAML, and `evaluate_observable` is the few lines of NetObs's evaluation loop the
estimators of `deephall_b200.observables` need (walk, evaluate, collect the step values, digest).
One process per GPU: every rank restores its walker shard; parameters are replicated.
"""
from __future__ import annotations

import os
from typing import Any

import torch
import yaml

from . import constants, mcmc
from .checkpoint import restore_checkpoint
from .config import Config
from .hamiltonian import make_local_kinetic_energy, make_potential
from .networks import make_network


class HallSystem(dict):
    """netobs_bridge/hall_system.py:18-19: an `ElectronGas` mapping with `flux`."""


class DeepHallAdaptor:
    def __init__(self, config: Any = None, args: list[str] | None = None) -> None:
        self.config, self.args = config, args or []

    def restore(self, ckpt_file: str | None = None):
        # adaptor.py:40-65
        if ckpt_file is None:
            raise ValueError("Must specify a checkpoint")
        config_path = os.path.join(os.path.dirname(os.fspath(ckpt_file)), "config.yml")
        with open(config_path) as f:
            self.cfg = cfg = Config.from_dict(yaml.safe_load(f))
        self.model = model = make_network(cfg.system, cfg.network)
        self.network = model.apply
        self.batch_per_device = cfg.batch_size // constants.world_size()
        self.kinetic_energy = make_local_kinetic_energy(self.network, cfg.system)
        self.potential_energy = make_potential(cfg.system, model)
        _, state = restore_checkpoint(ckpt_file, model)
        data = state.data
        if data.shape[0] != self.batch_per_device:  # a whole-batch file: this rank takes its shard (train.py:60)
            r = constants.rank()
            data = data[r * self.batch_per_device:(r + 1) * self.batch_per_device].contiguous()
        return (state.params, data, HallSystem(spins=list(cfg.system.nspins), ndim=2, flux=cfg.system.flux),
                {"mcmc_width": state.mcmc_width})

    def call_network(self, params, electrons, system=None):
        del system
        return self.network(params, electrons)

    def call_signed_network(self, params, electrons, system=None):
        # adaptor.py:67-71
        return torch.ones((), device=electrons.device), self.call_network(params, electrons, system)

    def make_walking_step(self, batch_log_psi=None, steps: int = 10, system=None):
        # adaptor.py:73-92 (the sampler is the fused sweep of this network; batch_log_psi is not re-traced)
        del batch_log_psi, system
        mcmc_step = mcmc.make_mcmc_step(self.network, self.batch_per_device, steps)

        def walk(key, params, electrons, aux_data):
            new_data, _ = mcmc_step(params, electrons, key, aux_data["mcmc_width"])
            return new_data, aux_data

        return walk

    def call_local_kinetic_energy(self, params, key, electrons, system=None):
        del key, system
        return self.kinetic_energy(params, electrons)[0]  # adaptor.py:94-102

    def call_local_potential_energy(self, params, key, electrons, system=None):
        del params, key, system
        return self.potential_energy(electrons) * self.cfg.system.interaction_strength  # adaptor.py:104-112


DEFAULT = DeepHallAdaptor


def evaluate_observable(adaptor: DeepHallAdaptor, estimator, ckpt_file: str, steps: int, mcmc_steps: int = 10, seed: int = 0):
    """restore -> `steps` x (walk, evaluate) -> digest.  Step values with a walker axis are averaged over the walkers of
    every rank, as NetObs does before it stores them.  Returns (digest, all_values, state)."""
    params, data, system, aux = adaptor.restore(ckpt_file)
    walk = adaptor.make_walking_step(None, mcmc_steps, system)
    all_values, state = estimator.empty_val_state(steps)
    key = mcmc.PhiloxKey(seed)
    for i in range(steps):
        key, k_walk = key.split()
        key, k_eval = key.split()
        data, aux = walk(k_walk, params, data, aux)
        values, state = estimator.evaluate(i, params, k_eval, data, system, state, aux)
        for name, v in values.items():
            if v.dim() > all_values[name].dim() - 1:  # per-walker values: mean over the (global) batch
                # NaN-aware mean over all walkers of all ranks (the reference's digest uses nanmean): masked sum and
                # count, reduced together, so one NaN ratio drops that walker, not the step
                bad = torch.isnan(torch.view_as_real(v)).any(-1) if v.is_complex() else torch.isnan(v)
                cnt = constants.pmean((~bad).sum(0).to(torch.float32))
                tot = constants.pmean(torch.where(bad, torch.zeros_like(v), v).sum(0))
                v = tot / cnt.clamp(min=1e-30)
            all_values[name][i] = v
    if hasattr(estimator, "gathered_state"):
        state = estimator.gathered_state(state)
    return estimator.digest(all_values, state), all_values, state
