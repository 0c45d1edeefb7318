"""Minimal VMC driver: the call sequence of deephall/train.py:80-167 on the B200 engine.

This is NOT a re-implementation of the reference's training driver (logging, checkpoints and
KFAC are out of scope); it exists so that `VMC steps/sec` is measured on the
reference's own sequence: mcmc_step -> update_mcmc_width -> training_step.
One process per GPU (torchrun); walkers are sharded, parameters replicated.
"""
from __future__ import annotations

import numpy as np
import torch

from . import constants, mcmc
from .config import Config
from .networks import make_network
from .optimizers import CheckpointState, make_optimizer_step


class VMC:
    def __init__(self, cfg: Config, device="cuda"):
        self.cfg = cfg
        self.world = constants.world_size()
        self.rank = constants.rank()
        assert cfg.batch_size % self.world == 0
        self.batch_per_device = cfg.batch_size // self.world
        self.model = make_network(cfg.system, cfg.network)
        self.network = self.model.apply
        self.mcmc_step = mcmc.make_mcmc_step(self.network, self.batch_per_device, cfg.mcmc.steps)
        self.opt_init, self.training_step = make_optimizer_step(cfg, self.network)
        self.key = mcmc.PhiloxKey(cfg.seed)
        # train.py:57-65: same parameters on every replica, distinct walkers per replica
        params = self.model.init(cfg.seed + 1, device=device)
        data = mcmc.init_guess(cfg.seed, self.batch_per_device, self.model.nelec, self.model,
                               subsequence0=self.rank * self.batch_per_device)
        self.state = CheckpointState(params, data, self.opt_init(params, None, data), cfg.mcmc.width)
        self.pmoves = np.zeros(cfg.mcmc.adapt_frequency)
        self.t = 0

    def burn_in(self, n=None):
        for _ in range(self.cfg.mcmc.burn_in if n is None else n):  # train.py:108-110
            self.key, sub = self.key.split()
            self.mcmc_step(self.state.params, self.state.data, sub, self.state.mcmc_width)

    def _range_overflow(self) -> bool:
        """True if any rank's fp16-piece contraction saturated an operand since the last check (dh_plan_status)."""
        bits = 0
        for plan in {id(p): p for p in (self.model.plan(), self.model.plan(self.cfg.system))}.values():
            bits |= plan.status()
        flag = torch.tensor([float(bits & 1)], device=self.state.params.device)
        return bool(constants.pmean(flag).item() > 0)

    def step(self, sync_stats=True):
        """One iteration of train.py:126-140.  Returns (pmove, stats).

        With sync_stats (the reference's loop synchronises here as well, mcmc.py:180) the fp16 range guard is checked: if
        a contraction saturated an fp16 operand piece during the iteration, the iteration is redone from the saved state
        with TF32 pieces (fp32 exponent range, the reference's range) and the network stays in that mode."""
        if sync_stats and getattr(self.model, "contraction", None) == "f16":
            saved = (self.state._replace(data=self.state.data.clone()), self.key, self.t, self.pmoves.copy())
            out = self._step(sync_stats)
            if self._range_overflow():
                self.model.contraction = "tf32"
                self.state, self.key, self.t, self.pmoves = saved
                out = self._step(sync_stats)
            return out
        return self._step(sync_stats)

    def _step(self, sync_stats=True):
        st = self.state
        self.key, sub = self.key.split()
        data, pmove = self.mcmc_step(st.params, st.data, sub, st.mcmc_width)
        width = st.mcmc_width
        if sync_stats:  # host-side width adaptation needs pmove on the host (train.py:131, mcmc.py:180)
            width, self.pmoves = mcmc.update_mcmc_width(self.t, width, self.cfg.mcmc.adapt_frequency, pmove, self.pmoves)
        self.state = st._replace(data=data, mcmc_width=width)
        self.key, sub = self.key.split()
        self.state, stats = self.training_step(self.state, sub)
        self.t += 1
        return pmove, stats


def train(cfg: Config, log=print):
    vmc = VMC(cfg)
    vmc.burn_in()
    for it in range(cfg.optim.iterations):
        pmove, stats = vmc.step()
        if log and constants.rank() == 0:
            e = stats["energy"]
            log(f"step={it} pmove={float(pmove):.2f} energy={float(e.real):.4f} energy_imag={float(e.imag):+.4f} "
                f"variance={float(stats['variance']):.4f} L_square={float(stats['angular_momentum_square']):.4f}")
        if torch.isnan(stats["energy"].real).any():  # train.py:159,166
            raise SystemExit("=" * 30 + " ABORT " + "=" * 30)
    return vmc
