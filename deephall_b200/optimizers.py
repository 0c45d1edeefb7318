"""Optimizer step factories with the reference's signature (deephall/optimizers/__init__.py:25-35).

`make_optimizer_step(cfg, network) -> (init, step)`; `step(state, key) -> (state, stats)`.
adam (optimizers/adam.py:24-43, optax.adam defaults b1=0.9, b2=0.999, eps=1e-8, no weight
decay), none (optimizers/none.py:22-35) and kfac (optimizers/kfac.py:198-241, the reference default; kfac.py here,
curvature statistics from dh_kfac_factors) are provided.

Unlike the reference's Adam path, which applies the device-local gradient un-reduced
(SURVEY 2.1), the gradient is all-reduced (mean) over ranks so replicas cannot drift.
"""
from __future__ import annotations

from typing import NamedTuple

import torch

from . import constants
from .config import Config
from .loss import make_loss_fn


class CheckpointState(NamedTuple):  # types.py:42-46
    params: torch.Tensor
    data: torch.Tensor
    opt_state: object
    mcmc_width: float


class AdamState(NamedTuple):
    count: int
    mu: torch.Tensor
    nu: torch.Tensor


def make_adam_training_step(optim_cfg, loss_grad_fn, b1=0.9, b2=0.999, eps=1e-8):
    def init(params, key, data):
        del key, data
        return AdamState(0, torch.zeros_like(params), torch.zeros_like(params))

    def step(state: CheckpointState, key):
        del key
        params, data, opt, width = state
        stats, grads = loss_grad_fn(params, data)
        grads = constants.pmean(grads)
        count = opt.count + 1
        mu = opt.mu.mul(b1).add_(grads, alpha=1 - b1)
        nu = opt.nu.mul(b2).addcmul_(grads, grads, value=1 - b2)
        lr = optim_cfg.lr.schedule(opt.count)  # optax passes the pre-increment count to the schedule
        mhat = mu / (1 - b1**count)
        vhat = nu / (1 - b2**count)
        params = params - lr * mhat / (vhat.sqrt() + eps)
        return CheckpointState(params, data, AdamState(count, mu, nu), width), stats

    return init, step


def make_inference_step(loss_grad_fn):
    def init(params, key, data):
        del params, key, data
        return None

    def step(state: CheckpointState, key):
        del key
        stats, _ = loss_grad_fn(state.params, state.data)
        return state, stats

    return init, step


def make_optimizer_step(cfg: Config, network):
    name = str(cfg.optim.optimizer) if cfg.optim.optimizer is not None else "none"
    if name == "adam":
        return make_adam_training_step(cfg.optim.adam, make_loss_fn(network, cfg.system))
    if name == "none":
        from .loss import LossMode

        return make_inference_step(make_loss_fn(network, cfg.system, LossMode.ENERGY_DIFF))
    if name == "kfac":
        from .kfac import make_kfac_training_step

        return make_kfac_training_step(cfg.optim.kfac, make_loss_fn(network, cfg.system), network, cfg.system)
    raise ValueError(f"Optimizer {cfg.optim.optimizer} is not implemented!")
