"""Walker-ensemble estimators with the reference's estimator surface (deephall/netobs_bridge/observables).

The reference plugs these into NetObs (`netobs.observables.Estimator`): `empty_val_state(steps)`,
`evaluate(i, params, key, data, system, state, aux_data) -> (values, state)`, `digest(all_values, state)`.
NetObs is not a dependency here; the three methods keep the reference's names, argument order and return shapes,
so the NetObs evaluation loop can drive them unchanged.  The reductions run in CUDA (csrc/observables.cu); walkers
are sharded one process per GPU, and where the reference sums over its pmap axis the states / sums are all-reduced.

  PairCorrelationEstimator  netobs_bridge/observables/pair_corr.py:29-65
  DensityEstimator          netobs_bridge/observables/density.py:24-57
  OverlapEstimator          netobs_bridge/observables/overlap.py:32-72
"""
from __future__ import annotations

import dataclasses

import torch
import torch.distributed as dist

from . import _native, constants
from .config import Network, System
from .networks import make_network


def _psum(t: torch.Tensor) -> torch.Tensor:
    if constants.world_size() > 1:
        if t.is_complex():
            r = torch.view_as_real(t.clone().contiguous())
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
            return torch.view_as_complex(r)
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class PairCorrelationEstimator:
    """g(theta_12) histogram (pair_corr.py:42-60).  `state["pair_corr"]` accumulates, un-normalised by the number
    of evaluation steps as in the reference (pair_corr.py:57); with sharded walkers every rank holds its share of
    the sum (normalised by the global batch) and `gathered_state` all-reduces it."""

    def __init__(self, options: dict | None = None):
        self.options = options or {}
        self.bins = int(self.options.get("bins", 200))

    def empty_val_state(self, steps: int):
        del steps
        return {}, {"pair_corr": torch.zeros(self.bins, dtype=torch.float32, device="cuda")}

    def evaluate(self, i, params, key, data, system, state, aux_data=None):
        del i, params, key, system, aux_data
        data = data.reshape(-1, *data.shape[-2:]).contiguous()
        _native.pair_correlation(data, state["pair_corr"], self.bins, batch_norm=data.shape[0] * constants.world_size())
        return {}, state

    def digest(self, all_values, state):
        del all_values, state
        return {}

    @staticmethod
    def gathered_state(state):
        return {"pair_corr": _psum(state["pair_corr"])}


class DensityEstimator:
    """Polar-angle histogram of the electrons (density.py:42-49); integer counts."""

    def __init__(self, options: dict | None = None):
        self.options = options or {}
        self.hist_bins = int(self.options.get("bins", 50))

    def empty_val_state(self, steps: int):
        del steps
        return {}, {"map": torch.zeros(self.hist_bins, dtype=torch.int64, device="cuda")}

    def evaluate(self, i, params, key, data, system, state, aux_data=None):
        del i, params, key, system, aux_data
        data = data.reshape(-1, *data.shape[-2:]).contiguous()
        _native.density_histogram(data, state["map"])
        return {}, state

    def digest(self, all_values, state):
        del all_values, state
        return {}

    @staticmethod
    def gathered_state(state):
        return {"map": _psum(state["map"])}


class OverlapEstimator:
    """Overlap of the network's wavefunction with the analytic Laughlin state of the same system (overlap.py:32-72).

    `ratio` carries one global phase that depends on the branch of Im(log) both implementations happen to return
    (the reference's `shift` is a mean of complex logs, overlap.py:60); |ratio|^2 and the digested overlap do not."""

    def __init__(self, network_apply, system: System, network: Network | None = None):
        self.network_apply = network_apply
        self.system = system
        laughlin = make_network(system, dataclasses.replace(network or Network(), type="laughlin"))  # overlap.py:44-46
        self.laughlin = laughlin
        self._no_params = torch.zeros(0, device="cuda")

    def empty_val_state(self, steps: int):
        return {"ratio": torch.zeros(steps, dtype=torch.complex64, device="cuda"),
                "ratio_square": torch.zeros(steps, dtype=torch.float32, device="cuda")}, {}

    def evaluate(self, i, params, key, data, system, state, aux_data=None):
        del i, key, system, aux_data
        data = data.reshape(-1, *data.shape[-2:]).contiguous()
        logpsi = self.network_apply(params, data)
        logphi = self.laughlin.apply(self._no_params, data)
        total = _psum(_native.overlap_sum(logphi, logpsi))                 # overlap.py:60 (mean over every device)
        shift = total / (data.shape[0] * constants.world_size())
        ratio, ratio_square = _native.overlap_ratio(logphi, logpsi, shift)  # overlap.py:61-62
        return {"ratio": ratio, "ratio_square": ratio_square}, state

    def digest(self, all_values, state):
        del state
        ratio, ratio_square = all_values["ratio"], all_values["ratio_square"]
        ok = ~(torch.isnan(ratio.real) | torch.isnan(ratio.imag))
        okq = ~torch.isnan(ratio_square)
        # nanmean over every step and walker of every rank (overlap.py:68-70)
        num = _psum(torch.where(ok, ratio, torch.zeros_like(ratio)).to(torch.complex128).sum())
        n1 = _psum(ok.sum().to(torch.float64))
        den = _psum(torch.where(okq, ratio_square, torch.zeros_like(ratio_square)).to(torch.float64).sum())
        n2 = _psum(okq.sum().to(torch.float64))
        return {"overlap": ((num / n1).abs() ** 2 / (den / n2)).to(torch.float32)}
