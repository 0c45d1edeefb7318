"""Walker-ensemble estimators with the reference's estimator surface (deephall/netobs_bridge/observables).

The reference plugs these into NetObs (`netobs.observables.Estimator`): `empty_val_state(steps)`,
`evaluate(i, params, key, data, system, state, aux_data) -> (values, state)`, `digest(all_values, state)`.
NetObs is not a dependency here; the three methods keep the reference's names, argument order and return shapes,
so the NetObs evaluation loop can drive them unchanged.  The reductions run in CUDA (csrc/observables.cu); walkers
are sharded one process per GPU, and where the reference sums over its pmap axis the states / sums are all-reduced.

  PairCorrelationEstimator  netobs_bridge/observables/pair_corr.py:29-65
  DensityEstimator          netobs_bridge/observables/density.py:24-57
  OverlapEstimator          netobs_bridge/observables/overlap.py:32-72
  OneRDMEstimator           netobs_bridge/observables/one_rdm.py:28-124
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


class OneRDMEstimator:
    """One-body reduced density matrix in the lowest-Landau-level basis Y_{Q,Q,m} (one_rdm.py:67-124).

    Per walker one uniform point r' (one_rdm.py:61-64,116), N displaced copies of the walker evaluated by the network
    in one batch of B N walkers, and rdm_ij = 4 pi sum_a Psi(R_a') / Psi(R) phi_i(r_a) conj(phi_j(r')).  `key` is the
    Philox seed (or a PhiloxKey) the r' are drawn with; `evaluate` returns the per-walker matrices as the reference
    does, or with `options["reduce"]` the batch mean only (an (L, L) matrix instead of B of them)."""

    def __init__(self, network_apply, system: System, options: dict | None = None):
        self.network_apply = network_apply
        self.system = system
        self.options = options or {}
        self.flux = int(system.flux)
        self.norbs = self.flux + 1
        net = getattr(network_apply, "__self__", None)
        self._plan = net.plan(system) if net is not None else None

    def empty_val_state(self, steps: int):
        return {"one_rdm": torch.zeros((steps, self.norbs, self.norbs), dtype=torch.complex64, device="cuda")}, {}

    def uniform_sample(self, key, batch: int) -> torch.Tensor:
        # one_rdm.py:61-64: theta = arccos(U(-1, 1)), phi = U(-pi, pi) -- the distribution dh_init_walkers draws
        seed = int(getattr(key, "seed", key) or 0)
        if hasattr(key, "offset"):  # PhiloxKey: every subkey (offset block of 2^20) draws its own points
            seed = (seed + 0x9E3779B97F4A7C15 * ((int(key.offset) >> 20) + 1)) & 0x7FFFFFFFFFFFFFFF
        if self._plan is not None:
            # one subsequence per GLOBAL walker index: ranks draw different r' (as mcmc.make_mcmc_step does)
            return self._plan.init_walkers(batch, seed=seed, subsequence0=constants.rank() * batch)[:, 0, :].contiguous()
        g = torch.Generator(device="cuda").manual_seed(seed)
        u = torch.rand((batch, 2), generator=g, device="cuda")
        return torch.stack([torch.acos(2 * u[:, 0] - 1), (2 * u[:, 1] - 1) * torch.pi], -1).contiguous()

    def evaluate(self, i, params, key, data, system, state, aux_data=None, r_prime=None):
        del i, system, aux_data
        data = data.reshape(-1, *data.shape[-2:]).contiguous()
        B, N = data.shape[0], data.shape[1]
        if r_prime is None:
            r_prime = self.uniform_sample(key, B)
        data_prime = _native.one_rdm_scatter(data, r_prime)                      # one_rdm.py:92-94
        logpsi = self.network_apply(params, data)                                # :96
        logpsi_prime = self.network_apply(params, data_prime.reshape(B * N, N, 2)).reshape(B, N)  # :97
        varphi = _native.lll_orbitals(data, self.flux)                           # :98
        varphi_prime = _native.lll_orbitals(r_prime, self.flux)                  # :99
        if self.options.get("reduce"):
            total = torch.zeros((self.norbs, self.norbs), dtype=torch.complex128, device=data.device)
            _native.one_rdm_product(logpsi, logpsi_prime, varphi, varphi_prime, per_walker=False, out_sum=total)
            total = _psum(total) / (B * constants.world_size())
            return {"one_rdm": total.to(torch.complex64)}, state
        return {"one_rdm": _native.one_rdm_product(logpsi, logpsi_prime, varphi, varphi_prime)}, state

    def digest(self, all_values, state):
        del state
        one_rdm = all_values["one_rdm"].mean(0)  # one_rdm.py:121-124
        return {"diagonal": torch.diagonal(one_rdm), "trace": torch.trace(one_rdm)}
