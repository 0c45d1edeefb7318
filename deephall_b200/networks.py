"""Wavefunction factory with the reference's constructor contract.

Mirrors deephall/networks/__init__.py:22-37 (`make_network(system, network)`) and the
`model.init(key, x)` / `model.apply(params, x)` surface of the flax module
(deephall/train.py:62,84; types.py:68-70).  The arithmetic runs in libdeephall_b200.so.

Parameters are ONE flat fp32 CUDA tensor whose layout follows the flax tree
(`Psiformer.param_tree` gives the nested `{'params': {...}}` view, `Psiformer.from_tree`
packs a reference checkpoint tree).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import _native
from .config import Network, System

_PLANS: dict = {}


def get_plan(nspins, flux, ndets, num_heads, heads_dim, num_layers, interaction_type="coulomb",
             interaction_strength=1.0, radius=None, chunk_walkers=0) -> _native.Plan:
    key = (tuple(nspins), int(flux), ndets, num_heads, heads_dim, num_layers, str(interaction_type),
           float(interaction_strength), radius, chunk_walkers, torch.cuda.current_device())
    if key not in _PLANS:
        _PLANS[key] = _native.Plan(nspins, flux, ndets, num_heads, heads_dim, num_layers, interaction_type,
                                   interaction_strength, radius, chunk_walkers)
    return _PLANS[key]


class Psiformer:
    """B200 Psiformer (networks/psiformer.py:63-91).  Constructor arguments are the ones
    `make_network` passes in the reference."""

    def __init__(self, nspins, Q, ndets=1, num_heads=4, heads_dim=64, num_layers=2, orbital_type="full"):
        if str(orbital_type) != "full":
            raise NotImplementedError("orbital_type='sparse' is a 'next' row (SURVEY 8f N4)")
        self.nspins = (int(nspins[0]), int(nspins[1]))
        self.Q = float(Q)
        self.flux = int(round(2 * self.Q))
        self.ndets, self.num_heads, self.heads_dim, self.num_layers = int(ndets), int(num_heads), int(heads_dim), int(num_layers)
        self.orbital_type = str(orbital_type)

    # ---- plan access (system-dependent parts default to the reference defaults)
    def plan(self, system: System | None = None) -> _native.Plan:
        if system is None:
            return get_plan(self.nspins, self.flux, self.ndets, self.num_heads, self.heads_dim, self.num_layers)
        return get_plan(self.nspins, self.flux, self.ndets, self.num_heads, self.heads_dim, self.num_layers,
                        system.interaction_type, system.interaction_strength, system.radius)

    @property
    def nelec(self) -> int:
        return sum(self.nspins)

    def param_layout(self) -> "OrderedDict[str, tuple[int, tuple[int, ...]]]":
        return self.plan().param_layout()

    # ---- model.init
    def init(self, key, x=None, device="cuda") -> torch.Tensor:
        """flax init distributions: lecun-normal (truncated at 2 sigma) kernels with
        fan_in = contracted input size, zero biases, unit LayerNorm scales, ee_par = 1.
        `key` is an int seed (or anything with a `.seed` attribute)."""
        seed = int(getattr(key, "seed", key))
        gen = torch.Generator().manual_seed(seed)
        lay = self.param_layout()
        flat = torch.zeros(self.plan().num_params, dtype=torch.float32)
        D = self.num_heads * self.heads_dim
        for name, (off, shape) in lay.items():
            n = int(np.prod(shape))
            if name.endswith("/kernel"):
                fan_in = 4 if name.endswith("Dense_0/kernel") else D
                std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
                t = torch.empty(n, dtype=torch.float64)
                torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=gen)
                flat[off : off + n] = (t * std).float()
            elif name.endswith("/scale") or name.startswith("Jastrow_0/"):
                flat[off : off + n] = 1.0
        return flat.to(device)

    def param_tree(self, flat: torch.Tensor) -> dict:
        """Nested {'params': {...}} dict of views into `flat`, keyed like the flax tree."""
        root: dict = {}
        for name, (off, shape) in self.param_layout().items():
            node = root
            parts = name.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = flat[off : off + int(np.prod(shape))].view(shape)
        return {"params": root}

    def from_tree(self, tree: dict, device="cuda") -> torch.Tensor:
        """Pack a (reference-checkpoint style) nested tree of arrays into the flat layout."""
        root = tree.get("params", tree)
        flat = torch.empty(self.plan().num_params, dtype=torch.float32)
        for name, (off, shape) in self.param_layout().items():
            node = root
            for p in name.split("/"):
                node = node[p]
            if isinstance(node, torch.Tensor):
                arr = node.detach().to("cpu", torch.float32)
            else:
                arr = torch.as_tensor(np.asarray(node), dtype=torch.float32)
            if tuple(arr.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {tuple(arr.shape)}")
            flat[off : off + arr.numel()] = arr.reshape(-1)
        return flat.to(device)

    # ---- model.apply
    def apply(self, params: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """complex64 log psi.  x: (N, 2) -> scalar, or (B, N, 2) -> (B,)."""
        single = x.dim() == 2
        xb = (x[None] if single else x).contiguous().float()
        out = self.plan().logpsi(params, xb)
        return out[0] if single else out

    __call__ = apply


def make_network(system: System, network: Network) -> Psiformer:
    """networks/__init__.py:22-37."""
    if str(network.type) == "laughlin":
        raise NotImplementedError("the analytic Laughlin network is a 'next' row (SURVEY 8f N3); "
                                  "oracle/laughlin.py holds its CPU restatement for tests")
    ps = network.psiformer
    return Psiformer(
        Q=system.flux / 2,
        nspins=system.nspins,
        ndets=ps.determinants,
        num_heads=ps.num_heads,
        num_layers=ps.num_layers,
        heads_dim=ps.heads_dim,
        orbital_type=network.orbital,
    )
