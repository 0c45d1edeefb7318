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
             interaction_strength=1.0, radius=None, chunk_walkers=0, network_type="psiformer", cf_flux=1,
             orbital_type="full", excitation_lz=0.0, contraction="f16") -> _native.Plan:
    key = (tuple(nspins), int(flux), ndets, num_heads, heads_dim, num_layers, str(interaction_type),
           float(interaction_strength), radius, chunk_walkers, str(network_type), int(cf_flux), str(orbital_type),
           float(excitation_lz), str(contraction), torch.cuda.current_device())
    if key not in _PLANS:
        _PLANS[key] = _native.Plan(nspins, flux, ndets, num_heads, heads_dim, num_layers, interaction_type,
                                   interaction_strength, radius, chunk_walkers, network_type, cf_flux, orbital_type,
                                   excitation_lz, contraction)
    return _PLANS[key]


class B200Network:
    """Common surface of the networks this engine evaluates: `plan(system)`, `init`, `apply`, `nelec`."""

    nspins: tuple

    @property
    def nelec(self) -> int:
        return sum(self.nspins)

    def apply(self, params: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """complex64 log psi.  x: (N, 2) -> scalar, or (B, N, 2) -> (B,)."""
        single = x.dim() == 2
        xb = (x[None] if single else x).contiguous().float()
        out = self.plan().logpsi(params, xb)
        return out[0] if single else out

    __call__ = apply


class Laughlin(B200Network):
    """Analytic Laughlin ground state, quasihole and quasiparticle (networks/laughlin.py:19-100), evaluated by the same tail kernels
    (log-determinant jets, local-energy assembly, Metropolis sweep).  It has no parameters."""

    def __init__(self, nspins, flux, cf_flux=1, excitation_lz=0):
        self.nspins = (int(nspins[0]), int(nspins[1]))
        self.flux = int(flux)
        self.cf_flux = int(cf_flux)
        self.excitation_lz = float(excitation_lz)
        n = sum(self.nspins)
        two_q1 = self.flux - 2 * self.cf_flux * (n - 1)
        if two_q1 not in (n - 2, n - 1, n) or two_q1 < 0:
            raise ValueError("Filling not supported")  # laughlin.py:47
        if self.nspins[1] != 0:
            raise NotImplementedError("the analytic Laughlin states are built for spin-polarised systems")
        if two_q1 != n - 1:  # quasihole / quasiparticle: laughlin.py:38-45,49-52
            d = self.excitation_lz - two_q1 / 2
            reach = abs(two_q1 / 2) + (1 if two_q1 == n - 2 else 0)
            if abs(d - round(d)) > 1e-9 or abs(self.excitation_lz) > reach:
                raise AssertionError(f"Impossible Lz={self.excitation_lz} for excitation")

    def plan(self, system: System | None = None) -> _native.Plan:
        kw = dict(network_type="laughlin", cf_flux=self.cf_flux, excitation_lz=self.excitation_lz)
        if system is None:
            return get_plan(self.nspins, self.flux, 1, 4, 64, 0, **kw)
        return get_plan(self.nspins, self.flux, 1, 4, 64, 0, system.interaction_type, system.interaction_strength,
                        system.radius, **kw)

    def init(self, key=None, x=None, device="cuda") -> torch.Tensor:
        return torch.zeros(0, dtype=torch.float32, device=device)

    def param_layout(self):
        return OrderedDict()


class Psiformer(B200Network):
    """B200 Psiformer (networks/psiformer.py:63-91).  Constructor arguments are the ones
    `make_network` passes in the reference."""

    def __init__(self, nspins, Q, ndets=1, num_heads=4, heads_dim=64, num_layers=2, orbital_type="full"):
        orbital_type = str(getattr(orbital_type, "value", orbital_type))  # OrbitalType enum or its string
        if orbital_type not in ("full", "sparse"):
            raise ValueError(f"orbital_type must be 'full' or 'sparse' (config.py:87-89), got {orbital_type!r}")
        self.nspins = (int(nspins[0]), int(nspins[1]))
        self.Q = float(Q)
        self.flux = int(round(2 * self.Q))
        self.ndets, self.num_heads, self.heads_dim, self.num_layers = int(ndets), int(num_heads), int(heads_dim), int(num_layers)
        self.orbital_type = str(orbital_type)
        # arithmetic of the contractions: "f16" = fp16 hi/lo pieces (default); "tf32" = TF32 pieces, the fp32-exponent-range
        # fallback a caller switches to when `plan().status()` reports a saturated fp16 piece; "fp32" = plain FMA
        self.contraction = "f16"

    # ---- plan access (system-dependent parts default to the reference defaults)
    def plan(self, system: System | None = None) -> _native.Plan:
        if system is None:
            return get_plan(self.nspins, self.flux, self.ndets, self.num_heads, self.heads_dim, self.num_layers,
                            orbital_type=self.orbital_type, contraction=self.contraction)
        return get_plan(self.nspins, self.flux, self.ndets, self.num_heads, self.heads_dim, self.num_layers,
                        system.interaction_type, system.interaction_strength, system.radius, orbital_type=self.orbital_type,
                        contraction=self.contraction)

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
                fan_in = 4 if name.endswith("Dense_0/kernel") else (8 if name.endswith("lll_weight/kernel") else D)
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

def make_network(system: System, network: Network) -> B200Network:
    """networks/__init__.py:22-37."""
    if str(network.type) == "laughlin":
        return Laughlin(system.nspins, system.flux, excitation_lz=getattr(system, "lz_center", 0.0) or 0.0)
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
