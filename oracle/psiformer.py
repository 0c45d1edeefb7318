"""Torch-CPU restatement of the reference Psiformer (TEST INFRASTRUCTURE ONLY).

Follows, line by line:
  deephall/networks/psiformer.py:32-60   PsiformerLayers (features, Dense, MHA, LN, tanh)
  deephall/networks/psiformer.py:63-91   Psiformer.__call__/orbitals (slogdet tail)
  deephall/networks/blocks.py:23-35      FeaturedOrbitals (two real DenseGeneral -> complex)
  deephall/networks/blocks.py:38-70      Orbitals (monopole-harmonic envelope, `full`)
  deephall/networks/blocks.py:73-121     Jastrow
  deephall/networks/__init__.py:22-37    make_network constructor contract
and the published semantics of flax 0.10.2 (Dense, DenseGeneral, MultiHeadAttention,
LayerNorm(use_fast_variance=True)) and jax 0.4.35 (`jnp.linalg.slogdet`), neither of
which is vendored in /root/reference (pyproject.toml:16-23).

All functions broadcast over leading batch dimensions: ``x`` is ``(..., N, 2)``.
Parameters are a flat ``dict[str, Tensor]`` keyed by the flax tree path, e.g.
``PsiformerLayers_0/MultiHeadAttention_0/query/kernel``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
from scipy import special as ss


@dataclass(frozen=True)
class NetCfg:
    """Arguments `make_network` passes to `Psiformer` (networks/__init__.py:28-36)."""

    nspins: tuple[int, int] = (3, 0)
    flux: int = 2
    ndets: int = 1
    num_heads: int = 4
    heads_dim: int = 64
    num_layers: int = 2
    orbital_type: str = "full"  # config.py:87-89: "full" | "sparse"

    @property
    def Q(self) -> float:
        return self.flux / 2

    @property
    def nelec(self) -> int:
        return sum(self.nspins)

    @property
    def norb(self) -> int:  # L = 2Q + 1 lowest-Landau-level orbitals
        return int(self.flux) + 1

    @property
    def dim(self) -> int:
        return self.num_heads * self.heads_dim


def param_shapes(cfg: NetCfg) -> "OrderedDict[str, tuple[int, ...]]":
    """The flax parameter tree (auto-naming order), flattened to path -> shape."""
    D, H, hd, N, L, K = cfg.dim, cfg.num_heads, cfg.heads_dim, cfg.nelec, cfg.norb, cfg.ndets
    s: OrderedDict[str, tuple[int, ...]] = OrderedDict()
    p = "PsiformerLayers_0/"
    s[p + "Dense_0/kernel"] = (4, D)
    for l in range(cfg.num_layers):
        a = f"{p}MultiHeadAttention_{l}/"
        for nm in ("query", "key", "value"):
            s[a + nm + "/kernel"] = (D, H, hd)
            s[a + nm + "/bias"] = (H, hd)
        s[a + "out/kernel"] = (H, hd, D)
        s[a + "out/bias"] = (D,)
        s[f"{p}Dense_{1 + 2 * l}/kernel"] = (D, D)
        s[f"{p}LayerNorm_{2 * l}/scale"] = (D,)
        s[f"{p}LayerNorm_{2 * l}/bias"] = (D,)
        s[f"{p}Dense_{2 + 2 * l}/kernel"] = (D, D)
        s[f"{p}Dense_{2 + 2 * l}/bias"] = (D,)
        s[f"{p}LayerNorm_{2 * l + 1}/scale"] = (D,)
        s[f"{p}LayerNorm_{2 * l + 1}/bias"] = (D,)
    o = "Orbitals_0/featured_orbitals/"
    idx = 0
    F = L if cfg.orbital_type == "full" else 8  # blocks.py:47-56: sparse orbitals project to 8 features first
    for n_alpha in cfg.nspins:  # blocks.py:29-34: one (re, im) pair per non-empty spin block
        if n_alpha:
            for _ in range(2):
                s[f"{o}DenseGeneral_{idx}/kernel"] = (D, F, N, K)
                s[f"{o}DenseGeneral_{idx}/bias"] = (F, N, K)
                idx += 1
    if cfg.orbital_type == "sparse":  # blocks.py:57: nn.DenseGeneral(2Q+1, axis=1), real parameters on complex input
        s["Orbitals_0/lll_weight/kernel"] = (8, L)
        s["Orbitals_0/lll_weight/bias"] = (L,)
    n_up, n_dn = cfg.nspins
    if n_up * (n_up - 1) // 2 + n_dn * (n_dn - 1) // 2 > 0:  # blocks.py:91
        s["Jastrow_0/ee_par"] = (1,)
    if n_up > 0:  # blocks.py:99: r_ees[0][1] has shape (n_up, n_dn) and the test is on shape[0] -> inert leaf when n_dn = 0
        s["Jastrow_0/ee_anti"] = (1,)
    return s


def num_params(cfg: NetCfg) -> int:
    return sum(int(np.prod(v)) for v in param_shapes(cfg).values())


def _lecun_normal(gen: torch.Generator, shape, fan_in: int) -> torch.Tensor:
    # flax default_kernel_init = variance_scaling(1.0, "fan_in", "truncated_normal"):
    # stddev = sqrt(1/fan_in) / 0.87962566103423978, truncated at +-2 (in unit-normal units).
    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
    out = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(out, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
    return out * std


def init_params(cfg: NetCfg, seed: int = 0, dtype=torch.float64, perturb: float = 0.0):
    """`model.init` distributions (lecun-normal kernels, zero biases, unit LN scale,
    ee_par = ee_anti = 1).  ``perturb`` adds N(0, perturb^2) to biases/scales/jastrow so
    that parity tests exercise every term."""
    gen = torch.Generator().manual_seed(seed)
    D = cfg.dim
    out: OrderedDict[str, torch.Tensor] = OrderedDict()
    for name, shape in param_shapes(cfg).items():
        if name.endswith("/kernel"):
            if name.endswith("Dense_0/kernel"):
                fan_in = 4
            elif name.endswith("lll_weight/kernel"):
                fan_in = 8
            else:
                fan_in = D  # every other kernel contracts a D-dim (or HxHd = D) input
            t = _lecun_normal(gen, shape, fan_in)
        elif name.endswith("/scale") or name.startswith("Jastrow_0/"):
            t = torch.ones(shape, dtype=torch.float64)
        else:
            t = torch.zeros(shape, dtype=torch.float64)
        if perturb and not name.endswith("/kernel"):
            t = t + perturb * torch.randn(shape, generator=gen, dtype=torch.float64)
        out[name] = t.to(dtype)
    return out


def cast_params(params, dtype):
    return OrderedDict((k, v.to(dtype)) for k, v in params.items())


# ----------------------------------------------------------------------------- flax layers
def layer_norm(x, scale, bias, eps=1e-5):
    # flax LayerNorm(epsilon=1e-5, use_fast_variance=True)
    mu = x.mean(-1, keepdim=True)
    var = torch.clamp((x * x).mean(-1, keepdim=True) - mu * mu, min=0.0)
    return (x - mu) * (torch.rsqrt(var + eps) * scale) + bias


def multi_head_attention(params, prefix, h, cfg: NetCfg):
    # flax MultiHeadAttention(num_heads=H): qkv_features = out_features = D, biases on.
    H, hd = cfg.num_heads, cfg.heads_dim
    q = torch.einsum("...nd,dhe->...nhe", h, params[prefix + "query/kernel"]) + params[prefix + "query/bias"]
    k = torch.einsum("...nd,dhe->...nhe", h, params[prefix + "key/kernel"]) + params[prefix + "key/bias"]
    v = torch.einsum("...nd,dhe->...nhe", h, params[prefix + "value/kernel"]) + params[prefix + "value/bias"]
    q = q / math.sqrt(hd)
    w = torch.einsum("...qhd,...khd->...hqk", q, k)
    w = torch.softmax(w, dim=-1)
    o = torch.einsum("...hqk,...khd->...qhd", w, v)
    return torch.einsum("...qhd,hde->...qe", o, params[prefix + "out/kernel"]) + params[prefix + "out/bias"]


def input_feature(x, cfg: NetCfg):
    # psiformer.py:51-60, spins psiformer.py:81
    theta, phi = x[..., 0], x[..., 1]
    spins = torch.tensor([1.0] * cfg.nspins[0] + [-1.0] * cfg.nspins[1], dtype=x.dtype)
    spins = spins.expand(theta.shape)
    return torch.stack(
        [torch.cos(theta), torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), spins], dim=-1
    )


def psiformer_layers(params, x, cfg: NetCfg):
    # psiformer.py:37-49
    p = "PsiformerLayers_0/"
    h = input_feature(x, cfg) @ params[p + "Dense_0/kernel"]
    for l in range(cfg.num_layers):
        attn = multi_head_attention(params, f"{p}MultiHeadAttention_{l}/", h, cfg)
        h = h + attn @ params[f"{p}Dense_{1 + 2 * l}/kernel"]
        h = layer_norm(h, params[f"{p}LayerNorm_{2 * l}/scale"], params[f"{p}LayerNorm_{2 * l}/bias"])
        h = h + torch.tanh(h @ params[f"{p}Dense_{2 + 2 * l}/kernel"] + params[f"{p}Dense_{2 + 2 * l}/bias"])
        h = layer_norm(h, params[f"{p}LayerNorm_{2 * l + 1}/scale"], params[f"{p}LayerNorm_{2 * l + 1}/bias"])
    return h


def norm_factor(cfg: NetCfg, dtype=torch.float64):
    # blocks.py:45-46
    Q = cfg.Q
    m = np.arange(-Q, Q + 1)
    return torch.tensor(np.sqrt(ss.comb(2 * Q, Q - m)), dtype=dtype)


def spinors(x):
    # blocks.py:65-66
    theta, phi = x[..., 0], x[..., 1]
    cdt = torch.complex128 if x.dtype == torch.float64 else torch.complex64
    ph = torch.polar(torch.ones_like(phi), 0.5 * phi).to(cdt)
    u = torch.cos(theta / 2) * ph
    v = torch.sin(theta / 2) * ph.conj()
    return u, v


def envelope(x, cfg: NetCfg):
    # blocks.py:64-67: sqrt(C(2Q,Q-m)) u^(Q+m) v^(Q-m); exponents are the integers 0..2Q.
    u, v = spinors(x)
    twoQ = int(cfg.flux)
    a = torch.arange(0, twoQ + 1)  # Q + m
    b = twoQ - a  # Q - m
    return norm_factor(cfg, x.dtype) * u[..., None] ** a * v[..., None] ** b  # (..., N, L)


def featured_orbitals(params, h, cfg: NetCfg):
    # blocks.py:28-35
    o = "Orbitals_0/featured_orbitals/"
    outs, idx, start = [], 0, 0
    for n_alpha in cfg.nspins:
        if n_alpha:
            ha = h[..., start : start + n_alpha, :]
            re = torch.einsum("...nd,dljk->...nljk", ha, params[f"{o}DenseGeneral_{idx}/kernel"]) + params[f"{o}DenseGeneral_{idx}/bias"]
            im = torch.einsum("...nd,dljk->...nljk", ha, params[f"{o}DenseGeneral_{idx + 1}/kernel"]) + params[f"{o}DenseGeneral_{idx + 1}/bias"]
            outs.append(torch.complex(re, im))
            idx += 2
        start += n_alpha
    c = torch.cat(outs, dim=-4)  # (..., N, L or 8, N, K)
    if cfg.orbital_type == "sparse":  # blocks.py:61-62: contract the 8 features with the (8, L) kernel, add the (real) bias
        w, b = params["Orbitals_0/lll_weight/kernel"], params["Orbitals_0/lll_weight/bias"]
        c = torch.einsum("...nsjk,sl->...nljk", c, w.to(c.dtype)) + b.to(c.dtype)[:, None, None]
    return c  # (..., N, L, N, K)


def jastrow(params, x, cfg: NetCfg):
    # blocks.py:77-121
    theta, phi = x[..., 0], x[..., 1]
    cart = torch.stack([torch.cos(theta), torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi)], dim=-1)
    N = cfg.nelec
    diff = cart[..., None, :, :] - cart[..., :, None, :]
    eye = torch.eye(N, dtype=x.dtype)
    r_ee = torch.linalg.norm(diff + eye[..., None], dim=-1) * (1.0 - eye)
    n_up, n_dn = cfg.nspins
    total = torch.zeros(x.shape[:-2], dtype=x.dtype)
    iu = torch.triu_indices(n_up, n_up, offset=1)
    idn = torch.triu_indices(n_dn, n_dn, offset=1)
    par = torch.cat([r_ee[..., iu[0], iu[1]], r_ee[..., n_up + idn[0], n_up + idn[1]]], dim=-1)
    if par.shape[-1] > 0:
        a = params["Jastrow_0/ee_par"]
        total = total + (-(0.25 * a**2) / (a + par)).sum(-1)
    if n_up > 0:  # blocks.py:99 (shape[0] of the (n_up, n_dn) block); an empty sum when n_dn = 0
        a = params["Jastrow_0/ee_anti"]
        anti = r_ee[..., :n_up, n_up:]
        total = total + (-(0.5 * a**2) / (a + anti)).sum((-1, -2))
    return total


def orbitals(params, x, cfg: NetCfg):
    """psiformer.py:78-91 -> (..., K, N, N): rows = electrons, columns = orbitals."""
    h = psiformer_layers(params, x, cfg)
    c = featured_orbitals(params, h, cfg)  # (..., N, L, N, K)
    env = envelope(x, cfg)  # (..., N, L)
    orb = (c * env[..., None, None]).sum(-3)  # (..., N, N, K)
    orb = torch.movedim(orb, -1, -3)  # blocks.py:70
    jas = jastrow(params, x, cfg)
    return torch.exp(jas / cfg.nelec)[..., None, None, None] * orb


def slogdet_tail(orb):
    """psiformer.py:74-76 == laughlin.py:55-57 : complex log of a sum of determinants."""
    signs, logdets = torch.linalg.slogdet(orb)
    logmax = logdets.max(dim=-1, keepdim=True).values
    tot = (signs * torch.exp(logdets - logmax)).sum(-1)
    return torch.log(tot) + logmax[..., 0]


def logpsi(params, x, cfg: NetCfg):
    """`model.apply(params, x)`: complex log psi; x is (..., N, 2)."""
    return slogdet_tail(orbitals(params, x, cfg))


# ----------------------------------------------------------------------------- flat layout
def flatten_params(params) -> torch.Tensor:
    return torch.cat([v.reshape(-1) for v in params.values()])


def unflatten_params(flat: torch.Tensor, cfg: NetCfg):
    out, off = OrderedDict(), 0
    for name, shape in param_shapes(cfg).items():
        n = int(np.prod(shape))
        out[name] = flat[off : off + n].reshape(shape)
        off += n
    assert off == flat.numel()
    return out
