"""Torch-CPU restatement of deephall/loss.py (TEST INFRASTRUCTURE ONLY).

Single-shard version: `constants.pmean` (constants.py:40-41) over one device is the
identity; the multi-shard semantic (shard-local quantiles, pmean of the means) is
``loss_stats_sharded`` below.
"""
from __future__ import annotations

import torch
from torch.func import grad, vmap


def iqr_clip_real(x, scale=100.0):
    # loss.py:30-34 ; jnp.nanquantile == linear interpolation
    q1 = torch.nanquantile(x, 0.25)
    q3 = torch.nanquantile(x, 0.75)
    iqr = q3 - q1
    return torch.clip(x, q1 - scale * iqr, q3 + scale * iqr)


def iqr_clip(x, scale=100.0):
    # loss.py:37-38
    if not x.is_complex():
        return iqr_clip_real(x, scale)
    return torch.complex(iqr_clip_real(x.real, scale), iqr_clip_real(x.imag, scale))


def nanmean_c(x):
    if x.is_complex():
        # jnp.nanmean on complex: an element is NaN if either part is NaN
        bad = torch.isnan(x.real) | torch.isnan(x.imag)
        return torch.where(bad, torch.zeros_like(x), x).sum() / (~bad).sum()
    return torch.nanmean(x)


def loss_stats(el, obs, lz_penalty=0.0, lz_center=0.0, l2_penalty=0.0):
    """loss.py:66-92 for one shard.  el complex (B,), obs dict of (B,) tensors.
    Returns (stats, diff)."""
    stats = {k: v.mean() for k, v in obs.items()}  # loss.py:68-71 (plain mean)
    loss = nanmean_c(el)
    clipped = nanmean_c(iqr_clip(el))
    diff_to_clip = el - clipped
    if lz_penalty:
        lz2, lz = obs["angular_momentum_z_square"], obs["angular_momentum_z"]
        diff_to_clip = diff_to_clip + lz_penalty * (
            (lz2 - torch.nanmean(iqr_clip(lz2))) - 2 * lz_center * (lz - torch.nanmean(iqr_clip(lz)))
        )
    if l2_penalty:
        l2 = obs["angular_momentum_square"]
        diff_to_clip = diff_to_clip + l2_penalty * (l2 - torch.nanmean(iqr_clip(l2)))
    diff = iqr_clip(diff_to_clip)
    variance = torch.nanmean(el.real**2) - loss.real**2
    stats["energy"] = loss
    stats["variance"] = variance
    return stats, diff


def energy_grad(f_single, params_flat, x, diff, chunk=64):
    """loss.py:53-64,96-106 (ENERGY_GRAD).  f_single(params_flat, x[N,2]) -> complex.
    Per-walker parameter gradients are materialised exactly as the reference does."""
    B = x.shape[0]

    def f_re(p, xx):
        return f_single(p, xx).real

    def f_im(p, xx):
        return f_single(p, xx).imag

    acc = torch.zeros_like(params_flat)
    cnt = torch.zeros_like(params_flat)
    for s in range(0, B, chunk):
        xs, d = x[s : s + chunk], diff[s : s + chunk]
        gr = vmap(grad(f_re), in_dims=(None, 0))(params_flat, xs)
        gi = vmap(grad(f_im), in_dims=(None, 0))(params_flat, xs)
        prod = gr * d.real[:, None] + gi * d.imag[:, None]  # Re[(gr - i gi) * d]
        ok = ~torch.isnan(prod)
        acc += torch.where(ok, prod, torch.zeros_like(prod)).sum(0)
        cnt += ok.sum(0)
    return torch.nan_to_num(2 * acc / cnt)


def energy_grad_vjp(f_batch, params_flat, x, diff):
    """Same quantity by one reverse pass (valid when no walker is NaN)."""
    p = params_flat.detach().clone().requires_grad_(True)
    lp = f_batch(p, x)
    obj = (2.0 / x.shape[0]) * (lp.real * diff.real + lp.imag * diff.imag).sum()
    (g,) = torch.autograd.grad(obj, p)
    return g


def sr_f_vector(f_batch, params_flat, x, diff):
    """LossMode.SR_F_VECTOR (loss.py:99-108): the complex vector 2 * mean_b[(dRe_b - i dIm_b) diff_b], by two
    reverse passes (valid when no walker is NaN)."""
    p = params_flat.detach().clone().requires_grad_(True)
    lp = f_batch(p, x)
    s = 2.0 / x.shape[0]
    (gr,) = torch.autograd.grad(s * (lp.real * diff.real + lp.imag * diff.imag).sum(), p, retain_graph=True)
    (gi,) = torch.autograd.grad(s * (lp.real * diff.imag - lp.imag * diff.real).sum(), p)
    return torch.complex(gr, gi)
