"""Torch-CPU restatement of deephall/mcmc.py and train.py:40-54 (TEST INFRASTRUCTURE ONLY).

The reference draws from jax threefry, which cannot be reproduced here; every function
therefore takes the random numbers as explicit arguments ("injected randoms"):
``normal`` ~ N(0,1) and ``uphi`` ~ U[0,1) of shape (B,N) and ``uacc`` ~ U[0,1) of shape (B,).
"""
from __future__ import annotations

import math

import torch


def init_guess(gen: torch.Generator, batch: int, nelec: int, dtype=torch.float32):
    # train.py:40-54: theta = arccos(U(-1,1)), phi = U(-pi,pi)
    u1 = torch.rand((batch, nelec), generator=gen, dtype=torch.float64) * 2 - 1
    u2 = (torch.rand((batch, nelec), generator=gen, dtype=torch.float64) * 2 - 1) * math.pi
    return torch.stack([torch.arccos(u1), u2], dim=-1).to(dtype)


def sph_sampling(x1, normal, uphi, stddev):
    # mcmc.py:67-102
    theta, phi = x1[..., 0], x1[..., 1]
    theta_p = torch.arctan(normal * stddev)
    phi_p = uphi * 2 * math.pi
    xp = torch.sin(theta_p) * torch.cos(phi_p)
    yp = torch.sin(theta_p) * torch.sin(phi_p)
    zp = torch.cos(theta_p)
    # R_z(phi) R_y(theta) (xp, yp, zp)
    X = torch.cos(theta) * xp + torch.sin(theta) * zp
    Y = yp
    Z = -torch.sin(theta) * xp + torch.cos(theta) * zp
    x2 = torch.cos(phi) * X - torch.sin(phi) * Y
    y2 = torch.sin(phi) * X + torch.cos(phi) * Y
    z2 = Z
    theta2 = torch.arccos(torch.clip(z2, -1, 1))
    phi2 = torch.sign(y2) * torch.arccos(torch.clip(x2 / torch.sin(theta2), -1, 1))
    return torch.stack([theta2, phi2], dim=-1)


def log_uniform(uacc):
    """log(U) rounded once from fp64 -- the convention both sides use so that the accept
    test is bit-reproducible (include/deephall_b200.h, dh_mcmc_accept)."""
    return torch.log(uacc.double()).to(uacc.dtype)


def mh_accept(lp_1, lp_2, uacc):
    # mcmc.py:56-59: strict >, NaN proposal -> reject
    return (lp_2 - lp_1) > log_uniform(uacc)


def mh_update(batch_f, x1, lp_1, normal, uphi, uacc, stddev):
    # mcmc.py:25-64
    x2 = sph_sampling(x1, normal, uphi, stddev)
    lp_2 = 2.0 * batch_f(x2).real
    cond = mh_accept(lp_1, lp_2, uacc)
    x_new = torch.where(cond[..., None, None], x2, x1)
    lp_new = torch.where(cond, lp_2, lp_1)
    return x_new, lp_new, cond


def mcmc_step(batch_f, data, randoms, width):
    """mcmc.py:122-148.  randoms: list of (normal, uphi, uacc) per MH move."""
    lp = 2.0 * batch_f(data).real
    naccept = 0
    for normal, uphi, uacc in randoms:
        data, lp, cond = mh_update(batch_f, data, lp, normal, uphi, uacc, width)
        naccept += int(cond.sum())
    pmove = naccept / (len(randoms) * data.shape[0])
    return data, pmove


def draw_randoms(gen: torch.Generator, steps: int, batch: int, nelec: int, dtype=torch.float32):
    out = []
    for _ in range(steps):
        out.append(
            (
                torch.randn((batch, nelec), generator=gen, dtype=dtype),
                torch.rand((batch, nelec), generator=gen, dtype=dtype),
                torch.rand((batch,), generator=gen, dtype=dtype),
            )
        )
    return out


def update_mcmc_width(t, width, adapt_frequency, pmove, pmoves, pmove_max=0.55, pmove_min=0.5):
    # mcmc.py:153-186
    t_since = t % adapt_frequency
    pmoves[t_since] = pmove
    if t > 0 and t_since == 0:
        m = sum(pmoves) / len(pmoves)
        if m > pmove_max:
            width *= 1.1
        elif m < pmove_min:
            width /= 1.1
    return width, pmoves
