"""Torch-CPU restatement of deephall/networks/laughlin.py:59-100 (ground state, quasihole, quasiparticle).

TEST INFRASTRUCTURE ONLY.  Used as an analytic oracle: the reference pins
energy = 2.58... and L^2 = 0 for nspins [3,0], flux 6 (tests/cli_test.py:41-42).
"""
from __future__ import annotations

import torch

from .psiformer import slogdet_tail, spinors


def laughlin_orbitals(x, flux: int, cf_flux: int = 1, excitation_lz: float = 0.0):
    # laughlin.py:35-37,59-71 (full_orbitals) and :73-83 (quasihole_orbitals)
    N = x.shape[-2]
    Q1 = flux / 2 - cf_flux * (N - 1)
    u, v = spinors(x)
    u, v = u[..., None], v[..., None]
    twoQ1 = int(round(2 * Q1))
    if N == 2 * Q1 + 1:  # ground state: m = -Q .. Q  ->  exponents a = Q + m = 0 .. 2Q
        a = torch.arange(0, twoQ1 + 1)
    elif N == 2 * Q1:  # quasihole: m = -Q .. -lz-1, then Q, Q-1, .. -lz+1  (the orbital m = -lz is left out)
        skip = Q1 - excitation_lz
        assert abs(skip - round(skip)) < 1e-9 and -abs(Q1) <= excitation_lz <= abs(Q1)  # laughlin.py:49-52,39
        skip = int(round(skip))
        a = torch.cat([torch.arange(0, skip), torch.arange(twoQ1, skip, -1)])
    elif N == 2 * Q1 + 2:  # quasiparticle (laughlin.py:85-100): the filled shell plus one LLL-projected orbital
        return quasiparticle_orbitals(u, v, Q1, excitation_lz)
    else:
        raise ValueError("Filling not supported")
    eye = torch.eye(N, dtype=u.dtype)
    element = u * v[..., :, 0][..., None, :] - u[..., :, 0][..., None, :] * v + eye
    jas = element.prod(-1, keepdim=True)
    return u**a * v ** (twoQ1 - a) * jas


def quasiparticle_orbitals(u, v, Q, m1):
    # laughlin.py:85-100; u, v: [N, 1]
    assert abs((m1 - Q) - round(m1 - Q)) < 1e-9 and -abs(Q) - 1 <= m1 <= abs(Q) + 1  # laughlin.py:49-52,43
    N = u.shape[-2]
    twoQ = int(round(2 * Q))
    a = torch.arange(0, twoQ + 1)
    orbitals = u**a * v ** (twoQ - a)
    eye = torch.eye(N, dtype=u.dtype)
    u_row, v_row = u[..., :, 0][..., None, :], v[..., :, 0][..., None, :]
    element = u * v_row - u_row * v + eye
    jastrow = element.prod(-1, keepdim=True)
    # LLL projection (u* -> d/du, v* -> d/dv)
    jastrow_dv = jastrow * ((-u_row / element).sum(-1, keepdim=True) + u)
    jastrow_du = jastrow * ((v_row / element).sum(-1, keepdim=True) - v)
    ea, eb = int(round(Q + m1)), int(round(Q - m1))
    excited = (u**ea * v**eb) * ((Q + 1 + m1) * v * jastrow_dv - (Q + 1 - m1) * u * jastrow_du)
    return torch.cat([orbitals * jastrow, excited], -1)


def logpsi(x, flux: int, cf_flux: int = 1, excitation_lz: float = 0.0):
    return slogdet_tail(laughlin_orbitals(x, flux, cf_flux, excitation_lz)[..., None, :, :])
