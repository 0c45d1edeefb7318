"""Torch-CPU restatement of deephall/networks/laughlin.py:59-71 (ground state only).

TEST INFRASTRUCTURE ONLY.  Used as an analytic oracle: the reference pins
energy = 2.58... and L^2 = 0 for nspins [3,0], flux 6 (tests/cli_test.py:41-42).
"""
from __future__ import annotations

import torch

from .psiformer import slogdet_tail, spinors


def laughlin_orbitals(x, flux: int, cf_flux: int = 1):
    # laughlin.py:35-37,59-71 (full_orbitals)
    N = x.shape[-2]
    Q1 = flux / 2 - cf_flux * (N - 1)
    assert N == 2 * Q1 + 1, "only the Laughlin ground state is restated"
    u, v = spinors(x)
    u, v = u[..., None], v[..., None]
    twoQ1 = int(round(2 * Q1))
    a = torch.arange(0, twoQ1 + 1)
    eye = torch.eye(N, dtype=u.dtype)
    element = u * v[..., :, 0][..., None, :] - u[..., :, 0][..., None, :] * v + eye
    jas = element.prod(-1, keepdim=True)
    return u**a * v ** (twoQ1 - a) * jas


def logpsi(x, flux: int):
    return slogdet_tail(laughlin_orbitals(x, flux)[..., None, :, :])
