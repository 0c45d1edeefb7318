"""Torch-CPU restatement of deephall/hamiltonian.py (TEST INFRASTRUCTURE ONLY).

This is the reference's *algorithm*: complex gradient and full (2N x 2N) Hessian of
log psi in (theta, phi) coordinates (hamiltonian.py:105-114), then the kinetic energy
(hamiltonian.py:121-133) and L^2 / L_z / L_z^2 (hamiltonian.py:139-169).

torch 2.11 caveat: vmap(func.hessian) / jacfwd(jacfwd) are silently wrong
on this graph.  We use reverse-over-reverse, one walker at a time through
``torch.func.vmap(jacrev(jacrev(f)))``.
"""
from __future__ import annotations

import math

import torch
from torch.func import jacrev, vmap


def coulomb_potential(cos12, r):
    # hamiltonian.py:27-41
    r_ee = torch.sqrt(2 - 2 * cos12)
    return torch.triu(1 / r_ee, diagonal=1).sum((-1, -2)) / r


def harmonic_potential(cos12, Q):
    # hamiltonian.py:44-60
    return torch.triu(1 + (Q + 1) / Q * cos12, diagonal=1).sum((-1, -2))


def potential(x, Q, r, interaction_type="coulomb"):
    # hamiltonian.py:63-80 ; x is (..., N, 2)
    theta, phi = x[..., 0], x[..., 1]
    xyz = torch.stack([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), torch.cos(theta)], dim=-1)
    cos12 = torch.einsum("...ia,...ja->...ij", xyz, xyz)
    if interaction_type == "coulomb":
        return coulomb_potential(cos12, r)
    return harmonic_potential(cos12, Q)


def _assemble(x, g, hess, Q, r):
    """hamiltonian.py:96-170 given complex gradient g (N,2) and Hessian (N,2,N,2)."""
    theta, phi = x[..., 0], x[..., 1]
    sin, cos, tan = torch.sin, torch.cos, torch.tan
    grad_theta, grad_phi = g[..., 0], g[..., 1]
    square_grad = (grad_theta**2 + grad_phi**2 / sin(theta) ** 2).sum()
    lap = (
        grad_theta / tan(theta)
        + torch.diagonal(hess[:, 0, :, 0])
        + torch.diagonal(hess[:, 1, :, 1]) / sin(theta) ** 2
    ).sum()
    mag = ((Q / tan(theta)) ** 2 + 2j * Q * cos(theta) / sin(theta) ** 2 * grad_phi).sum()
    kinetic = (-lap - square_grad + mag) / 2 / r**2

    r_hat = torch.stack([sin(theta) * cos(phi), sin(theta) * sin(phi), cos(theta)])
    phi_hat = torch.stack([-sin(phi), cos(phi), torch.zeros_like(phi)])
    theta_hat_p = torch.stack([cos(phi) / tan(theta), sin(phi) / tan(theta), -torch.ones_like(theta)])
    i = (Ellipsis, slice(None), None)
    j = (Ellipsis, None, slice(None))
    h_tt = hess[:, 0, :, 0] + grad_theta[i] * grad_theta[j]
    h_tp = hess[:, 0, :, 1] + grad_theta[i] * grad_phi[j]
    h_pp = hess[:, 1, :, 1] + grad_phi[i] * grad_phi[j]
    magnetic = Q * (theta_hat_p * cos(theta) + r_hat)
    l2 = (
        2 * phi_hat[i] * theta_hat_p[j] * h_tp
        - phi_hat[i] * phi_hat[j] * h_tt
        - theta_hat_p[i] * theta_hat_p[j] * h_pp
        - (2j * magnetic[j]) * (phi_hat[i] * grad_theta[i] - theta_hat_p[i] * grad_phi[i])
        + magnetic[i] * magnetic[j]
    ).sum() - (grad_theta / tan(theta)).sum()
    return kinetic, {
        "angular_momentum_z": grad_phi.sum().imag,
        "angular_momentum_z_square": -h_pp.sum().real,
        "angular_momentum_square": l2.real,
    }


def make_local_kinetic_energy(f, Q, r):
    """f(x[N,2]) -> complex scalar.  Returns ke(x[N,2]) -> (kinetic, angular momenta)."""

    def f_re(x):
        return f(x).real

    def f_im(x):
        return f(x).imag

    def ke(x):
        g = torch.complex(jacrev(f_re)(x), jacrev(f_im)(x))
        h = torch.complex(jacrev(jacrev(f_re))(x), jacrev(jacrev(f_im))(x))
        return _assemble(x, g, h, Q, r)

    return ke


def batch_local_energy(f, x, Q, r=None, interaction_strength=1.0, interaction_type="coulomb", chunk=64):
    """vmap of hamiltonian.py:193-210 over walkers.  f: single-walker log psi."""
    r = r or math.sqrt(Q)

    def f_re(xx):
        return f(xx).real

    def f_im(xx):
        return f(xx).imag

    outs = {k: [] for k in ("energy", "kinetic", "potential", "angular_momentum_z", "angular_momentum_z_square", "angular_momentum_square")}
    for s in range(0, x.shape[0], chunk):
        xs = x[s : s + chunk]
        g = torch.complex(vmap(jacrev(f_re))(xs), vmap(jacrev(f_im))(xs))
        h = torch.complex(vmap(jacrev(jacrev(f_re)))(xs), vmap(jacrev(jacrev(f_im)))(xs))
        pot = potential(xs, Q, r, interaction_type) * interaction_strength
        for b in range(xs.shape[0]):
            kin, am = _assemble(xs[b], g[b], h[b], Q, r)
            outs["kinetic"].append(kin)
            outs["potential"].append(pot[b])
            outs["energy"].append(kin + pot[b])
            for k, v in am.items():
                outs[k].append(v)
    return {k: torch.stack(v) for k, v in outs.items()}
