"""Local energy with the reference's factory signatures (deephall/hamiltonian.py).

`local_energy(f, system)` returns `_e_l(params, data) -> (E_L, OtherObservables)` exactly as
hamiltonian.py:175-212 does, but instead of differentiating `f` (two grads + two full
Hessians, hamiltonian.py:105-113) it recognises that `f` is a deephall_b200 network and runs
the fused forward-Laplacian kernels (dh_local_energy).
"""
from __future__ import annotations

import torch

from .config import System
from .networks import B200Network as Psiformer  # any network of this engine


def _network_of(f) -> Psiformer:
    net = getattr(f, "__self__", f)
    if not isinstance(net, Psiformer):
        raise TypeError("deephall_b200.hamiltonian needs `f` to be `model.apply` of a deephall_b200 network "
                        "(there is no generic autodiff path and no CPU fallback)")
    return net


def make_potential(system: System, network: Psiformer):
    """hamiltonian.py:63-80: potential(data) WITHOUT the interaction_strength factor."""
    def potential(data: torch.Tensor) -> torch.Tensor:
        single = data.dim() == 2
        out = network.plan(system).potential((data[None] if single else data).contiguous().float())
        return out[0] if single else out

    return potential


def local_energy(f, system: System):
    """hamiltonian.py:175-212."""
    net = _network_of(f)

    def _e_l(params: torch.Tensor, data: torch.Tensor):
        single = data.dim() == 2
        out = net.plan(system).local_energy(params, (data[None] if single else data).contiguous().float())
        if single:
            out = {k: v[0] for k, v in out.items()}
        energy = out.pop("energy")
        _e_l.last_logpsi = out.pop("logpsi")  # by-product of the same pass (loss.py masks non-finite walkers with it)
        return energy, out  # keys: angular_momentum_z, _z_square, _square, potential, kinetic

    _e_l.last_logpsi = None
    return _e_l


def make_local_kinetic_energy(f, system: System):
    """hamiltonian.py:83-172 analogue: (kinetic, AngularMomenta) per walker."""
    e_l = local_energy(f, system)

    def _lapl_over_f(params, data):
        _, obs = e_l(params, data)
        kin = obs["kinetic"]
        return kin, {k: obs[k] for k in ("angular_momentum_z", "angular_momentum_z_square", "angular_momentum_square")}

    return _lapl_over_f
