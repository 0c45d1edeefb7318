"""Energy statistics and gradient with the reference's factory signature (deephall/loss.py).

`make_loss_fn(network, system, mode)` returns `loss_and_grad(params, data)` as loss.py:47-110.
The per-walker parameter gradients the reference materialises (loss.py:53-58, B x P x 2
floats) are replaced by ONE vector-Jacobian product with cotangent (2/B_valid) * diff_b
(dh_logpsi_vjp), which is the same number (loss.py:60-64,99-106).
"""
from __future__ import annotations

import enum

import torch

from . import constants
from .config import System
from .hamiltonian import local_energy
from .networks import B200Network as Psiformer  # any network of this engine


def iqr_clip_real(x: torch.Tensor, scale: float = 100.0) -> torch.Tensor:
    # loss.py:30-34 (jnp.nanquantile == linear interpolation == torch default)
    q = torch.nanquantile(x, torch.tensor([0.25, 0.75], device=x.device, dtype=x.dtype))
    iqr = q[1] - q[0]
    return torch.clamp(x, q[0] - scale * iqr, q[1] + scale * iqr)


def iqr_clip(x: torch.Tensor, scale: float = 100.0) -> torch.Tensor:
    # loss.py:37-38
    if not x.is_complex():
        return iqr_clip_real(x, scale)
    return torch.complex(iqr_clip_real(x.real, scale), iqr_clip_real(x.imag, scale))


def _nanmean(x: torch.Tensor) -> torch.Tensor:
    if x.is_complex():
        bad = torch.isnan(x.real) | torch.isnan(x.imag)
        return torch.where(bad, torch.zeros_like(x), x).sum() / (~bad).sum()
    return torch.nanmean(x)


class LossMode(enum.Enum):  # loss.py:41-44
    ENERGY_GRAD = enum.auto()
    ENERGY_DIFF = enum.auto()
    SR_F_VECTOR = enum.auto()


def make_loss_fn(network, system: System, mode: LossMode = LossMode.ENERGY_GRAD):
    net = getattr(network, "__self__", network)
    if not isinstance(net, Psiformer):
        raise TypeError("network must be `model.apply` of a deephall_b200 network")
    batch_local_energy = local_energy(net.apply, system)

    def masked_vjp(params, data, cot, valid, lp):
        """One VJP over the valid walkers only.  The reference's loss_prod is a per-parameter nanmean over walkers
        (loss.py:60-64): a walker whose log psi is not finite drops out.  Here such a walker gets a zero cotangent AND
        its coordinates replaced by those of a valid walker, so that no 0 * NaN reaches the batched reverse pass; all
        of it stream-ordered (no host synchronisation).  lp: log psi of the same walkers from the local-energy pass."""
        ok = valid & torch.isfinite(lp.real) & torch.isfinite(lp.imag)
        first = ok.to(torch.int8).argmax()
        data = torch.where(ok[:, None, None], data, data.index_select(0, first[None])).contiguous()
        # nanmean's denominator: the walkers that stay
        scale = (valid.sum().clamp(min=1) / ok.sum().clamp(min=1)).to(cot.dtype)
        cot = torch.where(ok[:, None], cot, torch.zeros_like(cot)) * scale
        return torch.nan_to_num(net.plan(system).logpsi_vjp(params, data, cot.contiguous()))

    def loss_and_grad(params: torch.Tensor, data: torch.Tensor):
        el, obs = batch_local_energy(params, data)  # loss.py:67
        lp = batch_local_energy.last_logpsi
        # loss.py:68-74,91: means, pmean'd in one packed all-reduce
        names = list(obs.keys())
        local = [obs[k].mean() for k in names] + [_nanmean(el), _nanmean(iqr_clip(el)), torch.nanmean(el.real**2)]
        red = constants.pmean_packed(local)
        stats = dict(zip(names, red[: len(names)]))
        loss, clipped_loss, mean_sq = red[len(names)], red[len(names) + 1], red[len(names) + 2]
        diff_to_clip = el - clipped_loss
        if system.lz_penalty:  # loss.py:76-84
            lz2, lz = obs["angular_momentum_z_square"], obs["angular_momentum_z"]
            c_lz2, c_lz = constants.pmean_packed([torch.nanmean(iqr_clip(lz2)), torch.nanmean(iqr_clip(lz))])
            diff_to_clip = diff_to_clip + system.lz_penalty * ((lz2 - c_lz2) - 2 * system.lz_center * (lz - c_lz))
        if system.l2_penalty:  # loss.py:85-88
            l2 = obs["angular_momentum_square"]
            diff_to_clip = diff_to_clip + system.l2_penalty * (l2 - constants.pmean(torch.nanmean(iqr_clip(l2))))
        diff = iqr_clip(diff_to_clip)  # loss.py:89
        stats["energy"] = loss
        stats["variance"] = mean_sq - loss.real**2  # loss.py:91 (pmean is linear)
        if mode == LossMode.ENERGY_DIFF:
            return stats, diff
        # loss.py:60-64,99-106: 2 * nanmean_b[(dRe_b - i dIm_b) diff_b], nan_to_num.  Its real part
        # 2 * nanmean_b[dRe_b Re(diff_b) + dIm_b Im(diff_b)] is one VJP with cotangent (2/B)(Re, Im) diff_b.
        d = torch.view_as_real(diff)
        valid = ~torch.isnan(d).any(-1)
        nvalid = valid.sum().clamp(min=1).to(torch.float32)
        cot = torch.where(valid[:, None], d, torch.zeros_like(d)) * (2.0 / nvalid)
        grads = masked_vjp(params, data, cot, valid, lp)
        if mode == LossMode.SR_F_VECTOR:
            # loss.py:107-108 keeps the complex vector; its imaginary part 2 * nanmean_b[dRe_b Im(diff_b) - dIm_b Re(diff_b)]
            # is a second VJP with cotangent (2/B)(Im, -Re) diff_b
            cot_i = torch.stack([cot[:, 1], -cot[:, 0]], dim=-1).contiguous()
            grads_i = masked_vjp(params, data, cot_i, valid, lp)
            return stats, torch.complex(grads, grads_i)
        return stats, grads

    return loss_and_grad
