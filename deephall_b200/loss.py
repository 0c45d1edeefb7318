"""Energy statistics and gradient with the reference's factory signature (deephall/loss.py).

`make_loss_fn(network, system, mode)` returns `loss_and_grad(params, data)` as loss.py:47-110.
The per-walker parameter gradients the reference materialises (loss.py:53-58, B x P x 2
floats) are replaced by ONE vector-Jacobian product with cotangent (2/B_valid) * diff_b
(dh_logpsi_vjp), which is the same number (loss.py:60-64,99-106).  The statistics themselves
(iqr_clip, nanmean, nanquantile: loss.py:30-38,66-92) run as two kernels of the library
(dh_energy_stats, dh_energy_diff) around ONE packed all-reduce.
"""
from __future__ import annotations

import enum

import torch

from . import _native, constants
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

    def vjp_over(params, data, cot, ok):
        """One VJP with the cotangents `cot` (already zero for walkers outside `ok`); those walkers' coordinates are replaced
        by an ok walker's, so that no 0 * NaN reaches the batched reverse pass (stream-ordered, no host synchronisation)."""
        first = ok.to(torch.int8).argmax()
        data = torch.where(ok[:, None, None], data, data.index_select(0, first[None])).contiguous()
        return torch.nan_to_num(net.plan(system).logpsi_vjp(params, data, cot.contiguous()))

    def loss_and_grad(params: torch.Tensor, data: torch.Tensor):
        el, obs = batch_local_energy(params, data)  # loss.py:67
        lp = batch_local_energy.last_logpsi
        if el.shape[0] > _native.ENERGY_STATS_MAX_BATCH:
            return loss_and_grad_tensor_ops(params, data, el, obs, lp)
        # loss.py:68-74,79-91: every rank-local mean in ONE kernel and ONE packed all-reduce (dh_energy_stats), then the
        # clipped differences and the gradient's cotangents in a second kernel (dh_energy_diff)
        red = constants.pmean(_native.energy_stats(el, obs))
        diff, cot, ok, _counts = _native.energy_diff(el, obs, red, system.lz_penalty or 0.0, system.lz_center or 0.0,
                                                     system.l2_penalty or 0.0, logpsi=lp)
        stats = {
            "angular_momentum_z": red[3], "angular_momentum_z_square": red[4], "angular_momentum_square": red[5],
            "potential": red[2], "kinetic": torch.complex(red[0], red[1]),
            "energy": torch.complex(red[6], red[7]), "variance": red[10] - red[6] * red[6],  # loss.py:91 (pmean is linear)
        }
        if mode == LossMode.ENERGY_DIFF:
            return stats, diff
        # loss.py:60-64,99-106: 2 * nanmean_b[(dRe_b - i dIm_b) diff_b], nan_to_num.  Its real part
        # 2 * nanmean_b[dRe_b Re(diff_b) + dIm_b Im(diff_b)] is one VJP with cotangent (2/B_ok)(Re, Im) diff_b.
        okb = ok > 0
        grads = vjp_over(params, data, cot, okb)
        if mode == LossMode.SR_F_VECTOR:
            # loss.py:107-108 keeps the complex vector; its imaginary part 2 * nanmean_b[dRe_b Im(diff_b) - dIm_b Re(diff_b)]
            # is a second VJP with cotangent (2/B)(Im, -Re) diff_b
            cot_i = torch.stack([cot[:, 1], -cot[:, 0]], dim=-1)
            return stats, torch.complex(grads, vjp_over(params, data, cot_i, okb))
        return stats, grads

    def loss_and_grad_tensor_ops(params, data, el, obs, lp):
        """The same statistics with tensor ops: shards of more than 32768 walkers per rank (beyond the kernels' range)."""
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
        d = torch.view_as_real(diff)
        valid = ~torch.isnan(d).any(-1)
        ok = valid & torch.isfinite(lp.real) & torch.isfinite(lp.imag)
        cot = torch.where(ok[:, None], d, torch.zeros_like(d)) * (2.0 / ok.sum().clamp(min=1).to(torch.float32))
        grads = vjp_over(params, data, cot, ok)
        if mode == LossMode.SR_F_VECTOR:
            cot_i = torch.stack([cot[:, 1], -cot[:, 0]], dim=-1)
            return stats, torch.complex(grads, vjp_over(params, data, cot_i, ok))
        return stats, grads

    return loss_and_grad
