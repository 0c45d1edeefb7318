"""KFAC training step with the reference's signature and settings (deephall/optimizers/kfac.py:198-241).

The reference hands the loss to `kfac_jax.Optimizer` (l2_reg=0, norm_constraint=1e-3, curvature_ema=0.95,
inverse_update_period=1, damping=1e-3, momentum=0, estimation_mode="fisher_exact", loss tag = unit normal predictive
distribution on Re log psi with kfac_jax's default variance 0.5, `loss.py:98`).  Here the curvature statistics come
from ONE extra forward + reverse pass of the CUDA path (`dh_kfac_factors`: the Kronecker-factor sums of every dense
layer as `repeated_dense` blocks over the electron axis, `optimizers/kfac.py:42-102`, plus diagonal blocks for the
LayerNorm and Jastrow parameters); the moving averages, the pi-adjusted damped Kronecker inverses and the
norm-constrained update run as small torch ops on the same device.  kfac_jax cannot be installed next to this code,
so the update rule is a restatement of the published algorithm (oracle/kfac.py states it in fp64; the GPU step is
tested against that) -- parity with the reference's optimizer trajectory is unpinned.
"""
from __future__ import annotations

import math
from typing import NamedTuple

import torch

from . import _native, constants
from .networks import Psiformer
from .optimizers import CheckpointState

VARIANCE = 0.5  # kfac_jax.register_normal_predictive_distribution(mean, targets=None, variance=0.5)


class KfacState(NamedTuple):
    step: int
    weight: float            # moving-average weight (de-biasing): w <- ema * w + 1
    stats: torch.Tensor      # moving average (raw) of the normalised curvature statistics, factor-vector layout
    dense0_xtx: torch.Tensor  # same for Dense_0's input factor (4 x 4), formed here from the walkers' features


def _features(data: torch.Tensor, n_up: int) -> torch.Tensor:
    """psiformer.py:51-60,81: (cos theta, sin theta cos phi, sin theta sin phi, spin) per electron."""
    theta, phi = data[..., 0], data[..., 1]
    spin = torch.ones_like(theta)
    spin[..., n_up:] = -1.0
    return torch.stack([torch.cos(theta), torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), spin], dim=-1)


_SIDE_STREAM = None


def _batched_inverses(groups: dict) -> dict:
    """{n: [n x n SPD matrices]} -> {n: (count, n, n) inverses}.

    dh_spd_inverse runs one Gauss-Jordan block per matrix, so a launch with 13 matrices leaves 135 SMs idle and two
    launches of 13 and 14 matrices take twice the time of one with 27.  All sizes up to 320 therefore go out as ONE
    batch: a smaller matrix is embedded as diag(A, I) in the largest size of the batch (its inverse is diag(A^-1, I)).
    The few larger factors (the orbital projection's 409 rows) are inverted by the library on a side stream at the
    same time."""
    global _SIDE_STREAM
    small = {n: ms for n, ms in groups.items() if n <= 320}
    large = {n: ms for n, ms in groups.items() if n > 320}
    out = {}
    cur = torch.cuda.current_stream()
    done = None
    if large:
        if _SIDE_STREAM is None:
            _SIDE_STREAM = torch.cuda.Stream()
        stacked = {n: torch.stack(ms) for n, ms in large.items()}
        _SIDE_STREAM.wait_stream(cur)
        with torch.cuda.stream(_SIDE_STREAM):
            for n, st in stacked.items():
                out[n] = torch.linalg.inv(st)
                out[n].record_stream(cur)
                st.record_stream(_SIDE_STREAM)
        done = _SIDE_STREAM
    if small:
        nmax = max(small)
        dev = next(iter(small.values()))[0].device
        total = sum(len(ms) for ms in small.values())
        batch = torch.zeros((total, nmax, nmax), dtype=torch.float32, device=dev)
        k = 0
        where = {}
        for n, ms in small.items():
            where[n] = (k, len(ms))
            batch[k : k + len(ms), :n, :n] = torch.stack(ms)
            if n < nmax:
                idx = torch.arange(n, nmax, device=dev)
                batch[k : k + len(ms), idx, idx] = 1.0
            k += len(ms)
        inv = _native.spd_inverse(batch)
        for n, (k0, cnt) in where.items():
            out[n] = inv[k0 : k0 + cnt, :n, :n]
    if done is not None:
        cur.wait_stream(done)
    return out


def make_kfac_training_step(optim_cfg, loss_grad_fn, network, system=None, norm_constraint=1e-3, curvature_ema=0.95,
                            damping=1e-3):
    """-> (init, step) like optimizers/kfac.py:198-241.  `network` is `model.apply` of a deephall_b200 Psiformer."""
    net = getattr(network, "__self__", network)
    if not isinstance(net, Psiformer):
        raise TypeError("network must be `model.apply` of a deephall_b200 Psiformer")
    plan = net.plan(system)
    layout, nfloats = plan.kfac_layout()
    n_up = net.nspins[0]
    t2 = 1.0 / VARIANCE  # squared loss tangent

    def normalised_stats(params, data):
        """This step's statistics (local shard), each already divided by its batch size."""
        B = data.shape[0]
        raw = plan.kfac_factors(params, data.contiguous())
        out = torch.zeros_like(raw)
        done = set()
        for e in layout:
            rows = float(B * e["rows_per_walker"])
            if e["kind"] == 0:
                for key, n in (("xtx_offset", e["in_dim"] ** 2), ("xsum_offset", e["in_dim"])):
                    o = e[key]
                    if o >= 0 and o not in done:
                        out[o : o + n] = raw[o : o + n] / rows
                        done.add(o)
                o, n = e["gtg_offset"], e["out_dim"] ** 2
                out[o : o + n] = raw[o : o + n] * (t2 / rows)
            elif e["kind"] == 1:  # per-walker squares
                o, n = e["diag_offset"], e["size"]
                out[o : o + n] = raw[o : o + n] * (t2 / B)
            else:  # naive diagonal: (batch-summed gradient)^2 / batch
                o, n = e["diag_offset"], e["size"]
                out[o : o + n] = raw[o : o + n] ** 2 * (t2 / B)
        feat = _features(data, n_up).reshape(-1, 4)
        return out, feat.T @ feat / feat.shape[0]

    def precondition(state: KfacState, grads: torch.Tensor) -> torch.Tensor:
        w = state.weight
        st = state.stats / w
        out = torch.zeros_like(grads)
        # pass 1: the damped, trace-normalised factors of every dense block, grouped by size so that each size is
        # inverted in ONE batched call (about 30 matrices of 256-410 rows per step)
        blocks, groups = [], {}
        for e in layout:
            ko = e["kernel_offset"]
            if e["kind"] != 0:
                n, o = e["size"], e["diag_offset"]
                out[ko : ko + n] = grads[ko : ko + n] / (st[o : o + n] + damping)
                continue
            din, dout, hb, npw = e["in_dim"], e["out_dim"], e["has_bias"], e["rows_per_walker"]
            if e["xtx_offset"] < 0:
                xtx = state.dense0_xtx / w
            else:
                xtx = st[e["xtx_offset"] : e["xtx_offset"] + din * din].view(din, din)
            if hb:
                xs = st[e["xsum_offset"] : e["xsum_offset"] + din]
                A = torch.empty(din + 1, din + 1, dtype=st.dtype, device=st.device)
                A[:din, :din] = xtx
                A[:din, din] = xs
                A[din, :din] = xs
                A[din, din] = 1.0
            else:
                A = xtx
            G = st[e["gtg_offset"] : e["gtg_offset"] + dout * dout].view(dout, dout)
            # (A (x) G + d I)^-1 ~ A_inv (x) G_inv, average-trace norms (kfac_jax.utils.pi_adjusted_kronecker_inverse);
            # the block is npw * A (x) G (fixed_scale, kfac.py:79-81), so d = damping / npw
            ca, cg = torch.trace(A) / A.shape[0], torch.trace(G) / G.shape[0]
            c = ca * cg
            ok = c > 0  # a factor that is still zero: plain damping
            d_hat = torch.where(ok, torch.sqrt((damping / npw) / torch.where(ok, c, torch.ones_like(c))), torch.ones_like(c))
            ck = torch.where(ok, torch.sqrt(torch.where(ok, c, torch.ones_like(c))), torch.full_like(c, math.sqrt(damping / npw)))
            ma = torch.where(ok, A / torch.where(ok, ca, torch.ones_like(ca)), torch.zeros_like(A))
            mg = torch.where(ok, G / torch.where(ok, cg, torch.ones_like(cg)), torch.zeros_like(G))
            ma = ma + d_hat * torch.eye(A.shape[0], dtype=A.dtype, device=A.device)
            mg = mg + d_hat * torch.eye(G.shape[0], dtype=G.dtype, device=G.device)
            ia = groups.setdefault(A.shape[0], [])
            ia.append(ma)
            ja = len(ia) - 1
            ig = groups.setdefault(G.shape[0], [])  # (the same list when both factors have the same size)
            ig.append(mg)
            blocks.append((e, (A.shape[0], ja), (G.shape[0], len(ig) - 1), ck))
        # (dh_spd_inverse: one Gauss-Jordan block per matrix, no library initialisation or workspace; it beats the
        # library's batched LU up to ~300 rows, scripts/gpu_kfac_timing.py)
        inv = _batched_inverses(groups)
        # pass 2: U = A_inv V G_inv / (c_k^2 npw)
        for e, (na, ja), (ng, jg), ck in blocks:
            din, dout, hb, npw, ko = e["in_dim"], e["out_dim"], e["has_bias"], e["rows_per_walker"], e["kernel_offset"]
            V = grads[ko : ko + din * dout].view(din, dout)
            if hb:
                bo = e["bias_offset"]
                V = torch.cat([V, grads[bo : bo + dout].view(1, dout)], dim=0)
            U = inv[na][ja] @ V @ inv[ng][jg] / (ck * ck * npw)
            if hb:
                out[bo : bo + dout] = U[din]
                U = U[:din]
            out[ko : ko + din * dout] = U.reshape(-1)
        return out

    def init(params, key, data):
        del key, data
        return KfacState(0, 0.0, torch.zeros(nfloats, dtype=torch.float32, device=params.device),
                         torch.zeros(4, 4, dtype=torch.float32, device=params.device))

    def step(state: CheckpointState, key):
        del key
        params, data, opt, width = state
        stats, grads = loss_grad_fn(params, data)
        grads = constants.pmean(grads)
        # kfac_jax order: curvature estimate (same batch) and inverses first, then the update
        new, x0 = normalised_stats(params, data)
        new, x0 = constants.pmean(new), constants.pmean(x0)
        opt = KfacState(opt.step, opt.weight * curvature_ema + 1.0, opt.stats * curvature_ema + new,
                        opt.dense0_xtx * curvature_ema + x0)
        pg = precondition(opt, grads)
        lr = optim_cfg.lr.schedule(opt.step)
        sq = float((pg * grads).sum()) * lr * lr
        coeff = min(1.0, math.sqrt(norm_constraint / sq)) if sq > 0 else 1.0
        params = params - (lr * coeff) * pg
        return CheckpointState(params, data, KfacState(opt.step + 1, opt.weight, opt.stats, opt.dense0_xtx), width), stats

    return init, step
