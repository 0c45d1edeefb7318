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


def _batched_inverses(groups: dict) -> dict:
    """{n: [n x n SPD matrices]} -> {n: (count, n, n) inverses}, by the library's own kernel (dh_spd_inverse).

    All sizes up to 1024 go out as ONE batch: a smaller matrix is embedded as diag(A, I) in the largest size of the
    batch (its inverse is diag(A^-1, I)); dh_spd_inverse runs one 8-CTA cluster per matrix with the matrix in
    distributed shared memory, so the batch takes the time of its largest member (~0.2 ms at 410 rows).
    With more than one rank the matrices are dealt out round-robin -- every rank inverts 1/world of them -- and the
    inverses are exchanged with one all-gather (the statistics they come from are already identical on all ranks).
    Factors with more than 1024 rows (orbital projections of multi-determinant c5-sized networks) are outside the
    kernel's range and go to torch.linalg.inv."""
    small = {n: ms for n, ms in groups.items() if n <= 1024}
    out = {n: torch.linalg.inv(torch.stack(ms)) for n, ms in groups.items() if n > 1024}
    if not small:
        return out
    nmax = max(small)
    dev = next(iter(small.values()))[0].device
    world, rank = constants.world_size(), constants.rank()
    total = sum(len(ms) for ms in small.values())
    padded = (total + world - 1) // world * world
    batch = torch.zeros((padded, nmax, nmax), dtype=torch.float32, device=dev)
    diag = torch.arange(nmax, device=dev)
    k = 0
    where = {}
    for n, ms in small.items():
        where[n] = (k, len(ms))
        batch[k : k + len(ms), :n, :n] = torch.stack(ms)
        if n < nmax:
            batch[k : k + len(ms), diag[n:], diag[n:]] = 1.0
        k += len(ms)
    if padded > total:
        batch[total:, diag, diag] = 1.0
    if world == 1:
        inv = _native.spd_inverse(batch, inplace=True)
    else:
        # matrix m belongs to rank m % world: a strided view keeps the all-gather layout trivial
        per = padded // world
        mine = batch.view(per, world, nmax, nmax)[:, rank].contiguous()
        _native.spd_inverse(mine, inplace=True)
        gathered = torch.empty((world, per, nmax, nmax), dtype=torch.float32, device=dev)
        torch.distributed.all_gather_into_tensor(gathered, mine)
        inv = gathered.permute(1, 0, 2, 3).reshape(padded, nmax, nmax)
    for n, (k0, cnt) in where.items():
        out[n] = inv[k0 : k0 + cnt, :n, :n]
    return out


_SIDE_STREAMS: dict = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


def _inverse_batch(batch: torch.Tensor) -> torch.Tensor:
    """In-place inverses of a (count, n, n) batch of SPD matrices by dh_spd_inverse; with more than one rank the matrices are
    dealt out round-robin -- every rank inverts 1/world of them -- and exchanged with one all-gather (the statistics
    they come from are identical on all ranks)."""
    world, rank = constants.world_size(), constants.rank()
    count, n = batch.shape[0], batch.shape[1]
    if count == 0:
        return batch
    if world == 1:
        return _native.spd_inverse(batch, inplace=True)
    per = (count + world - 1) // world
    mine = torch.eye(n, dtype=torch.float32, device=batch.device).repeat(per, 1, 1)
    own = batch[rank::world]
    mine[: own.shape[0]] = own
    _native.spd_inverse(mine, inplace=True)
    gathered = torch.empty((world, per, n, n), dtype=torch.float32, device=batch.device)
    torch.distributed.all_gather_into_tensor(gathered, mine)
    return gathered.permute(1, 0, 2, 3).reshape(per * world, n, n)[:count].contiguous()


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

    scale_cache: dict = {}

    def stat_scales(B: int, device):
        """Per-entry normalisation of the factor vector for a shard of B walkers (one multiply per step instead of one
        small kernel per curvature block), and the slices of the naive-diagonal entries, which are squared first."""
        if B not in scale_cache:
            sc = torch.zeros(nfloats, dtype=torch.float32)
            naive = []
            for e in layout:
                rows = float(B * e["rows_per_walker"])
                if e["kind"] == 0:
                    for key, n in (("xtx_offset", e["in_dim"] ** 2), ("xsum_offset", e["in_dim"])):
                        if e[key] >= 0:
                            sc[e[key] : e[key] + n] = 1.0 / rows
                    sc[e["gtg_offset"] : e["gtg_offset"] + e["out_dim"] ** 2] = t2 / rows
                else:  # kind 1: per-walker squares; kind 2 (naive diagonal): (batch-summed gradient)^2 / batch
                    sc[e["diag_offset"] : e["diag_offset"] + e["size"]] = t2 / B
                    if e["kind"] == 2:
                        naive.append((e["diag_offset"], e["size"]))
            scale_cache[B] = (sc.to(device), naive)
        return scale_cache[B]

    def normalised_stats(params, data):
        """This step's statistics (local shard), each already divided by its batch size."""
        # (the step's previous plan op is the gradient's VJP on these params and walkers: its forward is reused)
        raw = net.plan(system).kfac_factors(params, data.contiguous(), reuse_forward=True)
        sc, naive = stat_scales(data.shape[0], raw.device)
        for o, n in naive:
            raw[o : o + n] = raw[o : o + n] ** 2
        feat = _features(data, n_up).reshape(-1, 4)
        return raw * sc, feat.T @ feat / feat.shape[0]

    def precondition(state: KfacState, grads: torch.Tensor) -> torch.Tensor:
        """The update direction.  Library path (dh_kfac_damped_factors -> dh_spd_inverse -> dh_kfac_update): the damped
        factors of all blocks are assembled by one kernel, inverted in two batches (<= 288 rows / larger) and applied by the
        library's fp32 contraction; `precondition_tensor_ops` is the same rule as small tensor ops, for plans with a
        factor beyond dh_spd_inverse's 1024 rows."""
        pl = net.plan(system)
        if pl.kfac_update_shape() is None:
            return precondition_tensor_ops(state, grads)
        coef, ms, ml = pl.kfac_damped_factors(state.stats, state.dense0_xtx, state.weight, damping)
        # the two batches are inverted side by side (27 x 4 + 2 x 8 CTAs at c3: both fit the machine at once)
        cur = torch.cuda.current_stream()
        side = _side_stream(grads.device)
        side.wait_stream(cur)
        ml.record_stream(side)
        with torch.cuda.stream(side):
            inv_l = _inverse_batch(ml)
        inv_s = _inverse_batch(ms)
        cur.wait_stream(side)
        inv_l.record_stream(cur)
        return pl.kfac_update(inv_s, inv_l, coef, state.stats, state.weight, damping, grads.contiguous())

    def precondition_tensor_ops(state: KfacState, grads: torch.Tensor) -> torch.Tensor:
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
            # A_inv V G_inv with the library's fp32 contraction (dh_gemm), not a vendor GEMM
            U = _native.gemm(_native.gemm(inv[na][ja].contiguous(), V.contiguous()), inv[ng][jg].contiguous()) / (ck * ck * npw)
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
        grads_red = constants.pmean_async(grads)  # NCCL runs it while the curvature pass below computes
        # kfac_jax order: curvature estimate (same batch) and inverses first, then the update
        new, x0 = normalised_stats(params, data)
        new, x0 = constants.pmean(new), constants.pmean(x0)
        grads = grads_red()
        opt = KfacState(opt.step, opt.weight * curvature_ema + 1.0, opt.stats * curvature_ema + new,
                        opt.dense0_xtx * curvature_ema + x0)
        pg = precondition(opt, grads)
        lr = optim_cfg.lr.schedule(opt.step)
        # norm constraint on the device (no host synchronisation): coeff = min(1, sqrt(c / (lr^2 <pg, g>)))
        sq = (pg * grads).sum() * (lr * lr)
        coeff = torch.where(sq > 0, torch.clamp(torch.sqrt(norm_constraint / sq.clamp(min=1e-38)), max=1.0), torch.ones_like(sq))
        params = params - (lr * coeff) * pg
        return CheckpointState(params, data, KfacState(opt.step + 1, opt.weight, opt.stats, opt.dense0_xtx), width), stats

    step.precondition, step.precondition_tensor_ops = precondition, precondition_tensor_ops  # (tests compare the two routes)
    return init, step
