"""TEST INFRASTRUCTURE ONLY (CPU oracle): KFAC curvature factors and update rule.

The reference's default optimizer is `kfac_jax.Optimizer` (kfac_jax==0.0.6, `optimizers/kfac.py:198-219`) with the
settings  l2_reg=0, norm_constraint=1e-3, curvature_ema=0.95, inverse_update_period=1, damping=1e-3, momentum=0,
estimation_mode="fisher_exact"  on the loss tag `register_normal_predictive_distribution(Re log psi)` (`loss.py:98`,
kfac_jax's default variance 0.5).  kfac_jax is not installed here and the reference holds no golden vectors for an
optimizer step, so this is a restatement of the PUBLISHED algorithm -- PARITY UNPINNED -- of:

* `RepeatedDenseBlock` (`optimizers/kfac.py:42-102`): every Dense / DenseGeneral is a two-Kronecker-factor block whose
  batch is (walkers x electrons): A = mean[x~ x~^T] (x~ = input with a trailing 1 when the layer has a bias),
  G = mean[dy dy^T], dy = tangent of the layer output for the loss tangent 1/sqrt(variance) per walker; the block's
  curvature is fixed_scale * A (x) G with fixed_scale = number of electrons that pass through the layer;
* LayerNorm scale / bias: scale-and-shift diagonal blocks, mean_b[(per-walker gradient)^2];
* Jastrow ee_par / ee_anti: no pattern -> generic tag -> naive diagonal (batch-summed gradient)^2 / batch;
* sparse orbitals (`blocks.py:52-62`): the 8-feature projections are repeated-dense blocks like the full ones (output
  [8][N][K]); `lll_weight` = `nn.DenseGeneral(2Q+1, axis=1)` on a COMPLEX tensor contracts axis 1 of a 4-d operand, the
  dimension numbers of none of `optimizers/kfac.py:148-195`'s example graphs (they contract the last axis; its
  `repeated_dense_complex_no_bias` is 3-d, last axis) -> generic tag for kernel and bias, like the Jastrow parameters;
* factors are exponential moving averages with weight de-biasing, the damped inverse of a Kronecker pair is the
  pi-adjusted factored Tikhonov form with average-trace norms, and the update is
  -lr * min(1, sqrt(norm_constraint / (lr^2 <P g, g>))) * P g.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import psiformer as OP

VARIANCE = 0.5  # kfac_jax.register_normal_predictive_distribution default


def dense_blocks(cfg: OP.NetCfg):
    """[(kernel name, bias name | None, electrons through the layer)] in parameter order."""
    p = "PsiformerLayers_0/"
    N = cfg.nelec
    out = [(p + "Dense_0/kernel", None, N)]
    for l in range(cfg.num_layers):
        a = f"{p}MultiHeadAttention_{l}/"
        for nm in ("query", "key", "value", "out"):
            out.append((a + nm + "/kernel", a + nm + "/bias", N))
        out.append((f"{p}Dense_{1 + 2 * l}/kernel", None, N))
        out.append((f"{p}Dense_{2 + 2 * l}/kernel", f"{p}Dense_{2 + 2 * l}/bias", N))
    o = "Orbitals_0/featured_orbitals/"
    idx = 0
    for n_alpha in cfg.nspins:
        if n_alpha:
            for _ in range(2):
                out.append((f"{o}DenseGeneral_{idx}/kernel", f"{o}DenseGeneral_{idx}/bias", n_alpha))
                idx += 1
    return out


def layer_io(params, x, cfg: OP.NetCfg):
    """Forward pass that records, per dense layer, its input rows and its (graph-attached) output; returns
    (log psi (B,), {kernel name: (input (B, n, in), output (B, n, out))}, {LayerNorm prefix: (xhat, y)})."""
    H, hd, D = cfg.num_heads, cfg.heads_dim, cfg.dim
    p = "PsiformerLayers_0/"
    io, ln = OrderedDict(), OrderedDict()

    def dense(name, xin, bias=None, kshape=None):
        k = params[name].reshape(kshape) if kshape else params[name]
        y = xin @ k
        if bias is not None:
            y = y + params[bias].reshape(-1)
        if y.requires_grad:
            y.retain_grad()
        io[name] = (xin, y)
        return y

    def lnorm(prefix, u):
        mu = u.mean(-1, keepdim=True)
        var = torch.clamp((u * u).mean(-1, keepdim=True) - mu * mu, min=0.0)
        xhat = (u - mu) * torch.rsqrt(var + 1e-5)
        y = xhat * params[prefix + "scale"] + params[prefix + "bias"]
        if y.requires_grad:
            y.retain_grad()
        ln[prefix] = (xhat, y)
        return y

    h = dense(p + "Dense_0/kernel", OP.input_feature(x, cfg))
    for l in range(cfg.num_layers):
        a = f"{p}MultiHeadAttention_{l}/"
        q = dense(a + "query/kernel", h, a + "query/bias", (D, D)).reshape(*h.shape[:-1], H, hd)
        k = dense(a + "key/kernel", h, a + "key/bias", (D, D)).reshape(*h.shape[:-1], H, hd)
        v = dense(a + "value/kernel", h, a + "value/bias", (D, D)).reshape(*h.shape[:-1], H, hd)
        w = torch.softmax(torch.einsum("...qhd,...khd->...hqk", q / math.sqrt(hd), k), dim=-1)
        att = torch.einsum("...hqk,...khd->...qhd", w, v).reshape(*h.shape[:-1], D)
        t1 = dense(a + "out/kernel", att, a + "out/bias", (D, D))
        h = lnorm(f"{p}LayerNorm_{2 * l}/", h + dense(f"{p}Dense_{1 + 2 * l}/kernel", t1))
        z = dense(f"{p}Dense_{2 + 2 * l}/kernel", h, f"{p}Dense_{2 + 2 * l}/bias")
        h = lnorm(f"{p}LayerNorm_{2 * l + 1}/", h + torch.tanh(z))
    o = "Orbitals_0/featured_orbitals/"
    N, L, K = cfg.nelec, cfg.norb, cfg.ndets
    F = L if cfg.orbital_type == "full" else 8  # blocks.py:47-56: sparse orbitals project to 8 features first
    cs, idx, start = [], 0, 0
    for n_alpha in cfg.nspins:
        if n_alpha:
            ha = h[..., start : start + n_alpha, :]
            re = dense(f"{o}DenseGeneral_{idx}/kernel", ha, f"{o}DenseGeneral_{idx}/bias", (D, F * N * K))
            im = dense(f"{o}DenseGeneral_{idx + 1}/kernel", ha, f"{o}DenseGeneral_{idx + 1}/bias", (D, F * N * K))
            cs.append(torch.complex(re, im).reshape(*ha.shape[:-1], F, N, K))
            idx += 2
        start += n_alpha
    c = torch.cat(cs, dim=-4)
    if cfg.orbital_type == "sparse":  # blocks.py:57,61-62: real (8, L) kernel and bias on the complex 8-feature tensor
        wl, bl = params["Orbitals_0/lll_weight/kernel"], params["Orbitals_0/lll_weight/bias"]
        c = torch.einsum("...nsjk,sl->...nljk", c, wl.to(c.dtype)) + bl.to(c.dtype)[:, None, None]
    env = OP.envelope(x, cfg)
    orb = torch.movedim((c * env[..., None, None]).sum(-3), -1, -3)
    jas = OP.jastrow(params, x, cfg)
    orb = torch.exp(jas / cfg.nelec)[..., None, None, None] * orb
    return OP.slogdet_tail(orb), io, ln


def curvature_stats(params_flat, x, cfg: OP.NetCfg):
    """The per-step curvature statistics kfac_jax would feed into its moving averages, for the walkers x (B, N, 2):
    {kernel name: (A, G)} for the dense blocks and {parameter name: diagonal} for the others."""
    B = x.shape[0]
    pf = params_flat.detach().clone().requires_grad_(True)
    params = OP.unflatten_params(pf, cfg)
    lp, io, ln = layer_io(params, x, cfg)
    lp.real.sum().backward()
    t2 = 1.0 / VARIANCE  # squared loss tangent
    dense = OrderedDict()
    for kname, bname, n_alpha in dense_blocks(cfg):
        xin, y = io[kname]
        xr = xin.detach().reshape(-1, xin.shape[-1])
        if bname is not None:
            xr = torch.cat([xr, torch.ones_like(xr[:, :1])], dim=1)
        g = y.grad.reshape(-1, y.shape[-1])
        rows = xr.shape[0]
        assert rows == B * n_alpha
        dense[kname] = (xr.T @ xr / rows, t2 * (g.T @ g) / rows)
    diag = OrderedDict()
    for prefix, (xhat, y) in ln.items():
        gy = y.grad
        diag[prefix + "scale"] = t2 * ((xhat.detach() * gy).sum(-2) ** 2).sum(0) / B
        diag[prefix + "bias"] = t2 * (gy.sum(-2) ** 2).sum(0) / B
    off = 0
    for name, shape in OP.param_shapes(cfg).items():
        n = int(torch.tensor(shape).prod())
        if name.startswith("Jastrow_0/") or name.startswith("Orbitals_0/lll_weight/"):
            diag[name] = t2 * pf.grad[off : off + n] ** 2 / B  # naive diagonal: (batch-summed gradient)^2 / batch
        off += n
    return dense, diag


def pi_adjusted_inverses(A, G, damping):
    """kfac_jax.utils.pi_adjusted_kronecker_inverse for two factors with average-trace norms:
    (A (x) G + damping I)^-1 ~ A_inv (x) G_inv."""
    ca, cg = torch.trace(A) / A.shape[0], torch.trace(G) / G.shape[0]
    c = ca * cg
    if not (c > 0):
        sd = math.sqrt(damping)
        return torch.eye(A.shape[0], dtype=A.dtype) / sd, torch.eye(G.shape[0], dtype=G.dtype) / sd
    d_hat = torch.sqrt(damping / c)
    ck = torch.sqrt(c)
    ai = torch.linalg.inv(A / ca + d_hat * torch.eye(A.shape[0], dtype=A.dtype)) / ck
    gi = torch.linalg.inv(G / cg + d_hat * torch.eye(G.shape[0], dtype=G.dtype)) / ck
    return ai, gi


class Kfac:
    """Stateful restatement of kfac_jax.Optimizer.step with the reference's settings (see the module docstring)."""

    def __init__(self, cfg: OP.NetCfg, lr_schedule, norm_constraint=1e-3, curvature_ema=0.95, damping=1e-3):
        self.cfg, self.lr, self.nc, self.ema, self.damping = cfg, lr_schedule, norm_constraint, curvature_ema, damping
        self.step_count, self.weight = 0, 0.0
        self.dense, self.diag = None, None

    def update_curvature(self, params_flat, x):
        dense, diag = curvature_stats(params_flat, x, self.cfg)
        if self.dense is None:
            self.dense = OrderedDict((k, (torch.zeros_like(a), torch.zeros_like(g))) for k, (a, g) in dense.items())
            self.diag = OrderedDict((k, torch.zeros_like(v)) for k, v in diag.items())
        self.weight = self.weight * self.ema + 1.0
        for k, (a, g) in dense.items():
            self.dense[k] = (self.dense[k][0] * self.ema + a, self.dense[k][1] * self.ema + g)
        for k, v in diag.items():
            self.diag[k] = self.diag[k] * self.ema + v

    def precondition(self, grad_flat):
        shapes = OP.param_shapes(self.cfg)
        offs, off = {}, 0
        for name, shape in shapes.items():
            n = 1
            for s_ in shape:
                n *= s_
            offs[name] = (off, n)
            off += n
        out = torch.zeros_like(grad_flat)
        for kname, bname, n_alpha in dense_blocks(self.cfg):
            A, G = (t / self.weight for t in self.dense[kname])
            o, n = offs[kname]
            V = grad_flat[o : o + n].reshape(-1, G.shape[0])
            if bname is not None:
                ob, nb = offs[bname]
                V = torch.cat([V, grad_flat[ob : ob + nb].reshape(1, -1)], dim=0)
            ai, gi = pi_adjusted_inverses(A, G, self.damping / n_alpha)
            U = ai @ V @ gi / n_alpha
            if bname is not None:
                out[ob : ob + nb] = U[-1]
                U = U[:-1]
            out[o : o + n] = U.reshape(-1)
        for name, d in self.diag.items():
            o, n = offs[name]
            out[o : o + n] = grad_flat[o : o + n] / (d / self.weight + self.damping)
        return out

    def step(self, params_flat, grad_flat, x):
        """kfac_jax order: curvature estimate and inverses first, then the preconditioned, norm-constrained update."""
        self.update_curvature(params_flat, x)
        pg = self.precondition(grad_flat)
        lr = self.lr(self.step_count)
        sq = (pg * grad_flat).sum() * lr * lr
        coeff = min(1.0, math.sqrt(self.nc / float(sq))) if float(sq) > 0 else 1.0
        self.step_count += 1
        return params_flat - lr * coeff * pg
