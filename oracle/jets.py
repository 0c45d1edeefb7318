"""Forward-Laplacian ("jet") evaluation of the local energy -- torch-CPU prototype.

TEST INFRASTRUCTURE ONLY.  This is *our* bookkeeping (the reference differentiates the
network with jax.grad / jax.hessian, hamiltonian.py:105-114); it is the executable spec
of what the CUDA kernels in deephall_b200/csrc compute, and is itself checked against
``oracle.hamiltonian`` (the reference's formula) in tests/test_oracle_jets.py.

Every intermediate quantity y carries R = 2N + 8 "rows" (leading axis):

  row 0                value y
  rows 1 .. 2N         J_k = delta_k y      k = 2 i + t: electron i rotated about its own
                                            tangent axis t (t=0: theta_hat, t=1: phi_hat)
  row 2N+1             S = sum_k delta_k^2 y
  rows 2N+2 .. 2N+4    D_a = delta_a y      a = x, y, z: ALL electrons rotated about e_a
  rows 2N+5 .. 2N+7    T_a = delta_a^2 y

where delta is the derivative along the rotation flow r(t) = exp(t [n]x) r,
(u,v)(t) = exp(i t (n.sigma)^T / 2) (u,v)  (SURVEY Appendix B, F12).  No cot(theta) or
1/sin^2(theta) appears anywhere.  With delta = i L:

  kinetic = -(S + sum_k J_k^2) / (2 r^2)                 (of log psi)
  L^2     = -sum_a (T_a + D_a^2),   L_z = Im D_z,   L_z^2 = -Re (T_z + D_z^2)
"""
from __future__ import annotations

import math

import torch

from .psiformer import NetCfg, norm_factor, spinors


class Rows:
    def __init__(self, N: int):
        self.N = N
        self.R = 2 * N + 8
        self.J = slice(1, 2 * N + 1)
        self.S = 2 * N + 1
        self.D = slice(2 * N + 2, 2 * N + 5)
        self.T = slice(2 * N + 5, 2 * N + 8)


# ----------------------------------------------------------------------------- jet algebra
def jet_mul(p, q, rw: Rows):
    """Product rule for two jets (row axis 0, trailing shapes broadcast)."""
    out_shape = torch.broadcast_shapes(p.shape, q.shape)
    out = torch.empty(out_shape, dtype=torch.result_type(p, q))
    out[0] = p[0] * q[0]
    out[rw.J] = p[rw.J] * q[0] + p[0] * q[rw.J]
    out[rw.S] = p[rw.S] * q[0] + p[0] * q[rw.S] + 2 * (p[rw.J] * q[rw.J]).sum(0)
    out[rw.D] = p[rw.D] * q[0] + p[0] * q[rw.D]
    out[rw.T] = p[rw.T] * q[0] + p[0] * q[rw.T] + 2 * p[rw.D] * q[rw.D]
    return out


def jet_fn(x, f0, f1, f2, rw: Rows):
    """Elementwise y = f(x): f0, f1, f2 are f, f', f'' evaluated at the value row."""
    out = torch.empty_like(x)
    out[0] = f0
    out[rw.J] = f1 * x[rw.J]
    out[rw.S] = f1 * x[rw.S] + f2 * (x[rw.J] ** 2).sum(0)
    out[rw.D] = f1 * x[rw.D]
    out[rw.T] = f1 * x[rw.T] + f2 * x[rw.D] ** 2
    return out


def jet_linear(x, W, b=None):
    """y = x @ W (+ b on the value row only)."""
    y = x @ W
    if b is not None:
        y[0] = y[0] + b
    return y


def jet_const(c, rw: Rows):
    out = torch.zeros((rw.R,) + tuple(c.shape), dtype=c.dtype)
    out[0] = c
    return out


# ----------------------------------------------------------------------------- seeds
def frames(x):
    th, ph = x[..., 0], x[..., 1]
    st, ct, sp, cp = torch.sin(th), torch.cos(th), torch.sin(ph), torch.cos(ph)
    rhat = torch.stack([st * cp, st * sp, ct], -1)
    that = torch.stack([ct * cp, ct * sp, -st], -1)
    phat = torch.stack([-sp, cp, torch.zeros_like(sp)], -1)
    return rhat, that, phat


def seed_rhat(x, rw: Rows):
    """Jet of r_hat: (R, B, N, 3)."""
    B, N = x.shape[0], x.shape[1]
    rhat, that, phat = frames(x)
    out = torch.zeros((rw.R, B, N, 3), dtype=x.dtype)
    out[0] = rhat
    for i in range(N):
        out[1 + 2 * i, :, i] = -phat[:, i]  # theta_hat x r_hat
        out[2 + 2 * i, :, i] = that[:, i]  # phi_hat x r_hat
    out[rw.S] = -2 * rhat
    eye = torch.eye(3, dtype=x.dtype)
    for a in range(3):
        ea = eye[a].expand_as(rhat)
        out[2 * N + 2 + a] = torch.linalg.cross(ea, rhat)
        out[2 * N + 5 + a] = ea * rhat[..., a : a + 1] - rhat
    return out


def _spinor_gen(n, u, v):
    """(i/2) (n.sigma)^T (u, v)."""
    nx, ny, nz = n[..., 0], n[..., 1], n[..., 2]
    du = 0.5j * (nz * u + (nx + 1j * ny) * v)
    dv = 0.5j * ((nx - 1j * ny) * u - nz * v)
    return du, dv


def seed_envelope(x, cfg: NetCfg, rw: Rows):
    """Jet of env[i, m] = sqrt(C(2Q,Q-m)) u_i^(Q+m) v_i^(Q-m): (R, B, N, L) complex."""
    B, N = x.shape[0], x.shape[1]
    twoQ = int(cfg.flux)
    u, v = spinors(x)
    cdt = u.dtype
    a = torch.arange(0, twoQ + 1)
    b = twoQ - a
    nf = norm_factor(cfg, x.dtype).to(cdt)

    def pw(z, e):  # z^(max(e,0)), shape (B,N,L)
        return z[..., None] ** torch.clamp(e, min=0)

    e0 = nf * pw(u, a) * pw(v, b)
    eu = nf * a * pw(u, a - 1) * pw(v, b)  # d e / d u
    ev = nf * b * pw(u, a) * pw(v, b - 1)
    euu = nf * a * (a - 1) * pw(u, a - 2) * pw(v, b)
    euv = nf * a * b * pw(u, a - 1) * pw(v, b - 1)
    evv = nf * b * (b - 1) * pw(u, a) * pw(v, b - 2)

    def first(du, dv):
        return eu * du[..., None] + ev * dv[..., None]

    def second(du, dv):  # includes u'' = -u/4, v'' = -v/4
        return (
            euu * (du * du)[..., None]
            + 2 * euv * (du * dv)[..., None]
            + evv * (dv * dv)[..., None]
            - 0.25 * (eu * u[..., None] + ev * v[..., None])
        )

    out = torch.zeros((rw.R, B, N, twoQ + 1), dtype=cdt)
    out[0] = e0
    _, that, phat = frames(x)
    S = torch.zeros_like(e0)
    for t, axis in enumerate((that, phat)):
        du, dv = _spinor_gen(axis, u, v)
        f1, f2 = first(du, dv), second(du, dv)
        for i in range(N):
            out[1 + 2 * i + t, :, i] = f1[:, i]
        S = S + f2
    out[rw.S] = S
    eye = torch.eye(3, dtype=x.dtype)
    for a_ in range(3):
        du, dv = _spinor_gen(eye[a_].expand(B, N, 3), u, v)
        out[2 * N + 2 + a_] = first(du, dv)
        out[2 * N + 5 + a_] = second(du, dv)
    return out


# ----------------------------------------------------------------------------- network
def jet_layer_norm(x, scale, bias, rw: Rows, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    c = x - mu  # linear: every row centred by its own mean
    var = jet_mul(c, c, rw).mean(-1, keepdim=True)
    v0 = var[0] + eps
    rho = jet_fn(var, v0**-0.5, -0.5 * v0**-1.5, 0.75 * v0**-2.5, rw)
    y = jet_mul(c, rho, rw) * scale
    y[0] = y[0] + bias
    return y


def jet_tanh(x, rw: Rows):
    t = torch.tanh(x[0])
    return jet_fn(x, t, 1 - t * t, -2 * t * (1 - t * t), rw)


def jet_attention(params, prefix, h, cfg: NetCfg, rw: Rows):
    """h: (R, B, N, D) -> MHA output (R, B, N, D)."""
    H, hd, D = cfg.num_heads, cfg.heads_dim, cfg.dim
    R, B, N, _ = h.shape

    def proj(nm):
        y = jet_linear(h, params[prefix + nm + "/kernel"].reshape(D, D), params[prefix + nm + "/bias"].reshape(D))
        return y.reshape(R, B, N, H, hd)

    q, k, v = proj("query") / math.sqrt(hd), proj("key"), proj("value")
    # scores s[b,h,i,j]
    qi = q.permute(0, 1, 3, 2, 4)[:, :, :, :, None, :]  # R,B,H,N,1,hd
    kj = k.permute(0, 1, 3, 2, 4)[:, :, :, None, :, :]  # R,B,H,1,N,hd
    s = jet_mul(qi, kj, rw).sum(-1)  # R,B,H,N,N
    m = s[0].max(-1, keepdim=True).values
    e0 = torch.exp(s[0] - m)
    ex = jet_fn(s, e0, e0, e0, rw)  # exp(s - m) has the same derivatives pattern
    den = ex.sum(-1, keepdim=True)
    inv = jet_fn(den, 1 / den[0], -1 / den[0] ** 2, 2 / den[0] ** 3, rw)
    p = jet_mul(ex, inv, rw)  # R,B,H,N,N
    vj = v.permute(0, 1, 3, 2, 4)[:, :, :, None, :, :]  # R,B,H,1,N,hd
    o = jet_mul(p[..., None], vj, rw).sum(-2)  # R,B,H,N,hd
    o = o.permute(0, 1, 3, 2, 4).reshape(R, B, N, D)
    return jet_linear(o, params[prefix + "out/kernel"].reshape(D, D), params[prefix + "out/bias"])


def jet_psiformer_layers(params, x, cfg: NetCfg, rw: Rows):
    B, N = x.shape[0], x.shape[1]
    rj = seed_rhat(x, rw)  # R,B,N,3  (x,y,z)
    feat = torch.zeros((rw.R, B, N, 4), dtype=x.dtype)
    feat[..., 0] = rj[..., 2]  # cos(theta)
    feat[..., 1] = rj[..., 0]
    feat[..., 2] = rj[..., 1]
    feat[0, ..., 3] = torch.tensor([1.0] * cfg.nspins[0] + [-1.0] * cfg.nspins[1], dtype=x.dtype)
    p = "PsiformerLayers_0/"
    h = jet_linear(feat, params[p + "Dense_0/kernel"])
    for l in range(cfg.num_layers):
        attn = jet_attention(params, f"{p}MultiHeadAttention_{l}/", h, cfg, rw)
        h = h + jet_linear(attn, params[f"{p}Dense_{1 + 2 * l}/kernel"])
        h = jet_layer_norm(h, params[f"{p}LayerNorm_{2 * l}/scale"], params[f"{p}LayerNorm_{2 * l}/bias"], rw)
        h = h + jet_tanh(jet_linear(h, params[f"{p}Dense_{2 + 2 * l}/kernel"], params[f"{p}Dense_{2 + 2 * l}/bias"]), rw)
        h = jet_layer_norm(h, params[f"{p}LayerNorm_{2 * l + 1}/scale"], params[f"{p}LayerNorm_{2 * l + 1}/bias"], rw)
    return h, rj


def jet_jastrow(params, rj, cfg: NetCfg, rw: Rows):
    """Jet of the Jastrow factor, (R, B), from the r_hat jet.  Polarised pairs only use
    ee_par; anti-parallel pairs use ee_anti (blocks.py:91-105)."""
    R, B, N, _ = rj.shape
    n_up = cfg.nspins[0]
    tot = torch.zeros((R, B), dtype=rj.dtype)
    for i in range(N):
        for j in range(i + 1, N):
            par = (i < n_up) == (j < n_up)
            a = params["Jastrow_0/ee_par"] if par else params["Jastrow_0/ee_anti"]
            coef = 0.25 if par else 0.5
            c = jet_mul(rj[:, :, i], rj[:, :, j], rw).sum(-1)  # cos12 jet (R,B)
            w = 2 - 2 * c[0]
            r = jet_fn(-2 * c, torch.sqrt(w), 0.5 * w**-0.5, -0.25 * w**-1.5, rw)
            r[0] = torch.sqrt(w)
            d = a + r[0]
            tot = tot + jet_fn(r, -coef * a**2 / d, coef * a**2 / d**2, -2 * coef * a**2 / d**3, rw)
    return tot


def jet_logdet(M, rw: Rows):
    """M: (R, B, K, N, N) complex -> jets of log det, (R, B, K) complex."""
    Minv = torch.linalg.inv(M[0])
    X = Minv @ M  # broadcast over rows
    tr = torch.diagonal(X, dim1=-1, dim2=-2).sum(-1)
    trsq = (X * X.transpose(-1, -2)).sum((-1, -2))
    out = torch.empty(M.shape[:3], dtype=M.dtype)
    sign, logabs = torch.linalg.slogdet(M[0])
    out[0] = logabs + torch.log(sign)
    out[rw.J] = tr[rw.J]
    out[rw.S] = tr[rw.S] - trsq[rw.J].sum(0)
    out[rw.D] = tr[rw.D]
    out[rw.T] = tr[rw.T] - trsq[rw.D]
    return out


def jet_logsumexp(ld, rw: Rows):
    """ld: (R, B, K) complex -> (R, B) complex jets of log sum_k exp(ld_k)."""
    mx = ld[0].real.max(-1, keepdim=True).values
    e0 = torch.exp(ld[0] - mx)
    ex = jet_fn(ld, e0, e0, e0, rw).sum(-1)
    out = jet_fn(ex, torch.log(ex[0]) + mx[..., 0], 1 / ex[0], -1 / ex[0] ** 2, rw)
    return out


def jet_logpsi(params, x, cfg: NetCfg):
    """x: (B, N, 2) -> jets of log psi, (R, B) complex."""
    N, L, K = cfg.nelec, cfg.norb, cfg.ndets
    rw = Rows(N)
    B = x.shape[0]
    h, rj = jet_psiformer_layers(params, x, cfg, rw)  # R,B,N,D
    o = "Orbitals_0/featured_orbitals/"
    D = cfg.dim
    cs, idx, start = [], 0, 0
    for n_alpha in cfg.nspins:
        if n_alpha:
            ha = h[:, :, start : start + n_alpha]
            re = jet_linear(ha, params[f"{o}DenseGeneral_{idx}/kernel"].reshape(D, -1), params[f"{o}DenseGeneral_{idx}/bias"].reshape(-1))
            im = jet_linear(ha, params[f"{o}DenseGeneral_{idx + 1}/kernel"].reshape(D, -1), params[f"{o}DenseGeneral_{idx + 1}/bias"].reshape(-1))
            cs.append(torch.complex(re, im))
            idx += 2
        start += n_alpha
    if cfg.orbital_type == "sparse":  # blocks.py:61-62: a linear map of the 8 features (bias on the value row only)
        c8 = torch.cat(cs, dim=2).reshape(rw.R, B, N, 8, N, K)
        w, b = params["Orbitals_0/lll_weight/kernel"], params["Orbitals_0/lll_weight/bias"]
        c = torch.einsum("rbnsjk,sl->rbnljk", c8, w.to(c8.dtype))
        c[0] = c[0] + b.to(c8.dtype)[:, None, None]
    else:
        c = torch.cat(cs, dim=2).reshape(rw.R, B, N, L, N, K)
    env = seed_envelope(x, cfg, rw)  # R,B,N,L
    orb = jet_mul(c, env[..., None, None], rw).sum(3)  # R,B,N(i),N(j),K
    M = orb.permute(0, 1, 4, 2, 3)  # R,B,K,N,N
    ld = jet_logdet(M, rw)
    lp = jet_logsumexp(ld, rw)
    jas = jet_jastrow(params, rj, cfg, rw) if N > 1 else torch.zeros_like(lp)
    return lp + jas


def observables_from_jets(lp, x, cfg: NetCfg, radius=None, interaction_strength=1.0, interaction_type="coulomb"):
    """Assemble (E_L, observables) from the log psi jets -- hamiltonian.py:193-210 outputs."""
    from .hamiltonian import potential

    N = cfg.nelec
    rw = Rows(N)
    Q = cfg.Q
    r = radius or math.sqrt(Q)
    kinetic = -(lp[rw.S] + (lp[rw.J] ** 2).sum(0)) / (2 * r * r)
    D, T = lp[rw.D], lp[rw.T]
    pot = potential(x, Q, r, interaction_type) * interaction_strength
    return {
        "energy": kinetic + pot,
        "kinetic": kinetic,
        "potential": pot,
        "angular_momentum_z": D[2].imag,
        "angular_momentum_z_square": -(T[2] + D[2] ** 2).real,
        "angular_momentum_square": -(T + D**2).sum(0).real,
        "logpsi": lp[0],
    }


def local_energy(params, x, cfg: NetCfg, **kw):
    return observables_from_jets(jet_logpsi(params, x, cfg), x, cfg, **kw)
