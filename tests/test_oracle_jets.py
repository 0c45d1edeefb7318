"""The forward-Laplacian bookkeeping (oracle.jets, the spec the CUDA kernels implement) equals the
reference's gradient + Hessian formula (oracle.hamiltonian <- hamiltonian.py:96-170)."""
import pytest
import torch

from oracle import hamiltonian as OH
from oracle import jets as OJ
from oracle import mcmc as OM
from oracle import psiformer as OP

CASES = [
    dict(nspins=(3, 0), flux=2, num_heads=2, heads_dim=8, num_layers=2),
    dict(nspins=(5, 0), flux=11, ndets=2, num_heads=2, heads_dim=8),
    dict(nspins=(3, 2), flux=8, ndets=2, num_heads=2, heads_dim=8),  # spin-unpolarised (ee_anti, 2 orbital blocks)
    dict(nspins=(4, 0), flux=9, num_heads=2, heads_dim=8, orbital_type="sparse"),  # blocks.py:52-62
    dict(nspins=(2, 2), flux=5, ndets=2, num_heads=2, heads_dim=8, orbital_type="sparse"),
]


@pytest.mark.parametrize("kw", CASES)
def test_jets_equal_reference_formula(kw):
    cfg = OP.NetCfg(**kw)
    p = OP.init_params(cfg, 0, torch.float64, 0.1)
    x = OM.init_guess(torch.Generator().manual_seed(9), 3, cfg.nelec, torch.float64)
    x[0, 0, 0] = 0.01  # an electron next to the pole: the (theta, phi) formula loses digits, the jets do not
    ref = OH.batch_local_energy(lambda xx: OP.logpsi(p, xx, cfg), x, cfg.Q)
    out = OJ.local_energy(p, x, cfg)
    assert (out["logpsi"] - OP.logpsi(p, x, cfg)).abs().max() < 1e-11
    for k in ref:
        scale = ref[k].abs().clamp(min=1.0)
        assert ((ref[k] - out[k]).abs() / scale).max() < 1e-7, k


def test_jets_filled_lll():
    # nelec = 2Q+1 with constant orbital coefficients: KE = N/2, L^2 = 0 (hamiltonian_test.py:65-76 case (3,1,0))
    cfg = OP.NetCfg(nspins=(3, 0), flux=2, num_heads=2, heads_dim=8)
    p = OP.init_params(cfg, 1, torch.float64, 0.3)
    for k in list(p.keys()):
        if "DenseGeneral" in k and k.endswith("/kernel"):
            p[k] = torch.zeros_like(p[k])
    x = OM.init_guess(torch.Generator().manual_seed(1), 4, 3, torch.float64)
    out = OJ.local_energy(p, x, cfg, interaction_strength=0.0)
    # the Jastrow is still active, so only L^2 = 0 of the determinant part is not expected; switch it off:
    p["Jastrow_0/ee_par"] = torch.zeros(1, dtype=torch.float64)
    out = OJ.local_energy(p, x, cfg, interaction_strength=0.0)
    assert (out["kinetic"] - 1.5).abs().max() < 1e-9
    assert out["angular_momentum_square"].abs().max() < 1e-8


def test_kfac_factors_are_the_exact_fisher_block_for_one_row():
    """oracle/kfac.py (the restatement the GPU KFAC path is tested against): with one walker and one electron a dense
    layer sees a single (input, output-tangent) pair, so the Kronecker product A (x) G of its factors IS the Fisher
    block of the unit-variance-0.5 normal predictive distribution on Re log psi: (2 / B) sum_b vec(dW_b) vec(dW_b)^T
    with dW_b = [x_b; 1] g_b^T (optimizers/kfac.py:42-102, loss.py:98)."""
    from oracle import kfac as OK

    cfg = OP.NetCfg(nspins=(1, 0), flux=0, num_heads=2, heads_dim=4, num_layers=1)
    p = OP.init_params(cfg, 3, torch.float64, 0.2)
    pf = OP.flatten_params(p)
    x = OM.init_guess(torch.Generator().manual_seed(2), 1, 1, torch.float64)
    dense, diag = OK.curvature_stats(pf, x, cfg)
    q = pf.clone().requires_grad_(True)
    OP.logpsi(OP.unflatten_params(q, cfg), x, cfg).real.sum().backward()
    off, offs = 0, {}
    for name, shape in OP.param_shapes(cfg).items():
        n = 1
        for s_ in shape:
            n *= s_
        offs[name] = (off, n)
        off += n
    for kname, bname, _ in OK.dense_blocks(cfg):
        A, G = dense[kname]
        o, n = offs[kname]
        g = q.grad[o : o + n].reshape(-1, G.shape[0])
        if bname is not None:
            ob, nb = offs[bname]
            g = torch.cat([g, q.grad[ob : ob + nb].reshape(1, -1)], dim=0)
        fisher = 2.0 * torch.outer(g.reshape(-1), g.reshape(-1))  # variance 0.5 -> squared tangent 2
        assert (torch.kron(A, G) - fisher).abs().max() <= 1e-9 * max(1.0, fisher.abs().max().item()), kname
    for name, d in diag.items():  # diagonal blocks: squared gradients
        o, n = offs[name]
        assert (d - 2.0 * q.grad[o : o + n] ** 2).abs().max() < 1e-9, name


def test_kfac_preconditioner_inverts_the_damped_kronecker_product():
    """pi_adjusted_inverses: A_inv (x) G_inv equals (A (x) G + d I)^-1 up to the factored-Tikhonov cross terms, and
    exactly when one factor is a multiple of the identity."""
    from oracle import kfac as OK

    g = torch.Generator().manual_seed(0)
    a = torch.randn(5, 5, generator=g, dtype=torch.float64)
    A = a @ a.T / 5 + 0.1 * torch.eye(5, dtype=torch.float64)
    G = 0.7 * torch.eye(4, dtype=torch.float64)
    d = 1e-2
    ai, gi = OK.pi_adjusted_inverses(A, G, d)
    exact = torch.linalg.inv(torch.kron(A, G) + d * torch.eye(20, dtype=torch.float64))
    # with G = c I the factored form (A/ca + dh)(G/cg + dh) has the cross term dh * A/ca + dh^2 instead of d/(ca cg):
    # compare on the eigenbasis of A instead of entry by entry
    w = torch.linalg.eigvalsh(A)
    ca, cg = torch.trace(A) / 5, torch.tensor(0.7, dtype=torch.float64)
    dh = torch.sqrt(d / (ca * cg))
    approx_eigs = 1.0 / ((w / ca + dh) * (1.0 + dh) * ca * cg)
    assert torch.allclose(torch.linalg.eigvalsh(torch.kron(ai, gi))[::4].sort().values, approx_eigs.sort().values, rtol=1e-10)
    # and it is within the damping's order of the exact inverse
    assert (torch.kron(ai, gi) - exact).abs().max() / exact.abs().max() < 0.5
