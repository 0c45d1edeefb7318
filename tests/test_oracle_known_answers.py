"""Pins the CPU oracle on every known answer the reference's own tests hold for this path
(SURVEY 8c): tests/hamiltonian_test.py:42-76, tests/cli_test.py:41-42, tests/train_test.py:46-48."""
import math

import pytest
import torch

from oracle import hamiltonian as OH
from oracle import laughlin as OL
from oracle import loss as OLoss
from oracle import mcmc as OM
from oracle import psiformer as OP


def sample(batch, nelec, seed=1898):
    return OM.init_guess(torch.Generator().manual_seed(seed), batch, nelec, torch.float64)


def test_free_electron():
    # hamiltonian_test.py:42-62: Y_1m Slater determinant, Q=0, r=1 -> KE = 3, L^2 = 0
    def log_psi(x):
        th, ph = x[..., 0], x[..., 1]
        orb = torch.stack([torch.sin(th) * torch.cos(ph), torch.cos(th), torch.sin(th) * torch.sin(ph)], -1).to(torch.complex128)
        s, l = torch.linalg.slogdet(orb)
        return l + torch.log(s)

    ke = OH.make_local_kinetic_energy(log_psi, 0, 1.0)
    for x in sample(2, 3):
        k, am = ke(x)
        assert abs(k - 3) < 1e-3
        assert abs(am["angular_momentum_square"]) < 1e-3


def make_lll(nelec, Q):
    def log_psi(x):
        u, v = OP.spinors(x)
        orb = torch.stack([u**m * v ** (2 * Q - m) for m in range(nelec)], -1)
        s, l = torch.linalg.slogdet(orb)
        return l + torch.log(s)

    return log_psi


@pytest.mark.parametrize("nelec,Q,L_square", [(1, 1, 2), (3, 1, 0), (9, 4, 0)])
def test_kinetic_and_angular_momentum(nelec, Q, L_square):
    # hamiltonian_test.py:65-76
    ke = OH.make_local_kinetic_energy(make_lll(nelec, Q), Q, math.sqrt(Q))
    for x in sample(2, nelec):
        k, am = ke(x)
        assert abs(k - nelec / 2) < 1e-3
        assert abs(am["angular_momentum_square"] - L_square) < 1e-3


def test_laughlin_energy():
    # cli_test.py:37-42: Laughlin N=3, flux 6, Coulomb -> energy=2.58..., L_square=0.0000
    flux, N, B = 6, 3, 512
    gen = torch.Generator().manual_seed(42)
    x = OM.init_guess(gen, B, N, torch.float64)

    def f(xx):
        return OL.logpsi(xx, flux)

    for _ in range(30):  # burn-in
        x, _ = OM.mcmc_step(f, x, OM.draw_randoms(gen, 10, B, N, torch.float64), 0.4)
    energies, l2s = [], []
    for _ in range(6):
        x, pmove = OM.mcmc_step(f, x, OM.draw_randoms(gen, 10, B, N, torch.float64), 0.4)
        res = OH.batch_local_energy(f, x, flux / 2, chunk=512)
        stats, _ = OLoss.loss_stats(res["energy"], {k: res[k] for k in ("angular_momentum_square",)})
        energies.append(stats["energy"].real.item())
        l2s.append(stats["angular_momentum_square"].item())
    e = sum(energies) / len(energies)
    assert abs(e - 2.5866) < 0.02, e  # prints "energy=2.58" in the reference
    assert abs(sum(l2s) / len(l2s)) < 5e-4
    assert 0.2 < pmove < 0.9


def test_param_counts_match_survey():
    # SURVEY 8(a1) closed-form parameter counts of the flax tree, + 1: blocks.py:99 creates the (inert) ee_anti leaf
    # whenever n_up > 0 -- its test is r_ees[0][1].shape[0] > 0 on an (n_up, n_dn) block
    expect = {
        ((3, 0), 2, 1): 796_692,
        ((6, 0), 15, 1): 841_410,
        ((12, 0), 33, 1): 1_001_778,
        ((10, 0), 21, 1): 905_146,
        ((16, 0), 45, 4): 2_305_282,
        ((16, 0), 45, 16): 6_844_930,
    }
    for (nspins, flux, k), n in expect.items():
        assert OP.num_params(OP.NetCfg(nspins=nspins, flux=flux, ndets=k)) == n


def test_psiformer_antisymmetry_and_lll_floor():
    cfg = OP.NetCfg(nspins=(6, 0), flux=15, num_heads=2, heads_dim=16)
    p = OP.init_params(cfg, 3, torch.float64, 0.1)
    x = sample(3, 6, seed=2)
    lp = OP.logpsi(p, x, cfg)
    perm = [1, 0, 2, 3, 4, 5]
    lp2 = OP.logpsi(p, x[:, perm], cfg)
    assert torch.allclose(lp.real, lp2.real, atol=1e-10)
    d = (lp.imag - lp2.imag).abs()
    assert torch.allclose(torch.minimum(d, 2 * math.pi - d), torch.full_like(d, math.pi), atol=1e-9)
    # zero orbital kernels + no Jastrow -> constant-coefficient LLL determinant -> KE = N/2 (train_test.py:46-48 floor)
    for k in list(p.keys()):
        if "DenseGeneral" in k and k.endswith("/kernel"):
            p[k] = torch.zeros_like(p[k])
    p["Jastrow_0/ee_par"] = torch.zeros(1, dtype=torch.float64)
    res = OH.batch_local_energy(lambda xx: OP.logpsi(p, xx, cfg), x, cfg.Q, interaction_strength=0.0)
    assert torch.allclose(res["kinetic"].real, torch.full((3,), 3.0, dtype=torch.float64), atol=1e-8)
    assert res["kinetic"].imag.abs().max() < 1e-8


@pytest.mark.parametrize("N,lz", [(4, 1.0), (4, -2.0), (5, 0.5)])
def test_laughlin_quasihole_is_an_lll_lz_eigenstate(N, lz):
    # networks/laughlin.py:38-41,73-83: N = 2 Q1 electrons, the orbital m = -lz left out of the 2 Q1 + 1.  The reference
    # holds no number for it; the state it describes is a lowest-Landau-level state (KE = N/2) and an L_z eigenstate
    # whose eigenvalue is the omitted orbital's -(-lz) = lz (a filled shell has L_z = 0).
    flux = N + 2 * (N - 1)
    x = sample(3, N, seed=5)
    res = OH.batch_local_energy(lambda xx: OL.logpsi(xx, flux, 1, lz), x, flux / 2, chunk=3)
    assert (res["kinetic"].real - N / 2).abs().max() < 1e-6
    assert (res["angular_momentum_z"].real - lz).abs().max() < 1e-6
    assert OL.laughlin_orbitals(x[0], flux, 1, lz).shape == (N, N)


@pytest.mark.parametrize("N,lz", [(3, 1.5), (4, 0.0), (4, -2.0), (5, 0.5)])
def test_laughlin_quasiparticle_is_an_lll_state_of_l_q1_plus_1(N, lz):
    # networks/laughlin.py:42-46,85-100: N = 2 Q1 + 2 electrons, the filled shell plus one LLL-projected orbital.
    # No number in the reference; what it describes: KE = N/2 (LLL), L_z = lz, L = Q1 + 1.
    flux = (N - 2) + 2 * (N - 1)
    Q1 = (N - 2) / 2
    x = sample(3, N, seed=5)
    res = OH.batch_local_energy(lambda xx: OL.logpsi(xx, flux, 1, lz), x, flux / 2, chunk=3)
    assert (res["kinetic"].real - N / 2).abs().max() < 1e-6
    assert (res["angular_momentum_z"].real - lz).abs().max() < 1e-6
    assert (res["angular_momentum_square"].real - (Q1 + 1) * (Q1 + 2)).abs().max() < 1e-6


def test_estimator_oracle_closed_forms():
    # netobs_bridge/observables: the reference holds no numbers for these; closed forms pin the restatement.
    import numpy as np

    from oracle import observables as OO

    rng = np.random.default_rng(0)
    B, N, bins = 40000, 6, 10
    data = np.stack([np.arccos(rng.uniform(-1, 1, (B, N))), rng.uniform(-np.pi, np.pi, (B, N))], -1)
    # independent uniform points: theta_12 has density sin/2, so the 1/sin-weighted histogram is flat at (N-1)/N
    g = OO.pair_correlation_increment(data, bins)
    assert np.abs(g - (N - 1) / N).max() < 0.02
    # density: counts sum to B N and follow sin(theta)/2
    d = OO.density_increment(data, bins)
    assert d.sum() == B * N
    edges = np.linspace(0, np.pi, bins + 1)
    expect = B * N * (np.cos(edges[:-1]) - np.cos(edges[1:])) / 2
    assert np.abs(d - expect).max() < 5 * np.sqrt(expect.max())
    # overlap of a state with itself (up to a constant factor) is 1; with a random other state it is < 1
    lp = rng.normal(size=500) + 1j * rng.uniform(-np.pi, np.pi, 500)
    ev = OO.overlap_evaluate(lp + (0.3 - 0.2j), lp)
    assert abs(OO.overlap_digest(ev["ratio"], ev["ratio_square"]) - 1) < 1e-12
    ev = OO.overlap_evaluate(lp + 0.5 * rng.normal(size=500), lp)
    assert OO.overlap_digest(ev["ratio"], ev["ratio_square"]) < 0.95


def test_one_rdm_oracle_on_the_filled_shell():
    # one_rdm.py:91-124 restated.  Closed form: for the filled lowest Landau level (N = 2Q + 1 free fermions in the
    # orbitals Y_{Q,Q,m}) every orbital is occupied once: the 1-RDM is the identity, trace N.  Sampling |psi|^2 is not
    # needed for an exact check: with r' uniform the estimator is unbiased walker by walker only on average, so the
    # mean over an equilibrated batch is compared statistically.
    import numpy as np

    from oracle import observables as OO

    N, flux, B = 3, 2, 3000
    gen = torch.Generator().manual_seed(0)
    f = make_lll(N, flux / 2)
    x = OM.init_guess(gen, B, N, torch.float64)
    for _ in range(30):
        x, _ = OM.mcmc_step(f, x, OM.draw_randoms(gen, 10, B, N, torch.float64), 0.7)
    rng = np.random.default_rng(1)
    rp = np.stack([np.arccos(rng.uniform(-1, 1, B)), rng.uniform(-np.pi, np.pi, B)], -1)
    xn = x.numpy()
    dp = OO.one_rdm_data_prime(xn, rp)
    assert dp.shape == (B, N, N, 2) and (dp[5, 1, 1] == rp[5]).all() and (dp[5, 1, 0] == xn[5, 0]).all()
    lp = f(x).numpy()
    lpp = f(torch.from_numpy(dp.reshape(B * N, N, 2))).numpy().reshape(B, N)
    rdm = OO.one_rdm_product(flux, xn, rp, lp, lpp).mean(0)
    assert abs(rdm - np.eye(flux + 1)).max() < 0.12
    assert abs(np.trace(rdm) - N) < 0.1
