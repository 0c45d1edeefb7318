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
