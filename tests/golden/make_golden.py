"""Generates tests/golden/psiformer_small.npz.

The reference (JAX/flax) cannot be imported in this image, so the vectors come from the fp64
oracle running the REFERENCE ALGORITHM (oracle.hamiltonian: complex gradient + full Hessian in
(theta, phi), hamiltonian.py:105-170) -- not from the forward-Laplacian code path the CUDA
kernels mirror.  Parameters are stored explicitly (fp32 values), so the fixture does not depend
on torch's RNG.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import hamiltonian as OH  # noqa: E402
from oracle import mcmc as OM  # noqa: E402
from oracle import psiformer as OP  # noqa: E402

CASES = {
    "a": dict(nspins=(3, 0), flux=2, ndets=1, num_heads=2, heads_dim=16, num_layers=1),
    "b": dict(nspins=(6, 0), flux=15, ndets=2, num_heads=2, heads_dim=32, num_layers=2),
}


def main():
    out = {}
    for tag, kw in CASES.items():
        cfg = OP.NetCfg(**kw)
        p32 = OP.flatten_params(OP.init_params(cfg, seed=11, dtype=torch.float64, perturb=0.1)).float()
        params = OP.unflatten_params(p32.double(), cfg)
        B = 12
        x = OM.init_guess(torch.Generator().manual_seed(5), B, cfg.nelec, torch.float32)
        # a few MH moves so that determinants are not pathologically small
        rnd = OM.draw_randoms(torch.Generator().manual_seed(6), 10, B, cfg.nelec, torch.float64)
        x64, _ = OM.mcmc_step(lambda xx: OP.logpsi(params, xx, cfg), x.double(), rnd, 0.3)
        x = x64.float()
        x64 = x.double()
        res = OH.batch_local_energy(lambda xx: OP.logpsi(params, xx, cfg), x64, cfg.Q, interaction_strength=1.0)
        lp = OP.logpsi(params, x64, cfg)
        cot = torch.randn(B, 2, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
        p = p32.double().clone().requires_grad_(True)
        lpp = OP.logpsi(OP.unflatten_params(p, cfg), x64, cfg)
        (grad,) = torch.autograd.grad((lpp.real * cot[:, 0] + lpp.imag * cot[:, 1]).sum(), p)
        out[f"{tag}_cfg"] = np.array([kw["nspins"][0], kw["nspins"][1], kw["flux"], kw["ndets"], kw["num_heads"], kw["heads_dim"], kw["num_layers"]])
        out[f"{tag}_params"] = p32.numpy()
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_logpsi"] = lp.numpy()
        for k, v in res.items():
            out[f"{tag}_{k}"] = v.numpy()
        out[f"{tag}_cot"] = cot.numpy()
        out[f"{tag}_grad"] = grad.float().numpy()
        print(tag, "E_L", res["energy"][:3], "params", p32.numel())
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "psiformer_small.npz"), **out)


if __name__ == "__main__":
    main()
