#!/usr/bin/env python
"""Dump golden vectors from the REAL reference (peterzjx/DeepHall on JAX/flax) for the Psiformer hot path.

Run this on any machine where the reference runs (`pip install -e <DeepHall checkout>`; CPU is enough):

    JAX_PLATFORMS=cpu python scripts/dump_reference_golden.py --out tests/golden/reference_psiformer.npz

It imports the unmodified `deephall` package and, for each case below, records

    params      the flax tree `model.init(key, x)` returns, flattened to {"a/b/c": array} (NO ordering assumptions)
    x           walkers (B, N, 2) after `burn` calls of the reference's own `mcmc_step`
    logpsi      `vmap(model.apply)(params, x)`                       complex64 (networks/psiformer.py:72-76)
    energy, kinetic, potential, angular_momentum_z, _z_square, _square   `vmap(hamiltonian.local_energy(...))`
                                                                    (hamiltonian.py:175-212)
    loss_energy, loss_variance, grad    `loss.make_loss_fn(...)(params, x)` under `constants.pmap` with one device
                                        (loss.py:47-110); grad flattened like params

`tests/test_reference_golden.py` loads the file when it exists (skipped otherwise) and checks the fp64 oracle on CPU and
the CUDA path on the GPU against it.  With that file committed, rows a1-a6 of SURVEY 8 stop being "parity unpinned".

This script cannot run in the build image (no jax / flax / kfac_jax there); nothing in the product imports it.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

CASES = {
    # tag: (nspins, flux, psiformer kwargs, orbital, interaction_strength, walkers, burn-in calls of mcmc_step)
    "c1": ((3, 0), 2, dict(num_layers=2), "full", 0.0, 64, 5),
    "c1s": ((3, 0), 2, dict(num_heads=2, heads_dim=16, num_layers=1, determinants=2), "full", 1.0, 32, 5),
    "c2": ((6, 0), 15, dict(), "full", 1.0, 32, 5),
    "c3": ((12, 0), 33, dict(), "full", 1.0, 16, 5),
    "spin": ((3, 2), 8, dict(num_heads=2, heads_dim=16, num_layers=1, determinants=2), "full", 1.0, 16, 5),
    "sparse": ((4, 0), 9, dict(num_heads=2, heads_dim=16, num_layers=1), "sparse", 1.0, 16, 5),
}


def flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        name = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, dict) or hasattr(v, "items"):
            out.update(flatten(v, name))
        else:
            out[name] = np.asarray(v)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_psiformer.npz"))
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--perturb", type=float, default=0.1, help="N(0, s^2) added to biases / scales / Jastrow so every term is exercised")
    args = ap.parse_args()

    import jax
    import jax.numpy as jnp

    import deephall
    from deephall import constants, hamiltonian, loss, mcmc
    from deephall.config import Network, OrbitalType, PsiformerNetwork, System
    from deephall.networks import make_network
    from deephall.train import init_guess

    out = {"meta_deephall_file": np.array(deephall.__file__), "meta_jax_version": np.array(jax.__version__)}
    for tag in args.cases.split(","):
        nspins, flux, pkw, orbital, kappa, B, burn = CASES[tag]
        system = System(flux=flux, nspins=nspins, interaction_strength=kappa)
        network = Network(orbital=OrbitalType(orbital), psiformer=PsiformerNetwork(**pkw))
        model = make_network(system, network)
        key = jax.random.PRNGKey(1234)
        key, k_data, k_par, k_pert = jax.random.split(key, 4)
        x = init_guess(k_data, B, sum(nspins))
        params = model.init(k_par, x[0])
        if args.perturb:  # kernels stay at their init; every other leaf moves away from 0 / 1
            leaves, treedef = jax.tree_util.tree_flatten_with_path(params)
            ks = jax.random.split(k_pert, len(leaves))
            new = []
            for (path, leaf), kk in zip(leaves, ks):
                is_kernel = "kernel" in jax.tree_util.keystr(path)
                new.append(leaf if is_kernel else leaf + args.perturb * jax.random.normal(kk, leaf.shape, leaf.dtype))
            params = jax.tree_util.tree_unflatten(treedef, new)
        batch_network = jax.vmap(model.apply, in_axes=(None, 0))
        step = constants.pmap(mcmc.make_mcmc_step(batch_network, B, steps=10))
        xs, ps = x[None], jax.tree.map(lambda a: a[None], params)
        for i in range(burn):
            key, sub = jax.random.split(key)
            xs, pmove = step(ps, xs, sub[None], jnp.asarray([0.3]))
        x = xs[0]
        logpsi = batch_network(params, x)
        e_l = jax.vmap(hamiltonian.local_energy(model.apply, system), in_axes=(None, 0))
        el, obs = e_l(params, x)
        stats, grads = constants.pmap(loss.make_loss_fn(model.apply, system))(ps, xs)
        rec = {"x": x, "logpsi": logpsi, "energy": el, "loss_energy": stats["energy"][0], "loss_variance": stats["variance"][0],
               "pmove": pmove[0]}
        rec.update({k: v for k, v in obs.items()})
        for k, v in rec.items():
            out[f"{tag}/{k}"] = np.asarray(v)
        for k, v in flatten(jax.tree.map(lambda a: a, params)).items():
            out[f"{tag}/params/{k}"] = v
        for k, v in flatten(jax.tree.map(lambda a: a[0], grads)).items():
            out[f"{tag}/grad/{k}"] = v
        out[f"{tag}/cfg"] = np.array([nspins[0], nspins[1], flux, pkw.get("determinants", 1), pkw.get("num_heads", 4),
                                      pkw.get("heads_dim", 64), pkw.get("num_layers", 2), int(orbital == "sparse")])
        out[f"{tag}/kappa"] = np.float64(kappa)
        print(f"{tag}: B={B} pmove={float(pmove[0]):.3f} E={complex(stats['energy'][0]):.6f} "
              f"params={sum(int(np.prod(v.shape)) for v in flatten(params).values())}", file=sys.stderr)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, file=sys.stderr)


if __name__ == "__main__":
    main()
