"""Accuracy of the CUDA path vs the fp64 oracle for the three dense-contraction implementations
(per-walker distributions; SURVEY 8d: median / p90 / max).  Writes a markdown table to stdout."""
import math
import os
import sys

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402
from oracle import jets as OJ  # noqa: E402
from oracle import psiformer as OP  # noqa: E402

CASES = [("c1", dict(nspins=(3, 0), flux=2), 256, 0.0), ("c2", dict(nspins=(6, 0), flux=15), 256, 1.0),
         ("c3", dict(nspins=(12, 0), flux=33), 256, 1.0), ("c4", dict(nspins=(10, 0), flux=21), 128, 1.0),
         ("c5k4", dict(nspins=(16, 0), flux=45, ndets=4), 48, 1.0)]
# dh_config.contraction (round 2: no environment switches in the library)
MODES = [("fp32 FMA", "fp32"), ("fp16 pieces (default: tcgen05 contractions, tensor-core attention, fused epilogues)", "f16"),
         ("tf32 pieces (range-guard fallback)", "tf32")]


def q(v):
    v = v.flatten().sort().values
    return v[len(v) // 2].item(), v[int(len(v) * 0.9)].item(), v[-1].item()


print("| config | contraction | E_L rel. err median / p90 / max | kinetic rel. err median / p90 / max | Re log psi abs err max | phase err max |")
print("|---|---|---|---|---|---|")
for name, kw, B, kappa in CASES:
    cfg = OP.NetCfg(**kw)
    p64 = OP.init_params(cfg, 0, torch.float64, 0.1)
    flat32 = OP.flatten_params(p64).float()
    p64 = OP.unflatten_params(flat32.double(), cfg)
    ref = None
    for label, mode in MODES:
        plan = nat.Plan(nspins=cfg.nspins, flux=cfg.flux, ndets=cfg.ndets, interaction_strength=kappa, contraction=mode)
        flat = flat32.cuda()
        x = plan.init_walkers(B, seed=11)
        plan.mcmc_sweep(flat, x, 30, 0.2, seed=3)
        x = x.clone()
        if ref is None:
            xref = x.clone()
            ref = OJ.local_energy(p64, xref.double().cpu(), cfg, interaction_strength=kappa)
        out = plan.local_energy(flat, xref)
        e = (out["energy"].cpu().to(torch.complex128) - ref["energy"]).abs() / ref["energy"].abs()
        k = (out["kinetic"].cpu().to(torch.complex128) - ref["kinetic"]).abs() / ref["kinetic"].abs()
        lp = out["logpsi"].cpu().to(torch.complex128)
        dre = (lp.real - ref["logpsi"].real).abs().max().item()
        dim = ((lp.imag - ref["logpsi"].imag + math.pi) % (2 * math.pi) - math.pi).abs().max().item()
        fe, fk = q(e), q(k)
        print(f"| {name} | {label} | {fe[0]:.1e} / {fe[1]:.1e} / {fe[2]:.1e} | {fk[0]:.1e} / {fk[1]:.1e} / {fk[2]:.1e} | {dre:.1e} | {dim:.1e} |", flush=True)
        del plan
