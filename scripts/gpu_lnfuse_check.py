"""Developer check: local energy at c3 with the LayerNorm fused into the contraction epilogue (DH_LN_FUSE=1) against the
separate LayerNorm kernels; per-category times of both."""
import os
import sys

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402
from deephall_b200 import networks  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
outs, profs = {}, {}
params = networks.Psiformer((12, 0), 16.5).init(0)
for mode in ("0", "1"):
    os.environ["DH_LN_FUSE"] = mode
    plan = nat.Plan(nspins=(12, 0), flux=33)
    x = plan.init_walkers(B, seed=1)
    plan.mcmc_sweep(params, x, 20, 0.1, seed=3)
    for _ in range(2):
        out = plan.local_energy(params, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = plan.local_energy(params, x)
    e1.record()
    torch.cuda.synchronize()
    plan.profile_begin()
    plan.local_energy(params, x)
    prof = plan.profile_end()
    outs[mode] = out
    print(f"DH_LN_FUSE={mode}: {e0.elapsed_time(e1) / 3:.2f} ms", {k: round(v['ms'], 2) for k, v in prof.items() if v['count']}, flush=True)
for k in ("energy", "kinetic", "angular_momentum_square", "logpsi"):
    a, b = outs["1"][k], outs["0"][k]
    d = (a - b).abs() / b.abs().clamp(min=1.0)
    print(f"{k}: fused vs separate, relative difference median {d.median().item():.2e} max {d.max().item():.2e} finite {bool(torch.isfinite(torch.view_as_real(a) if a.is_complex() else a).all())}")
