"""Generic A/B of an environment knob on the c3 local-energy pass: python scripts/gpu_env_ab.py VAR a b [reps].
Prints the pass time and the per-category CUDA-event times (profiling pass, single stream) of each arm."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, ".")

if sys.argv[1] == "arm":
    from deephall_b200 import _native as nat

    var = sys.argv[2]
    plan = nat.Plan(nspins=(12, 0), flux=33)
    torch.manual_seed(0)
    params = torch.randn(plan.num_params, device="cuda") * 0.05
    x = plan.init_walkers(8192, seed=1)
    for _ in range(3):
        out = plan.local_energy(params, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = plan.local_energy(params, x)
    e1.record()
    torch.cuda.synchronize()
    for i in range(2):
        plan.mcmc_sweep(params, x, steps=10, width=0.1, seed=5, offset=i * 10)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(5):
        plan.mcmc_sweep(params, x, steps=10, width=0.1, seed=5, offset=100 + i * 10)
    s1.record()
    torch.cuda.synchronize()
    print(f"{var}={os.environ.get(var)} sweep(10) {s0.elapsed_time(s1) / 5:.3f} ms", flush=True)
    plan.profile_begin()
    out = plan.local_energy(params, x)
    prof = plan.profile_end()
    cats = {k: round(v["ms"], 3) for k, v in prof.items()}
    print(f"{var}={os.environ.get(var)} local_energy {e0.elapsed_time(e1) / 5:.3f} ms  {cats}", flush=True)
    torch.save(out["energy"].cpu(), f"/tmp/ab_{os.environ.get(var)}.pt")
else:
    var, a, b = sys.argv[1:4]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    for _ in range(reps):
        for v in (a, b):
            subprocess.run([sys.executable, __file__, "arm", var], env={**os.environ, var: v}, check=True)
    print("bit-identical:", torch.equal(torch.load(f"/tmp/ab_{a}.pt"), torch.load(f"/tmp/ab_{b}.pt")))
