"""Half-batch interleave of the value-only pass (DH_DUAL_VALUE_MIN) A/B: Metropolis sweep (10 moves) and log psi at
c3 / c4 / c2, CUDA-event timed; final walkers must agree bit for bit between the arms."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, ".")

if len(sys.argv) > 1 and sys.argv[1] == "arm":
    from deephall_b200 import _native as nat

    tag = os.environ.get("DH_DUAL_VALUE_MIN", "2048")
    cfgs = {"c3": dict(nspins=(12, 0), flux=33), "c4": dict(nspins=(10, 0), flux=21), "c2": dict(nspins=(6, 0), flux=15)}
    for name, kw in cfgs.items():
        B = 8192 if name != "c2" else 4096
        plan = nat.Plan(**kw)
        torch.manual_seed(0)
        params = torch.randn(plan.num_params, device="cuda") * 0.05
        x = plan.init_walkers(B, seed=1)
        for i in range(3):
            plan.mcmc_sweep(params, x, steps=10, width=0.1, seed=5, offset=i * 10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(5):
            plan.mcmc_sweep(params, x, steps=10, width=0.1, seed=5, offset=100 + i * 10)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name} value_min={tag} sweep(10) {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
        torch.save(x.cpu(), f"/tmp/dualv_{name}_{tag}.pt")
else:
    for d in ("2048", "0", "2048", "0"):
        subprocess.run([sys.executable, __file__, "arm"], env={**os.environ, "DH_DUAL_VALUE_MIN": d}, check=True)
    for name in ("c3", "c4", "c2"):
        same = torch.equal(torch.load(f"/tmp/dualv_{name}_0.pt"), torch.load(f"/tmp/dualv_{name}_2048.pt"))
        print(name, "walkers bit-identical:", same)
        assert same
