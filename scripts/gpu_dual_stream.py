"""Chunk interleave (DH_DUAL_STREAM) A/B: local-energy pass and logpsi at c3, B = 8192, CUDA-event timed; the two
arms must agree bit for bit (chunks are independent, the kernels deterministic)."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, ".")

if len(sys.argv) > 1 and sys.argv[1] == "arm":
    from deephall_b200 import _native as nat

    cfgs = {"c3": dict(nspins=(12, 0), flux=33), "c4": dict(nspins=(10, 0), flux=21), "c2": dict(nspins=(6, 0), flux=15)}
    for name, kw in cfgs.items():
        B = 8192 if name != "c2" else 4096
        plan = nat.Plan(**kw)
        torch.manual_seed(0)
        params = torch.randn(plan.num_params, device="cuda") * 0.05
        x = plan.init_walkers(B, seed=1)
        for _ in range(3):
            out = plan.local_energy(params, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = plan.local_energy(params, x)
        e1.record()
        torch.cuda.synchronize()
        cs = out["energy"].double().abs().sum().item() if "energy" in out else 0.0
        print(f"{name} dual={os.environ.get('DH_DUAL_STREAM', '1')} local_energy {e0.elapsed_time(e1) / 5:.3f} ms  checksum {cs!r}", flush=True)
        torch.save({k: v.cpu() for k, v in out.items()}, f"/tmp/dual_{name}_{os.environ.get('DH_DUAL_STREAM', '1')}.pt")
else:
    for d in ("1", "0", "1", "0"):
        subprocess.run([sys.executable, __file__, "arm"], env={**os.environ, "DH_DUAL_STREAM": d}, check=True)
    for name in ("c3", "c4", "c2"):
        a, b = torch.load(f"/tmp/dual_{name}_0.pt"), torch.load(f"/tmp/dual_{name}_1.pt")
        same = all(torch.equal(a[k], b[k]) for k in a)
        print(name, "bit-identical:", same)
        assert same
