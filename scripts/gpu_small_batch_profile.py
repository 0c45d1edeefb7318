"""Developer diagnostic: per-category kernel time of log psi and of the local-energy pass at small walkers-per-GPU counts."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
plan = nat.Plan(nspins=(12, 0), flux=33)
params = torch.randn(plan.num_params, device="cuda") * 0.05
for B in (8192, 1024):
    x = plan.init_walkers(B, seed=1)
    for name, fn in (("logpsi", lambda: plan.logpsi(params, x)), ("local_energy", lambda: plan.local_energy(params, x)),
                     ("vjp", lambda: plan.logpsi_vjp(params, x, torch.ones(B, 2, device="cuda") / B))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        wall = e0.elapsed_time(e1) / 5
        plan.profile_begin(); fn(); prof = plan.profile_end()
        tot = sum(v["ms"] for v in prof.values()); cnt = sum(v["count"] for v in prof.values())
        print(f"B={B} {name}: wall {wall:.3f} ms | kernels {tot:.3f} ms in {cnt} launches | " + ", ".join(f"{k} {v['ms']:.3f}/{v['count']}" for k, v in prof.items() if v["count"]))
