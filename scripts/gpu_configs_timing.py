"""Developer tool: log psi / local energy / sweep timings of the BASELINE.json configurations."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
from deephall_b200 import networks

CASES = [("c2", dict(nspins=(6, 0), flux=15), 4096), ("c3", dict(nspins=(12, 0), flux=33), 8192),
         ("c4", dict(nspins=(10, 0), flux=21), 8192), ("c5 K=4", dict(nspins=(16, 0), flux=45, ndets=4), 2048),
         ("c5 K=16", dict(nspins=(16, 0), flux=45, ndets=16), 2048)]


def t(fn, n=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, kw, B in CASES:
    plan = nat.Plan(**kw)
    params = networks.Psiformer(kw["nspins"], kw["flux"] / 2, ndets=kw.get("ndets", 1)).init(0)  # flax init distributions
    x = plan.init_walkers(B, seed=1)
    plan.mcmc_sweep(params, x, 20, 0.1, seed=3)
    ms_lp = t(lambda: plan.logpsi(params, x))
    ms_le = t(lambda: plan.local_energy(params, x))
    ms_mc = t(lambda: plan.mcmc_sweep(params, x, 10, 0.1, seed=5))
    cot = torch.randn(B, 2, device="cuda") / B
    ms_vjp = t(lambda: plan.logpsi_vjp(params, x, cot))
    out = plan.local_energy(params, x)
    ok = bool(torch.isfinite(out["energy"].real).all())
    ws = plan._ws.numel() / 2**30
    print(f"{name:8s} B={B}: logpsi {ms_lp:7.2f} ms | local_energy {ms_le:8.2f} ms ({B / ms_le * 1e3:9.0f} evals/s) | "
          f"10-move sweep {ms_mc:7.2f} ms ({B * 10 / ms_mc * 1e3:10.0f} walker-steps/s) | vjp {ms_vjp:7.2f} ms | "
          f"E mean {float(out['energy'].real.mean()):.4f} finite={ok} | workspace {ws:.1f} GiB", flush=True)
    del plan
    torch.cuda.empty_cache()
