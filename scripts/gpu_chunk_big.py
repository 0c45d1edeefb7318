import sys, torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
B = 8192
ref = None
for ch in (1024, 2048, 4096, 1024, 2048):
    plan = nat.Plan(nspins=(12, 0), flux=33, chunk_walkers=ch)
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
    el = out["energy"].cpu()
    if ref is None: ref = el
    print(f"chunk {ch}: {e0.elapsed_time(e1)/5:.3f} ms  identical to 1024: {torch.equal(el, ref)}  finite {torch.isfinite(torch.view_as_real(el)).all().item()}", flush=True)
    del plan, out
    torch.cuda.empty_cache()
