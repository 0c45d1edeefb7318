"""Developer diagnostic: dh_logpsi_vjp at c3, 8192 walkers: total and per-category CUDA-event times."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

plan = nat.Plan(nspins=(12, 0), flux=33)
torch.manual_seed(0)
params = torch.randn(plan.num_params, device="cuda") * 0.05
for B in (8192, 1024):
    x = plan.init_walkers(B, seed=1)
    cot = torch.randn(B, 2, device="cuda")
    for _ in range(3):
        g = plan.logpsi_vjp(params, x, cot)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(10):
        g = plan.logpsi_vjp(params, x, cot)
    e[1].record()
    torch.cuda.synchronize()
    plan.profile_begin()
    plan.logpsi_vjp(params, x, cot)
    prof = plan.profile_end()
    print(f"B={B}: vjp {e[0].elapsed_time(e[1]) / 10:.3f} ms", {k: (round(v['ms'], 3), v['count']) for k, v in prof.items() if isinstance(v, dict)})
