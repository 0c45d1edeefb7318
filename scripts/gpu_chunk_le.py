"""Developer diagnostic: local-energy pass time (c3, 8192 walkers) against the walkers-per-chunk setting."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

torch.manual_seed(0)
for chunk in (0, 512, 768, 1024, 1365, 2048, 4096):
    plan = nat.Plan(nspins=(12, 0), flux=33, chunk_walkers=chunk)
    params = torch.randn(plan.num_params, device="cuda") * 0.05
    x = plan.init_walkers(8192, seed=1)
    for _ in range(3):
        out = plan.local_energy(params, x)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(8):
        out = plan.local_energy(params, x)
    e[1].record()
    torch.cuda.synchronize()
    print(f"chunk_walkers={chunk}: {e[0].elapsed_time(e[1]) / 8:.3f} ms", flush=True)
    del plan
    torch.cuda.empty_cache()
