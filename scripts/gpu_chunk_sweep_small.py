"""Developer diagnostic: local-energy pass time at c3, 8192 walkers, against the walkers-per-chunk setting (L2 reuse
between producer and consumer kernels vs launch count and tail effects)."""
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
B = 8192
res = []
for chunk in (0, 2048, 1024, 512, 256, 128, 64):
    plan = nat.Plan(nspins=(12, 0), flux=33, chunk_walkers=chunk)
    params = torch.randn(plan.num_params, device="cuda") * 0.05
    x = plan.init_walkers(B, seed=1)
    for _ in range(2): plan.local_energy(params, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = plan.local_energy(params, x)
    e1.record(); torch.cuda.synchronize()
    print(f"chunk {chunk}: {e0.elapsed_time(e1) / 3:.2f} ms per pass, workspace {plan._ws.numel() / 2**30:.2f} GiB, E {float(out['energy'].real.mean()):.5f}", flush=True)
    del plan
