"""Developer diagnostic: local-energy time at c3, B = 8192, against the walkers-per-pass chunk size
(smaller chunks keep a pass's activations inside the 126 MB L2 at the price of more, smaller launches)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402

B = 8192
for chunk in [int(a) for a in sys.argv[1:]] or [0, 512, 256, 128, 96, 64, 37]:
    plan = nat.Plan(nspins=(12, 0), flux=33, chunk_walkers=chunk)
    flat = torch.randn(plan.num_params, device="cuda") * 0.05
    x = plan.init_walkers(B, seed=1)
    plan.mcmc_sweep(flat, x, 20, 0.1, seed=0)
    for _ in range(2):
        plan.local_energy(flat, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(3):
        plan.local_energy(flat, x)
    e1.record()
    t_launch = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()
    plan.profile_begin()
    plan.local_energy(flat, x)
    prof = plan.profile_end()
    cats = {k: (round(v["ms"], 2), v["count"]) for k, v in prof.items() if v["count"]}
    print(f"chunk {chunk:5d}: local energy {e0.elapsed_time(e1) / 3:7.2f} ms  (host launch loop {t_launch:6.2f} ms)  per category (ms, launches): {cats}", flush=True)
    del plan
