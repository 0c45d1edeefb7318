"""Where a KFAC iteration's time goes at c3 (8192 walkers): torch profiler over three steps, top CUDA kernels and
host-side totals."""
import dataclasses
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from deephall_b200.config import Config, Optim, System
from deephall_b200.train import VMC

cfg = Config(batch_size=8192, seed=0, system=System(flux=33, nspins=(12, 0)), optim=Optim(optimizer="kfac"))
vmc = VMC(cfg)
vmc.burn_in(3)
for _ in range(3):
    vmc.step(sync_stats=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    vmc.step(sync_stats=False)
e1.record()
torch.cuda.synchronize()
print(f"kfac step {e0.elapsed_time(e1) / 3:.2f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        vmc.step(sync_stats=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
