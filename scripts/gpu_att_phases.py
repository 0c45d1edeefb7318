"""Developer build only (make DEBUG=1): per-phase clock sums of the tensor-core jet attention (thread 0 of every block),
one local-energy pass at c3.  usage: python scripts/gpu_att_phases.py [walkers]"""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
# "l0" as second argument: a one-layer network, so that only the first-layer kernel runs (both kernels add into the same slots)
L0_ONLY = len(sys.argv) > 2 and sys.argv[2] == "l0"
plan = nat.Plan(nspins=(12, 0), flux=33, **({"num_layers": 1} if L0_ONLY else {}))
lib = nat.load()
torch.manual_seed(0)
params = torch.randn(plan.num_params, device="cuda") * 0.05
x = plan.init_walkers(B, seed=1)
plan.local_energy(params, x)
buf = (ctypes.c_ulonglong * 16)()
lib.dh_debug_at_prof(buf, 1)
plan.local_energy(params, x)
lib.dh_debug_at_prof(buf, 1)
names = ["blocks", "whole kernel", "score steps: wait + barrier", "score steps: multiply", "score write-out", "softmax",
         "P fragments", "P.V steps: wait + barrier", "P.V steps: multiply + stores"]
n = max(buf[0], 1)
print(f"blocks {buf[0]}")
for i in range(1, 9):
    print(f"{names[i]:32s} {buf[i] / n:10.0f} cycles per block  {100.0 * buf[i] / max(buf[1], 1):5.1f} %")
