"""Developer diagnostic: where a KFAC step's extra time goes at c3, B = 8192."""
import sys
import time

import torch

sys.path.insert(0, ".")
from deephall_b200 import _native as nat  # noqa: E402


def t(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


plan = nat.Plan(nspins=(12, 0), flux=33)
flat = torch.randn(plan.num_params, device="cuda") * 0.05
x = plan.init_walkers(8192, seed=1)
cot = torch.randn(8192, 2, device="cuda") / 8192
print(f"logpsi_vjp   {t(lambda: plan.logpsi_vjp(flat, x, cot)):7.2f} ms")
print(f"kfac_factors {t(lambda: plan.kfac_factors(flat, x)):7.2f} ms")
for n, b in ((257, 13), (256, 14), (408, 2)):
    m = torch.randn(b, n, n, device="cuda")
    m = m @ m.transpose(1, 2) / n + 0.1 * torch.eye(n, device="cuda")
    print(f"{b} x {n}^2: inv {t(lambda: torch.linalg.inv(m)):6.2f} ms  cholesky_inverse {t(lambda: torch.cholesky_inverse(torch.linalg.cholesky(m))):6.2f} ms  "
          f"one-by-one inv {t(lambda: [torch.linalg.inv(m[i]) for i in range(b)]):6.2f} ms  dh_spd_inverse {t(lambda: nat.spd_inverse(m)):6.2f} ms  "
          f"max |dh - torch| / max |torch| {((nat.spd_inverse(m) - torch.linalg.inv(m)).abs().max() / torch.linalg.inv(m).abs().max()).item():.1e}")
