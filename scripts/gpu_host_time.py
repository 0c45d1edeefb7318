import sys, time, torch
sys.path.insert(0, ".")
from deephall_b200 import _native as nat
plan = nat.Plan(nspins=(12, 0), flux=33)
params = torch.randn(plan.num_params, device="cuda") * 0.05
x = plan.init_walkers(8192, seed=1)
for _ in range(3): plan.mcmc_sweep(params, x, 10, 0.1, seed=3)
torch.cuda.synchronize()
t0 = time.perf_counter(); plan.mcmc_sweep(params, x, 10, 0.1, seed=3, offset=50); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"graph path: host enqueue {1e3*(t1-t0):.3f} ms, total {1e3*(t2-t0):.3f} ms")
t0 = time.perf_counter()
for st in range(10): plan.mcmc_sweep(params, x, 1, 0.1, seed=3, offset=70 + st)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"launch-by-launch (10 x 1 move, each with its own entry forward): host enqueue {1e3*(t1-t0):.3f} ms, total {1e3*(t2-t0):.3f} ms")
