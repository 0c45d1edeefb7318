"""Developer check: a few VMC iterations on N ranks (torchrun), one process per GPU, with either optimizer.
usage: torchrun --nproc-per-node 2 scripts/gpu_vmc_multigpu.py [kfac|adam]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from deephall_b200.config import Config, Network, Optim, PsiformerNetwork, System  # noqa: E402
from deephall_b200.train import VMC  # noqa: E402

opt = sys.argv[1] if len(sys.argv) > 1 else "kfac"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = Config(batch_size=1024 * world, seed=7, system=System(flux=2, nspins=(3, 0), interaction_strength=0.0),
             network=Network(psiformer=PsiformerNetwork(num_heads=2, heads_dim=16, num_layers=1)),
             optim=Optim(iterations=40, optimizer=opt))
vmc = VMC(cfg)
vmc.burn_in(10)
for it in range(40):
    _, st = vmc.step()
    if rank == 0 and it % 10 == 9:
        print(f"[{opt}, {world} rank(s)] iteration {it + 1}: energy {float(st['energy'].real):.4f} variance {float(st['variance']):.4f}", flush=True)
# replicas must hold identical parameters (all-reduced statistics and gradients)
p = vmc.state.params.clone()
if world > 1:
    ref = p.clone()
    dist.broadcast(ref, 0)
    print(f"rank {rank}: max |params - rank 0 params| = {float((p - ref).abs().max()):.3e}", flush=True)
    dist.destroy_process_group()
