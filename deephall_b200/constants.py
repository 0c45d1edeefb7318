"""Device parallelism of the hot path (deephall/constants.py:29-41).

The reference runs one `jax.pmap` replica per local device and `lax.pmean`s scalars.  Here
there is one *process* per GPU (torchrun); `pmean` is an NCCL (or gloo, in CPU tests)
all-reduce over the default process group and the identity when no group is initialised.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def pmean(x: torch.Tensor) -> torch.Tensor:
    """Mean over ranks (constants.py:40-41).  Complex tensors are reduced via their real view."""
    w = world_size()
    if w == 1:
        return x
    if x.is_complex():
        y = torch.view_as_real(x.clone().contiguous())
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return torch.view_as_complex(y) / w
    y = x.clone().contiguous()
    dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y / w


def pmean_async(x: torch.Tensor):
    """Mean over ranks started now, finished when the returned function is called: the collective runs on NCCL's own
    stream while the caller keeps launching work that does not depend on it (the KFAC curvature pass, optimizers/kfac.py)."""
    w = world_size()
    if w == 1:
        return lambda: x
    y = x.clone().contiguous()
    work = dist.all_reduce(y, op=dist.ReduceOp.SUM, async_op=True)

    def finish():
        work.wait()
        return y / w

    return finish


def pmean_packed(values: list[torch.Tensor]) -> list[torch.Tensor]:
    """One all-reduce for a list of scalars (packed stats vector)."""
    w = world_size()
    if w == 1:
        return values
    parts = [torch.view_as_real(v.reshape(1)).reshape(-1) if v.is_complex() else v.reshape(1) for v in values]
    packed = torch.cat([p.to(torch.float32) for p in parts])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    packed = packed / w
    out, off = [], 0
    for v in values:
        if v.is_complex():
            out.append(torch.complex(packed[off], packed[off + 1]))
            off += 2
        else:
            out.append(packed[off].to(v.dtype))
            off += 1
    return out
