"""Checkpoint files in the reference's wire format (deephall/log.py:49-67,174-216).

The reference writes `ckpt_{step:06d}.npz` = `np.savez_compressed(step, params, data, opt_state,
mcmc_width)`: `params` is the flax tree `{'params': {...}}` pickled inside a 0-d object array,
`data` the walkers `(B, N, 2)` with the device axis folded away, `mcmc_width` a scalar, and
`restore_checkpoint` resumes at `step + 1`.

`save_checkpoint` writes that layout with numpy leaves, so `LogManager.restore_checkpoint`
(log.py:193-216) and `netobs_bridge/adaptor.py:43-65` of the reference can read it.
`restore_checkpoint` reads files whose leaves are numpy arrays (ours, or a reference checkpoint
re-saved with `jax.tree.map(np.asarray, ...)` on a machine that has jax: unpickling jax.Array
leaves needs jax, which this image does not have).
One process per GPU: each rank saves / restores its own walker shard; parameters are replicated.
"""
from __future__ import annotations

import numpy as np
import torch

from .optimizers import AdamState, CheckpointState


def _is_kfac_state(opt) -> bool:
    return hasattr(opt, "_fields") and tuple(opt._fields) == ("step", "weight", "stats", "dense0_xtx")


def _to_numpy_tree(tree):
    if isinstance(tree, dict):
        return {k: _to_numpy_tree(v) for k, v in tree.items()}
    return tree.detach().cpu().numpy() if isinstance(tree, torch.Tensor) else np.asarray(tree)


def save_checkpoint(path, step: int, model, state: CheckpointState) -> None:
    params = _to_numpy_tree(model.param_tree(state.params)) if state.params.numel() else {"params": {}}
    opt = state.opt_state
    if isinstance(opt, AdamState):
        opt = {"count": opt.count, "mu": _to_numpy_tree(model.param_tree(opt.mu)), "nu": _to_numpy_tree(model.param_tree(opt.nu))}
    elif _is_kfac_state(opt):  # kfac.KfacState (the default optimizer): moving averages in the factor-vector layout
        opt = {"kfac_step": int(opt.step), "weight": float(opt.weight), "stats": opt.stats.detach().cpu().numpy(),
               "dense0_xtx": opt.dense0_xtx.detach().cpu().numpy()}
    elif opt is not None:
        raise TypeError(f"save_checkpoint: unknown optimizer state {type(opt).__name__}")
    with open(path, "wb") as f:
        np.savez_compressed(f, step=step, params=np.asarray(params, dtype="object"), data=state.data.detach().cpu().numpy(),
                            opt_state=np.asarray(opt, dtype="object"), mcmc_width=np.float32(state.mcmc_width))


def restore_checkpoint(path, model, device="cuda") -> tuple[int, CheckpointState]:
    """Returns (next step, state) like log.py:193-216."""
    with open(path, "rb") as npf, np.load(npf, allow_pickle=True) as f:
        step = int(f["step"].tolist()) + 1
        tree = f["params"].tolist()
        params = model.from_tree(tree, device=device) if tree.get("params", tree) else torch.zeros(0, device=device)
        data = torch.as_tensor(np.asarray(f["data"], dtype=np.float32)).to(device).contiguous()
        opt = f["opt_state"].tolist()
        if isinstance(opt, dict) and {"count", "mu", "nu"} <= set(opt):
            opt = AdamState(int(opt["count"]), model.from_tree(opt["mu"], device=device), model.from_tree(opt["nu"], device=device))
        elif isinstance(opt, dict) and {"kfac_step", "weight", "stats", "dense0_xtx"} <= set(opt):
            from .kfac import KfacState

            opt = KfacState(int(opt["kfac_step"]), float(opt["weight"]),
                            torch.as_tensor(np.asarray(opt["stats"], dtype=np.float32)).to(device),
                            torch.as_tensor(np.asarray(opt["dense0_xtx"], dtype=np.float32)).to(device))
        width = float(np.asarray(f["mcmc_width"]).reshape(-1)[0])
    return step, CheckpointState(params, data, opt, width)
