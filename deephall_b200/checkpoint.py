"""Checkpoint files in the reference's wire format (deephall/log.py:49-67,174-216).

The reference writes `ckpt_{step:06d}.npz` = `np.savez_compressed(step, params, data, opt_state,
mcmc_width)`: `params` is the flax tree `{'params': {...}}` pickled inside a 0-d object array,
`data` the walkers `(B, N, 2)` with the device axis folded away, `mcmc_width` a scalar, and
`restore_checkpoint` resumes at `step + 1`.

`save_checkpoint` writes that layout with numpy leaves, so `LogManager.restore_checkpoint`
(log.py:193-216) and `netobs_bridge/adaptor.py:43-65` of the reference can read it.
`restore_checkpoint` reads our files AND files written by the reference itself.  Those pickle
`jax.Array` leaves (`jax._src.array._reconstruct_array(fun, args, arr_state, aval_state)`: a numpy
array reconstructor plus the aval's weak_type) and, in `opt_state`, optax NamedTuples / kfac_jax
dataclasses.  Where jax is not installed (this image), the unpickling runs under `_foreign_pickles()`:
stand-in modules for `jax`, `jaxlib`, `flax`, `optax`, `kfac_jax`, `chex` whose `_reconstruct_array`
returns the numpy array and whose every other name is an inert stand-in class -- so parameters, walkers,
step and MCMC width of a reference-trained run load into the GPU path (the NetObs use case,
`netobs_bridge/adaptor.py:43-65`).  An optax Adam state (`ScaleByAdamState(count, mu, nu)`) maps to
`AdamState`; a kfac_jax state does not map to this library's factor layout: `opt_state` comes back
None (the curvature averages restart), with a warning.
One process per GPU: each rank saves / restores its own walker shard; parameters are replicated.
"""
from __future__ import annotations

import contextlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types
import warnings

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


_FOREIGN_ROOTS = ("jax", "jaxlib", "flax", "optax", "kfac_jax", "chex")


def _reconstruct_array(fun, args, arr_state, aval_state=None):
    """Stand-in for jax._src.array._reconstruct_array: the numpy value, without the device_put."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


class _OpaqueMeta(type):
    def __getattr__(cls, name):  # nested classes (kfac_jax's `Optimizer.State`) resolve to further stand-ins
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _OpaqueMeta(name, (_Opaque,), {"__module__": cls.__module__, "__qualname__": f"{cls.__qualname__}.{name}"})
        setattr(cls, name, sub)
        return sub


class _Opaque(metaclass=_OpaqueMeta):
    """An object of a class this process does not have: constructor arguments in `.args` (NamedTuples), attributes as
    pickled (dataclasses)."""

    def __new__(cls, *args, **kwargs):
        obj = object.__new__(cls)
        obj.args = args
        return obj

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):  # (dict, slots)
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.state = state


class _StandInModule(types.ModuleType):
    __path__: list = []  # a package: any sub-module can be imported

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _reconstruct_array if name == "_reconstruct_array" else _OpaqueMeta(name, (_Opaque,), {"__module__": self.__name__})
        setattr(self, name, obj)
        return obj


class _StandInFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _FOREIGN_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StandInModule(spec.name)

    def exec_module(self, module):
        pass


@contextlib.contextmanager
def _foreign_pickles():
    """Lets pickle resolve names of packages that are not installed here (see the module docstring).  Packages that ARE
    installed are left alone."""
    missing = [r for r in _FOREIGN_ROOTS if r not in sys.modules and importlib.util.find_spec(r) is None]
    if not missing:
        yield
        return
    finder = _StandInFinder()
    sys.meta_path.append(finder)  # (last: a package that is installed is found by the regular finders first)
    try:
        yield
    finally:
        sys.meta_path.remove(finder)
        for name in [m for m in sys.modules if m.split(".")[0] in missing and isinstance(sys.modules[m], _StandInModule)]:
            del sys.modules[name]


def _find_adam_state(opt):
    """optax.adam's state inside the reference's `opt_state` (optimizers/adam.py:30: chain(scale_by_adam, scale_by_schedule)):
    the element with (count, mu, nu), as a real NamedTuple or as its stand-in."""
    items = opt if isinstance(opt, (list, tuple)) else [opt]
    for it in items:
        fields = getattr(it, "_fields", None)
        if fields and tuple(fields[:3]) == ("count", "mu", "nu"):
            return it.count, it.mu, it.nu
        if isinstance(it, _Opaque) and type(it).__name__ == "ScaleByAdamState" and len(it.args) == 3:
            return it.args
        if isinstance(it, (list, tuple)) and not fields:
            got = _find_adam_state(list(it))
            if got is not None:
                return got
    return None


def restore_checkpoint(path, model, device="cuda") -> tuple[int, CheckpointState]:
    """Returns (next step, state) like log.py:193-216."""
    with open(path, "rb") as npf, _foreign_pickles(), np.load(npf, allow_pickle=True) as f:
        step = int(f["step"].tolist()) + 1
        tree = f["params"].tolist()
        params = model.from_tree(tree, device=device) if tree.get("params", tree) else torch.zeros(0, device=device)
        data = torch.as_tensor(np.asarray(f["data"], dtype=np.float32)).to(device).contiguous()
        opt = f["opt_state"].tolist()
        adam = None if isinstance(opt, dict) or opt is None else _find_adam_state(opt)
        if adam is not None:  # a reference checkpoint written with optimizer=adam
            opt = AdamState(int(np.asarray(adam[0])), model.from_tree(adam[1], device=device), model.from_tree(adam[2], device=device))
        elif opt is not None and not isinstance(opt, dict):
            warnings.warn(f"restore_checkpoint: optimizer state of type {type(opt).__name__} (written by the reference's "
                          "kfac_jax) has no counterpart in this library's factor layout; the optimizer starts afresh")
            opt = None
        if isinstance(opt, dict) and {"count", "mu", "nu"} <= set(opt):
            opt = AdamState(int(opt["count"]), model.from_tree(opt["mu"], device=device), model.from_tree(opt["nu"], device=device))
        elif isinstance(opt, dict) and {"kfac_step", "weight", "stats", "dense0_xtx"} <= set(opt):
            from .kfac import KfacState

            opt = KfacState(int(opt["kfac_step"]), float(opt["weight"]),
                            torch.as_tensor(np.asarray(opt["stats"], dtype=np.float32)).to(device),
                            torch.as_tensor(np.asarray(opt["dense0_xtx"], dtype=np.float32)).to(device))
        width = float(np.asarray(f["mcmc_width"]).reshape(-1)[0])
    return step, CheckpointState(params, data, opt, width)
