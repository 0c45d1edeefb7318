"""Host-side logic of the facade (CPU): config surface, IQR clipping / statistics vs the oracle,
width adaptation, RNG keys, and the world_size-2 collective path over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deephall_b200 import config as C
from deephall_b200 import constants, loss, mcmc
from oracle import loss as OLoss
from oracle import mcmc as OM


def test_config_defaults_match_reference():
    cfg = C.Config()
    assert cfg.batch_size == 3360 and cfg.system.flux == 2 and cfg.system.nspins == (3, 0)
    assert (cfg.network.psiformer.num_heads, cfg.network.psiformer.heads_dim, cfg.network.psiformer.num_layers,
            cfg.network.psiformer.determinants) == (4, 64, 2, 1)
    assert (cfg.mcmc.steps, cfg.mcmc.width, cfg.mcmc.burn_in, cfg.mcmc.adapt_frequency) == (10, 0.1, 200, 100)
    assert cfg.optim.optimizer == "kfac" and cfg.optim.adam.lr.rate == 0.005 and cfg.optim.kfac.lr.rate == 0.05
    assert abs(cfg.optim.adam.lr.schedule(2000) - 0.0025) < 1e-12


def test_config_from_dict_tolerates_extra_keys():
    cfg = C.Config.from_dict({"seed": 42, "system": {"nspins": [6, 0], "flux": 15, "bogus": 1}, "optim": {"iterations": 5}, "extra": 3})
    assert cfg.system.nspins == (6, 0) and cfg.system.flux == 15 and cfg.optim.iterations == 5 and cfg.seed == 42
    with pytest.raises(ValueError):
        C.from_dict(C.System, {"flux": 1, "nspins": 3})


def test_iqr_clip_and_stats_match_oracle():
    g = torch.Generator().manual_seed(0)
    el = torch.complex(torch.randn(257, generator=g), 0.1 * torch.randn(257, generator=g))
    el[3] = complex(1e4, -1e3)  # outlier gets clipped
    el[10] = complex(float("nan"), 0.0)  # NaN walker is skipped by nanmean / nanquantile
    a, b = loss.iqr_clip(el), OLoss.iqr_clip(el)
    ok = ~torch.isnan(b.real)
    assert torch.equal(a[ok], b[ok])
    assert a[3].real < 1e3
    assert torch.allclose(loss._nanmean(el), OLoss.nanmean_c(el))


def test_update_mcmc_width():
    pm = np.zeros(4)
    w = 0.1
    for t in range(9):
        w, pm = mcmc.update_mcmc_width(t, w, 4, torch.tensor([0.9]), pm)
        w2, _ = OM.update_mcmc_width(t, 0.1, 4, 0.9, [0.9] * 4) if t else (0.1, None)
    assert abs(w - 0.1 * 1.1 * 1.1) < 1e-12  # widened at t = 4 and t = 8
    w, pm = 0.1, np.zeros(2)
    for t in range(3):
        w, pm = mcmc.update_mcmc_width(t, w, 2, torch.tensor(0.1), pm)
    assert abs(w - 0.1 / 1.1) < 1e-12


def test_philox_key_split():
    k = mcmc.PhiloxKey(7)
    k2, sub = k.split()
    k3, sub2 = k2.split()
    assert sub.seed == 7 and sub.offset == 0 and sub2.offset == 1 << 20 and k3.offset == 2 << 20


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.tensor([float(rank + 1), 10.0 * (rank + 1)])
        m = constants.pmean(x)
        c = constants.pmean(torch.complex(torch.tensor(float(rank)), torch.tensor(2.0 * rank)))
        packed = constants.pmean_packed([torch.tensor(float(rank)), torch.complex(torch.tensor(1.0 + rank), torch.tensor(-1.0 * rank)),
                                         torch.tensor(3.0)])
        # sharded energy statistics: shard-local clip quantiles, pmean of the means (loss.py:31-32,73-74)
        g = torch.Generator().manual_seed(5)
        el_all = torch.complex(torch.randn(64, generator=g), torch.zeros(64))
        shard = el_all[rank * 32:(rank + 1) * 32]
        clipped = constants.pmean(loss._nanmean(loss.iqr_clip(shard)))
        q.put((rank, m.tolist(), complex(c), [complex(p) for p in packed], complex(clipped)))
    finally:
        dist.destroy_process_group()


def test_pmean_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    el_all = torch.complex(torch.randn(64, generator=g), torch.zeros(64))
    expect_clip = 0.5 * (OLoss.nanmean_c(OLoss.iqr_clip(el_all[:32])) + OLoss.nanmean_c(OLoss.iqr_clip(el_all[32:])))
    for rank, m, c, packed, clipped in res:
        assert m == [1.5, 15.0]
        assert abs(c - complex(0.5, 1.0)) < 1e-6
        assert abs(packed[0] - 0.5) < 1e-6 and abs(packed[1] - complex(1.5, -0.5)) < 1e-6 and abs(packed[2] - 3.0) < 1e-6
        assert abs(clipped - complex(expect_clip)) < 1e-6


def _write_reference_style_checkpoint(path, optimizer):
    """A checkpoint as deephall/log.py:174-178 writes it -- jax.Array leaves pickled through
    `jax._src.array._reconstruct_array`, optax NamedTuples / a kfac_jax dataclass in `opt_state` -- produced with throw-away
    modules of those names (jax is not installed here), which are removed again before the file is read."""
    import dataclasses
    import sys
    import types
    from typing import NamedTuple

    def module(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    names = ["jax", "jax._src", "jax._src.array", "optax", "optax._src", "optax._src.transform", "optax._src.base",
             "kfac_jax", "kfac_jax._src", "kfac_jax._src.optimizer"]
    assert not any(n in sys.modules for n in names)
    mods = {n: module(n) for n in names}
    try:
        def _reconstruct_array(fun, args, arr_state, aval_state):
            raise AssertionError("the writer's reconstructor must not run in the reader")

        class ArrayImpl:  # jax/_src/array.py: __reduce__ of a committed array
            def __init__(self, value):
                self._value = np.asarray(value)

            def __reduce__(self):
                fun, args, arr_state = self._value.__reduce__()
                return (_reconstruct_array, (fun, args, arr_state, {"weak_type": False}))

        _reconstruct_array.__module__ = ArrayImpl.__module__ = "jax._src.array"
        _reconstruct_array.__qualname__ = "_reconstruct_array"
        ArrayImpl.__qualname__ = "ArrayImpl"
        mods["jax._src.array"]._reconstruct_array = _reconstruct_array
        mods["jax._src.array"].ArrayImpl = ArrayImpl

        class ScaleByAdamState(NamedTuple):
            count: object
            mu: object
            nu: object

        class ScaleByScheduleState(NamedTuple):
            count: object

        for cls in (ScaleByAdamState, ScaleByScheduleState):
            cls.__module__, cls.__qualname__ = "optax._src.transform", cls.__name__
            setattr(mods["optax._src.transform"], cls.__name__, cls)

        class Optimizer:
            @dataclasses.dataclass
            class State:
                velocities: object
                estimator_state: object
                damping: object
                data_seen: object
                step_counter: object

        Optimizer.__module__, Optimizer.__qualname__ = "kfac_jax._src.optimizer", "Optimizer"
        Optimizer.State.__module__, Optimizer.State.__qualname__ = "kfac_jax._src.optimizer", "Optimizer.State"
        mods["kfac_jax._src.optimizer"].Optimizer = Optimizer

        g = np.random.default_rng(0)
        tree = {"params": {"Dense_0": {"kernel": g.normal(size=(4, 8)).astype(np.float32)},
                           "Jastrow_0": {"ee_par": g.normal(size=(1,)).astype(np.float32)}}}
        wrap = lambda t: {k: wrap(v) for k, v in t.items()} if isinstance(t, dict) else ArrayImpl(t)  # noqa: E731
        data = g.uniform(0, 3, size=(16, 3, 2)).astype(np.float32)
        if optimizer == "adam":
            mu = {"params": {"Dense_0": {"kernel": np.full((4, 8), 0.25, np.float32)}, "Jastrow_0": {"ee_par": np.full((1,), 0.5, np.float32)}}}
            nu = {"params": {"Dense_0": {"kernel": np.full((4, 8), 2.0, np.float32)}, "Jastrow_0": {"ee_par": np.full((1,), 3.0, np.float32)}}}
            opt = (ScaleByAdamState(ArrayImpl(np.int32(7)), wrap(mu), wrap(nu)), ScaleByScheduleState(ArrayImpl(np.int32(7))))
        elif optimizer == "kfac":
            opt = Optimizer.State(wrap(tree), {"blocks": [ArrayImpl(np.eye(3, dtype=np.float32))]}, ArrayImpl(np.float32(1e-3)),
                                  ArrayImpl(np.int32(64)), ArrayImpl(np.int32(4)))
        else:
            opt = None
        with open(path, "wb") as f:  # log.py:178 (np.asarray(..., dtype="object") of the optimizer state: log.py:55)
            np.savez_compressed(f, step=41, params=wrap(tree), data=data, opt_state=np.asarray(opt, dtype="object"),
                                mcmc_width=np.float32(0.17))
        return tree, data
    finally:
        for n in names:
            sys.modules.pop(n, None)


class _TreeModel:
    """from_tree of a two-leaf network (the real one needs a plan, i.e. a GPU)."""
    layout = {"Dense_0/kernel": (0, (4, 8)), "Jastrow_0/ee_par": (32, (1,))}

    def from_tree(self, tree, device="cpu"):
        root = tree.get("params", tree)
        flat = torch.empty(33)
        for name, (off, shape) in self.layout.items():
            node = root
            for part in name.split("/"):
                node = node[part]
            arr = torch.as_tensor(np.asarray(node), dtype=torch.float32)
            assert tuple(arr.shape) == shape
            flat[off : off + arr.numel()] = arr.reshape(-1)
        return flat.to(device)


@pytest.mark.parametrize("optimizer", ["adam", "kfac", "none"])
def test_restore_checkpoint_written_by_the_reference(tmp_path, optimizer):
    """restore_checkpoint reads the reference's own files (log.py:174-216) where jax / optax / kfac_jax are not installed:
    jax.Array leaves come back as numpy values, the optax Adam state maps to AdamState, a kfac_jax state is dropped with a
    warning; no stand-in module survives the call."""
    import sys
    import warnings

    from deephall_b200 import checkpoint
    from deephall_b200.optimizers import AdamState

    path = tmp_path / "ckpt_000041.npz"
    tree, data = _write_reference_style_checkpoint(path, optimizer)
    model = _TreeModel()
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        step, st = checkpoint.restore_checkpoint(path, model, device="cpu")
    assert step == 42 and abs(st.mcmc_width - 0.17) < 1e-7
    assert torch.equal(st.params, model.from_tree(tree)) and torch.equal(st.data, torch.as_tensor(data))
    if optimizer == "adam":
        assert isinstance(st.opt_state, AdamState) and st.opt_state.count == 7
        assert torch.all(st.opt_state.mu[:32] == 0.25) and st.opt_state.mu[32] == 0.5
        assert torch.all(st.opt_state.nu[:32] == 2.0) and st.opt_state.nu[32] == 3.0
    else:
        assert st.opt_state is None
    assert (len([w for w in caught if "kfac_jax" in str(w.message)]) == 1) == (optimizer == "kfac")
    assert not any(m.split(".")[0] in ("jax", "optax", "kfac_jax") for m in sys.modules)
