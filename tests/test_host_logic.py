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
