"""Configuration surface of the hot path.

Field names and defaults mirror deephall/config.py:56-214 so that a reference config (dict /
YAML) can be applied unchanged; OmegaConf is not required (plain dataclasses + `from_dict`).
Only the fields the walker-evaluation path reads are interpreted; `Log` is kept as a passive
record.
"""
from __future__ import annotations

import dataclasses as dc
import time
from typing import Any


@dc.dataclass
class System:  # config.py:56-79
    flux: int = 2
    radius: float | None = None
    nspins: tuple[int, int] = (3, 0)
    interaction_strength: float = 1.0
    lz_center: float = 0.0
    lz_penalty: float = 0.0
    l2_penalty: float = 0.0
    interaction_type: str = "coulomb"  # "coulomb" | "harmonic"


@dc.dataclass
class PsiformerNetwork:  # config.py:92-97
    num_heads: int = 4
    heads_dim: int = 64
    num_layers: int = 2
    determinants: int = 1


@dc.dataclass
class Network:  # config.py:100-104
    type: str = "psiformer"  # "psiformer" | "laughlin" (analytic ground state, networks/laughlin.py)
    orbital: str = "full"  # OrbitalType, config.py:87-89: "full" | "sparse"
    psiformer: PsiformerNetwork = dc.field(default_factory=PsiformerNetwork)


@dc.dataclass
class MCMC:  # config.py:107-122
    steps: int = 10
    width: float = 0.1
    burn_in: int = 200
    adapt_frequency: int = 100


@dc.dataclass
class LearningRate:  # config.py:125-137
    rate: float = 0.005
    decay: float = 1.0
    delay: float = 2000.0

    def schedule(self, t):
        return self.rate * (1.0 / (1.0 + (t / self.delay))) ** self.decay


@dc.dataclass
class OptimizerAdam:
    lr: LearningRate = dc.field(default_factory=LearningRate)


@dc.dataclass
class OptimizerKfac:
    lr: LearningRate = dc.field(default_factory=lambda: LearningRate(rate=0.05))


@dc.dataclass
class Optim:  # config.py:156-161
    iterations: int = 1000
    optimizer: str | None = "kfac"  # "adam" | "kfac" | "none"
    adam: OptimizerAdam = dc.field(default_factory=OptimizerAdam)
    kfac: OptimizerKfac = dc.field(default_factory=OptimizerKfac)


@dc.dataclass
class Log:  # config.py:164-198 (not interpreted by the hot path)
    save_path: str | None = None
    restore_path: str | None = None
    save_time_interval: int = 600
    save_step_interval: int = 1000
    initial_energy: bool = True


@dc.dataclass
class Config:  # config.py:201-214
    batch_size: int = 3360
    seed: int = dc.field(default_factory=lambda: int(time.time()))
    system: System = dc.field(default_factory=System)
    network: Network = dc.field(default_factory=Network)
    mcmc: MCMC = dc.field(default_factory=MCMC)
    optim: Optim = dc.field(default_factory=Optim)
    log: Log = dc.field(default_factory=Log)

    @classmethod
    def from_dict(cls, dikt: dict) -> "Config":
        return from_dict(cls, dikt)


def from_dict(cls, dikt: dict[str, Any]):
    """Nested dict -> dataclass; unknown keys are ignored (config.py:23-48 tolerates them)."""
    try:
        kwargs = {}
        for f in dc.fields(cls):
            if f.name not in dikt:
                continue
            v = dikt[f.name]
            sub = _DATACLASS_FIELDS.get((cls.__name__, f.name))
            if sub is not None and isinstance(v, dict):
                v = from_dict(sub, v)
            elif f.name == "nspins":
                v = tuple(int(s) for s in v)
            kwargs[f.name] = v
        return cls(**kwargs)
    except Exception as e:  # same error type as the reference
        raise ValueError(f"Error converting dictionary to {cls.__name__}: {e}")


_DATACLASS_FIELDS = {
    ("Config", "system"): System,
    ("Config", "network"): Network,
    ("Config", "mcmc"): MCMC,
    ("Config", "optim"): Optim,
    ("Config", "log"): Log,
    ("Network", "psiformer"): PsiformerNetwork,
    ("Optim", "adam"): OptimizerAdam,
    ("Optim", "kfac"): OptimizerKfac,
    ("OptimizerAdam", "lr"): LearningRate,
    ("OptimizerKfac", "lr"): LearningRate,
}
