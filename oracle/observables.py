"""numpy restatement of deephall/netobs_bridge/observables/{pair_corr,density,overlap}.py.

TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu_baseline): never imported by the product path.
Pinned by closed forms (tests/test_oracle_known_answers.py): independent uniform points give the flat
g = (N - 1) / N; a wavefunction's overlap with itself is 1.
"""
import numpy as np


def pair_correlation_increment(data, bins=200):
    # pair_corr.py:47-59 (fp64 here; the reference runs jnp in fp32)
    data = np.asarray(data, dtype=np.float64).reshape(-1, *np.shape(data)[-2:])
    batch_size, nelec, _ = data.shape
    theta, phi = data[..., 0], data[..., 1]
    xyz = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=-1)
    cos12 = np.sum(xyz[..., :, None, :] * xyz[..., None, :, :], axis=-1)
    iu = np.triu_indices(nelec, 1)
    theta12 = np.arccos(np.clip(cos12[:, iu[0], iu[1]].reshape(-1), -1.0, 1.0))
    to_add, _ = np.histogram(theta12, bins, (0, np.pi), weights=1 / np.sin(theta12))
    return to_add * 4 * bins / batch_size / nelec**2 / np.pi


def density_increment(data, bins=50):
    # density.py:46-48
    theta = np.asarray(data, dtype=np.float64)[..., 0].reshape(-1)
    return np.histogram(theta, bins, (0.0, np.pi))[0]


def overlap_evaluate(logphi, logpsi):
    # overlap.py:57-63
    logphi, logpsi = np.asarray(logphi, dtype=np.complex128), np.asarray(logpsi, dtype=np.complex128)
    shift = np.mean(logphi - logpsi)
    ratio = np.exp(logphi - logpsi - shift)
    return {"ratio": ratio, "ratio_square": np.abs(ratio) ** 2}


def overlap_digest(ratio, ratio_square):
    # overlap.py:65-70
    return np.abs(np.nanmean(ratio)) ** 2 / np.nanmean(ratio_square)
