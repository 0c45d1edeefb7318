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


def monopole_harmonic(q, l, m, electrons):  # noqa: E741
    # one_rdm.py:34-58 (make_monopole_harm), fp64
    from scipy import special as ss

    norm_factor = np.sqrt(((2 * l + 1) / (4 * np.pi)) * (ss.factorial(l - m) * ss.factorial(l + m))
                          / (ss.factorial(l - q) * ss.factorial(l + q)))
    s = np.arange(l - m + 1)
    sum_factors = (-1) ** (l - m - s) * ss.comb(l - q, s) * ss.comb(l + q, l - m - s)
    electrons = np.asarray(electrons, dtype=np.float64)
    theta, phi = electrons[..., 0], electrons[..., 1]
    x = np.clip(np.cos(theta), -1 + 1e-4, 1 - 1e-4)
    theta_part = np.sum(sum_factors * (1 - x[..., None]) ** (l - s - (m + q) / 2) * (1 + x[..., None]) ** (s + (m + q) / 2), axis=-1)
    return norm_factor / 2**l * theta_part * np.exp(1j * m * phi)


def lll_orbitals(flux, electrons):
    # one_rdm.py:71-72: Y_{Q,Q,m}, m = -Q .. Q, stacked on the last axis
    Q = flux / 2
    return np.stack([monopole_harmonic(Q, Q, m, electrons) for m in np.arange(-Q, Q + 1)], axis=-1)


def one_rdm_data_prime(data, r_prime):
    # one_rdm.py:92-94: N copies of the walker, copy a with electron a moved to r'
    data = np.asarray(data)
    nelec = data.shape[-2]
    dp = np.repeat(data[..., None, :, :], nelec, axis=-3).copy()
    idx = np.arange(nelec)
    dp[..., idx, idx, :] = np.asarray(r_prime)[..., None, :]
    return dp


def one_rdm_product(flux, data, r_prime, logpsi, logpsi_prime):
    # one_rdm.py:96-109 for a batch: data (B,N,2), r_prime (B,2), logpsi (B), logpsi_prime (B,N) -> (B,L,L)
    varphi = lll_orbitals(flux, data)                         # (B,N,L)
    varphi_prime = lll_orbitals(flux, np.asarray(r_prime)[:, None, :])  # (B,1,L)
    wf_ratio = np.exp(np.asarray(logpsi_prime, dtype=np.complex128) - np.asarray(logpsi, dtype=np.complex128)[:, None])
    return (4 * np.pi) * np.sum(wf_ratio[..., None, None] * varphi[..., None] * np.conj(varphi_prime)[..., None, :], axis=1)
