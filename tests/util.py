"""Shared helpers for the tests: the SURVEY.md 8(d) synthetic generator on the CPU."""
import numpy as np


def synth_params(G, C, seed=20261019, m0=-0.675):
    rng = np.random.default_rng(seed)
    beta = rng.lognormal(0.0, 0.4, C)
    beta = np.clip(beta / beta.mean() * 0.06, 0.005, 0.5)
    size = rng.lognormal(0.0, 0.75, G)
    mu = rng.lognormal(m0, 1.8, G)
    return mu, size, beta


def synth_counts(G, C, seed=20261019, m0=-0.675):
    """x_ij ~ NB(size phi_i, mean mu_i beta_j) as a dense float64 (G, C) array."""
    mu, size, beta = synth_params(G, C, seed, m0)
    rng = np.random.default_rng(seed + 1)
    lam = rng.gamma(size[:, None], (mu[:, None] * beta[None, :]) / size[:, None])
    X = rng.poisson(lam).astype(np.float64)
    return X, mu, size, beta


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    s = np.maximum(np.abs(b), 1e-300)
    return np.where(d == 0, 0.0, d / s)


# Philox4x32-10 in plain Python (independent of the CUDA implementation) -------------
def philox4x32_10(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF,
             p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


PHILOX_KAT = [  # Random123 kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
