"""Seeded synthetic inputs shared by the parity tests and bench.py."""
import numpy as np


def poisson_stream(n, K, rate, seed):
    """Poisson surrogate: inter-arrivals Exp(rate), nodes uniform on 1..K.  Window statistics match a
    stationary weakly-excited Hawkes stream of the same total rate."""
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.exponential(1.0 / rate, n))
    nodes = rng.integers(1, K + 1, n, dtype=np.int64)
    return t, nodes, float(t[-1] * (1.0 + 1e-9) if n else 1.0)


def ln_params(K, seed, wmax=None, density=None):
    """cfg2-style LogitNormal parameters: lambda0 = 1, W ~ U(0, wmax), mu ~ U(-1,1), tau ~ U(0.5,2)."""
    rng = np.random.default_rng(seed)
    wmax = 1.0 / K if wmax is None else wmax
    lam0 = np.ones(K)
    W = rng.uniform(0.0, wmax, (K, K))
    mu = rng.uniform(-1.0, 1.0, (K, K))
    tau = rng.uniform(0.5, 2.0, (K, K))
    A = None if density is None else (rng.random((K, K)) < density).astype(np.float64)
    return lam0, W, mu, tau, A


def exp_params(K, seed, wmax=None, density=None):
    rng = np.random.default_rng(seed)
    wmax = 1.0 / K if wmax is None else wmax
    lam0 = rng.uniform(0.5, 1.5, K)
    W = rng.uniform(0.0, wmax, (K, K))
    theta = rng.uniform(0.5, 2.0, (K, K))
    A = None if density is None else (rng.random((K, K)) < density).astype(np.float64)
    return lam0, W, theta, A


def philox_uniform(seed, index, counter):
    """numpy restatement of philox_uniform() in csrc/nhp_internal.cuh (Philox4x32-10), vectorised over index."""
    idx = np.asarray(index, dtype=np.uint64)
    c0 = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (idx >> np.uint64(32)).astype(np.uint64)
    c2 = np.full_like(c0, np.uint64(counter & 0xFFFFFFFF))
    c3 = np.full_like(c0, np.uint64((counter >> 32) & 0xFFFFFFFF))
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    x = ((c0 << np.uint64(32)) | c1) >> np.uint64(11)
    return x.astype(np.float64) * 2.0 ** -53
