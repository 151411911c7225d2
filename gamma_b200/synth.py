"""Seeded synthetic inputs for the hot path.

Base vectors are a mixture of Gaussians (so IVF lists have structure); queries
are fresh samples of the same mixture; the scalar-filter field is U{0..99}.
Used by tests/, bench.py and the golden-fixture generator — never by the
search path itself.
"""
import numpy as np

SEED_BASE, SEED_QUERY, SEED_FILTER = 20240601, 20240602, 20240603


def mixture(n, d, seed, n_clusters=4096, spread=0.3, normalize=False, centres_seed=SEED_BASE, chunk=1 << 20):
    """n samples of a `n_clusters`-component isotropic Gaussian mixture in R^d (float32)."""
    crng = np.random.default_rng(centres_seed + 7919)
    centres = crng.standard_normal((n_clusters, d), dtype=np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, d), np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        a = rng.integers(0, n_clusters, size=e - s)
        out[s:e] = centres[a] + spread * rng.standard_normal((e - s, d), dtype=np.float32)
    if normalize:
        out /= np.linalg.norm(out, axis=1, keepdims=True)
    return out


def base_vectors(n, d, normalize=False, n_clusters=4096):
    return mixture(n, d, SEED_BASE, n_clusters=n_clusters, normalize=normalize)


def query_vectors(n, d, normalize=False, n_clusters=4096):
    return mixture(n, d, SEED_QUERY, n_clusters=n_clusters, normalize=normalize)


def filter_field(n, lo=0, hi=100):
    """integer scalar field ~ U{lo..hi-1}; the range filter keeps value < 30 (~30 % pass)."""
    return np.random.default_rng(SEED_FILTER).integers(lo, hi, size=n).astype(np.int32)


def deleted_docs(n, frac=0.01):
    rng = np.random.default_rng(SEED_FILTER + 1)
    return np.sort(rng.choice(n, size=int(n * frac), replace=False)).astype(np.int64)
