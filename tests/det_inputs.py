"""Deterministic synthetic inputs built from integer hashing and exact IEEE +,-,* only
(no libm, no library RNG), so the golden generator, the CPU tests and the GPU tests
regenerate bit-identical tensors on any host instead of committing megabytes of floats."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(idx, seed):
    with np.errstate(over="ignore"):
        z = (idx.astype(np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) & _M
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def uniform01(n, seed):
    """float32 in [0,1) on a 2^-24 grid (exact)."""
    bits = _splitmix(np.arange(n, dtype=np.uint64), seed) >> np.uint64(40)
    return (bits.astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def bell(n, seed, scale=1.0):
    """Signed, dense, bell-shaped (sum of 6 uniforms, centred): what conv outputs look like."""
    acc = np.zeros(n, dtype=np.float32)
    for k in range(6):
        acc = (acc + uniform01(n, seed * 16 + k)).astype(np.float32)
    return ((acc - np.float32(3.0)) * np.float32(scale)).astype(np.float32)


def relu_bell(n, seed, scale=1.0):
    """~50 % exact zeros (post-ReLU / 'image'-like)."""
    x = bell(n, seed, scale)
    return np.where(x > 0, x, np.float32(0.0)).astype(np.float32)


def heavy_tail(n, seed, outlier=1.0e4):
    """Product of three centred uniforms (sharp peak at 0) plus one large outlier:
    almost everything lands in the lowest bins."""
    a = uniform01(n, seed * 16 + 1) - np.float32(0.5)
    b = uniform01(n, seed * 16 + 2) - np.float32(0.5)
    c = uniform01(n, seed * 16 + 3) - np.float32(0.5)
    x = (a * b * c * np.float32(8.0)).astype(np.float32)
    if n:
        x[n // 3] = np.float32(outlier)
    return x


def constant(n, value=1.0):
    return np.full(n, value, dtype=np.float32)


def int_grid(n, seed, lo=-200, hi=200, denom=16.0):
    """Values on a k/denom grid: exercises exact .5 ties of the half-even / half-away rounding."""
    u = _splitmix(np.arange(n, dtype=np.uint64), seed) % np.uint64(hi - lo + 1)
    return ((u.astype(np.int64) + lo).astype(np.float32) / np.float32(denom)).astype(np.float32)


def stats_cases():
    """name -> list of per-batch float32 vectors (ragged sizes, zeros, ties, outliers)."""
    return {
        "bell_dense": [bell(40000, 1), bell(30011, 2, 1.7)],
        "relu_half_zero": [relu_bell(25000, 3), relu_bell(25000, 4, 0.5)],
        "constant_one": [constant(5000, 1.0)],
        "heavy_tail_outlier": [heavy_tail(30000, 5)],
        "all_zero": [np.zeros(1000, dtype=np.float32)],
        "single_elem": [np.array([-3.25], dtype=np.float32)],
        "tiny_values": [bell(4097, 6, 1e-7)],
        "neg_dominant": [-np.abs(bell(12345, 7)) - np.float32(0.25), bell(77, 8, 0.1)],
        "second_batch_larger_than_first": [bell(8000, 9, 0.5), bell(8000, 10, 2.0)],
        "grid_ties": [int_grid(20000, 11)],
        "ragged_3": [bell(3, 12)],
    }
