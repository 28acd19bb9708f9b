"""Integer-only synthetic Hi-C-like count matrix for fixtures that must be reproduced bit for bit on another machine
without storing the matrix: every step is uint64 / int64 arithmetic (no floating point, no library RNG), so numpy's SIMD
dispatch, libm and the CPU model cannot change a single count.  Power-law-ish decay 240 // (|i-j| + 1), two nested block
levels (x3 inside a 37-bin block, x2 more inside an 11-bin block), hash noise, a few bins zeroed entirely."""
import numpy as np


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def golden_int_matrix(n, seed):
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.int64)
        lo = np.minimum(i[:, None], i[None, :]).astype(np.uint64)
        hi = np.maximum(i[:, None], i[None, :]).astype(np.uint64)
        d = (hi - lo).astype(np.int64)
        lam = 240 // (d + 1)
        lam = np.where((i[:, None] // 37) == (i[None, :] // 37), lam * 3, lam)
        lam = np.where((i[:, None] // 11) == (i[None, :] // 11), lam * 2, lam)
        h = _splitmix64(lo * np.uint64(1000003) + hi + np.uint64(seed) * np.uint64(0x51ED270B))      # symmetric in (i, j)
        noise = (h % (lam // 3 + 2).astype(np.uint64)).astype(np.int64)
        m = lam + noise
        dead = (_splitmix64(i.astype(np.uint64) + np.uint64(seed) * np.uint64(7919)) % np.uint64(173)) == 0
        m[dead, :] = 0
        m[:, dead] = 0
    return m.astype(np.float64)
