"""GPU tests of the tcgen05 int8 Gram kernel (csrc/igemm.cu) behind stage 2: for symmetric matrices of integer
counts the Gram matrix X X^T must be EXACT (compared with numpy int64), whatever the size and padding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sym_counts(n, scale, seed, signed=False):
    rng = np.random.default_rng(seed)
    a = rng.poisson(scale, size=(n, n)).astype(np.int64)
    if signed:
        a -= int(scale)
    a = np.triu(a) + np.triu(a, 1).T
    return a


@pytest.mark.parametrize("n,scale", [(64, 3), (128, 50), (200, 7), (257, 1000), (1000, 20), (1531, 300)])
def test_gram_exact(ctx, n, scale):
    a = _sym_counts(n, scale, seed=n)
    g = ctx.test_igram(a.astype(np.float64))
    assert g is not None
    ref = a @ a.T
    assert np.array_equal(g.astype(np.int64), ref)


def test_gram_exact_large_and_negative_counts(ctx):
    n = 384
    a = _sym_counts(n, 40, seed=9, signed=True)
    a[5, 7] = a[7, 5] = 1048575                  # largest admissible count (2^20 - 1)
    a[11, 11] = -1048575
    g = ctx.test_igram(a.astype(np.float64))
    assert g is not None
    assert np.array_equal(g.astype(np.int64), a @ a.T)


def test_non_integer_input_is_declined(ctx):
    a = _sym_counts(300, 10, seed=1).astype(np.float64)
    a[3, 4] = a[4, 3] = 2.5
    assert ctx.test_igram(a) is None
    a[3, 4] = a[4, 3] = 2.0 ** 20                # too large for three digits
    assert ctx.test_igram(a) is None
