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


# ---- sliced int8 Gram that forms M = Xc Xc^T in tp_pca (tp_igram_sliced) ------------------------------------------
def _centred_corr_like(n, seed):
    """a matrix shaped like the centred correlation matrix: unit-ish diagonal, small off-diagonal entries of both signs,
    rows of very different magnitude (one zero row, one tiny row, one huge row)"""
    rng = np.random.default_rng(seed)
    a = rng.normal(0.0, 0.03, size=(n, n))
    a = 0.5 * (a + a.T) + np.eye(n) * 0.9
    a -= a.mean(axis=0, keepdims=True)
    a[3, :] = 0.0
    a[5, :] *= 1e-9
    a[7, :] *= 1e6
    return a


def _exact_gram(a):
    al = a.astype(np.longdouble)
    return (al @ al.T)


@pytest.mark.parametrize("n", [64, 200, 257, 1000, 1403])
def test_sliced_gram_fp64_level(ctx, n):
    a = _centred_corr_like(n, seed=n)
    g = ctx.test_mgram(a)
    ref = _exact_gram(a)
    # digit pairs left out are below 2^-56 of (row scale x row scale); per element the bound is n * 8 * 2^-56 * s_i s_j
    s = np.maximum(np.abs(a).max(axis=1), 1e-300)
    s = 2.0 ** np.ceil(np.log2(s))
    bound = n * 16 * 2.0 ** -56 * np.outer(s, s) + 4 * np.finfo(float).eps * np.abs(ref).astype(float)
    err = np.abs(g - ref).astype(float)
    assert (err <= bound).all(), float((err / bound).max())
    assert np.array_equal(g, g.T)                     # mirrored elements are the same bits
    # relative to the diagonal (what the eigenvalues of M see): a few ulp -- the dropped scale-8 digit pairs are a
    # systematic +n * E[d^2] * 2^-68 s_i^2 on the diagonal, random elsewhere
    d = np.sqrt(np.abs(np.diag(ref).astype(float)))
    rel = err / np.maximum(np.outer(d, d), 1e-300)
    assert rel[d > 0][:, d > 0].max() <= 256 * np.finfo(float).eps, rel[d > 0][:, d > 0].max()


def test_sliced_gram_row_block_equals_symmetric_launch(ctx):
    n = 700
    a = _centred_corr_like(n, seed=5)
    full = ctx.test_mgram(a)
    blk = ctx.test_mgram(a, 256, 512, fill=-7.0)
    assert np.array_equal(blk[256:512], full[256:512])        # same bits from the full-width row block of a rank
    assert (blk[:256] == -7.0).all() and (blk[512:] == -7.0).all()


def test_pca_scores_same_with_sliced_and_fp64_gram(ctx):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(1500, seed=3)
    bad, _, _ = ctx.filter(m)
    keep = np.flatnonzero(~bad).astype(np.int32)
    out = {}
    try:
        for name, v in (("fp64", 0), ("sliced", 1024)):
            ctx.set("mgram_min_n", v)
            ctx.compact(keep); ctx.correlation()
            k = ctx.pca(200)
            out[name] = ctx.get_scores(keep.size, k)
    finally:
        ctx.set("mgram_min_n", 1024)
    a, b = out["fp64"], out["sliced"]
    sgn = np.sign((a * b).sum(axis=0)); sgn[sgn == 0] = 1
    assert np.abs(a - b * sgn).max() <= 1e-9 * np.abs(a).max()


# ---- symmetric products of a call spread over several GPUs (SymShard), emulated rank by rank on one GPU ---------------
@pytest.mark.parametrize("n,nranks", [(300, 2), (513, 3), (700, 4), (1000, 8), (1403, 5), (129, 2)])
def test_symshard_integer_gram_equals_one_gpu(ctx, n, nranks):
    a = _sym_counts(n, 30, seed=n + nranks)
    ref = a @ a.T
    g = ctx.test_symshard(a.astype(np.float64), nranks, kind=0)
    assert np.array_equal(g.astype(np.int64), ref)          # nothing left at `fill`, every element exact


@pytest.mark.parametrize("n,nranks", [(300, 2), (700, 4), (1000, 8), (1403, 3)])
def test_symshard_sliced_gram_equals_one_gpu(ctx, n, nranks):
    a = _centred_corr_like(n, seed=n)
    one = ctx.test_mgram(a)
    g = ctx.test_symshard(a, nranks, kind=1)
    assert np.array_equal(g, one)                            # same bits whoever computed the block pair
