"""GPU tests at sizes the CPU oracle cannot reach in seconds (BASELINE.json configs 2-5 in spirit): the CUDA path
is checked through size-independent properties of the reference's definitions instead of a full oracle run.

  * correlation: symmetric, unit diagonal, equal to numpy's corrcoef on sampled rows            (R/TADpole.R:94-100)
  * prcomp scores: columns are eigenvectors of Xc Xc^T scaled by sqrt(lambda), descending        (R/TADpole.R:366-367)
  * CONISS: the last merge height is the total sum of squares of the clustered columns           (rioja::chclust)
  * Calinski-Harabasz: recomputed in numpy from the returned scores and the cut of the returned dendrogram
                                                                                                  (fpc::calinhara, :115-120)
  * TAD tables: contiguous, disjoint, cover exactly the good bins                                 (R/TADpole.R:470-497)
  * diffT: zero against itself, symmetric in its arguments, cumulative, ends at 1                 (R/DiffT.R:41-49)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_LARGE = 5000


@pytest.fixture(scope="module")
def large(ctx):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(N_LARGE, seed=3)
    bad, _, _ = ctx.filter(m)
    keep = np.flatnonzero(~bad).astype(np.int32)
    ctx.compact(keep)
    x = ctx.get_filtered(keep.size)
    ctx.correlation()
    cor = ctx.get_correlation(keep.size)
    k = ctx.pca(200)
    scores = ctx.get_scores(keep.size, k)
    res = ctx.call(m)
    return dict(m=m, keep=keep, x=x, cor=cor, k=k, scores=scores, res=res)


def test_correlation_properties(large):
    cor, x = large["cor"], large["x"]
    assert np.array_equal(cor, cor.T)
    np.testing.assert_allclose(np.diag(cor), 1.0, atol=1e-12)
    rows = np.random.default_rng(0).choice(cor.shape[0], 40, replace=False)
    ref = np.corrcoef(x, rowvar=False)[rows]          # columns are the variables, as in sparse_cor
    np.testing.assert_allclose(cor[rows], ref, atol=1e-12, rtol=0)


def test_scores_are_scaled_eigenvectors(large):
    cor, s = large["cor"], large["scores"]
    xc = cor - cor.mean(axis=0, keepdims=True)          # prcomp centres the columns
    lam = (s * s).sum(axis=0)                           # |u sqrt(lambda)|^2 = lambda
    assert np.all(np.diff(lam) <= 1e-9 * lam[0])        # descending
    gram = s.T @ s
    off = gram - np.diag(np.diag(gram))
    assert np.abs(off).max() <= 1e-9 * lam[0]           # orthogonal columns
    ms = xc @ (xc.T @ s)                                # M s_j = lambda_j s_j
    resid = np.linalg.norm(ms - s * lam, axis=0) / (lam[0] * np.sqrt(lam))
    assert resid.max() <= 1e-10


def test_dendrogram_height_and_ch_scores(large):
    from tadpole_b200.hclust import cutree
    res, s = large["res"], large["scores"]
    i = res["n_pcs"]
    pcs = s[:, :i]
    tss = ((pcs - pcs.mean(axis=0)) ** 2).sum()
    assert res["seqdist"].max() == pytest.approx(tss, rel=1e-9)       # CONISS: last height = total SS
    row = res["scores"][i - 1]
    levels = np.flatnonzero(~np.isnan(row)) + 1
    n = s.shape[0]
    tot = ((s - s.mean(axis=0)) ** 2).sum()             # CH uses ALL k columns (quirk Q2)
    for ncl in (levels[0], levels[len(levels) // 2], levels[-1]):
        lab = cutree(res["seqdist"], int(ncl))
        w = sum(((s[lab == c] - s[lab == c].mean(axis=0)) ** 2).sum() for c in np.unique(lab))
        ch = (n - ncl) * (tot - w) / ((ncl - 1) * w)
        assert row[ncl - 1] == pytest.approx(ch, rel=1e-8)


def test_tad_tables_partition_the_good_bins(ctx, large):
    from tadpole_b200 import TADpole, api
    api.QUIET = True
    tp = TADpole(large["m"], ctx=ctx)
    bad = np.ones(N_LARGE, bool); bad[large["keep"]] = False
    for key, tab in tp.clusters.items():
        assert np.all(tab[:, 0] <= tab[:, 1]) and np.all(tab[1:, 0] > tab[:-1, 1])
        covered = np.zeros(N_LARGE, bool)
        for a, b in tab:
            covered[a - 1: b] = True
        # a TAD may absorb interior bad bins (fix_values), but never leaves a good bin out
        assert covered[~bad].all()
        assert tab.shape[0] <= int(key)


def test_centromere_arms_at_scale(ctx):
    from tadpole_b200 import TADpole, api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    n = 6000
    tp = TADpole(synth_hic(n, seed=5, centromere=True), centromere_search=True, ctx=ctx)
    assert tp.p.n_pcs >= 1 and tp.q.n_pcs >= 1
    ma = tp.merging_arms
    # (the q arm's table can run past the matrix: the reference re-inserts q-arm bad columns it never removed,
    #  SURVEY.md quirk Q3, replicated)
    assert np.all(ma[:, 0] <= ma[:, 1]) and np.all(ma[1:, 0] > ma[:-1, 1])
    # the arms do not share a TAD: the centromere gap separates the two tables
    assert ma.shape[0] == tp.p.cluster[str(tp.p.optimal_n_clusters)].shape[0] + tp.q.cluster[str(tp.q.optimal_n_clusters)].shape[0]


def test_difft_properties_full_size(ctx):
    from tadpole_b200.synth import synth_partition_pairs
    lx, ly = synth_partition_pairs(64, 15000, 500, seed=2)        # configs[4] shape, 64 of the 1000 pairs
    dxy = ctx.difft_batch(lx, ly)
    dyx = ctx.difft_batch(ly, lx)
    assert np.array_equal(dxy, dyx)
    assert np.all(np.diff(dxy, axis=1) >= 0) and np.all(dxy[:, -1] == 1.0)
    clean = np.maximum(lx, 1)                                     # no uncovered bins: a partition equals itself
    assert not ctx.difft_batch(clean, clean).any()
