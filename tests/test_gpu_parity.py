"""GPU parity tests: every stage of the CUDA path (through the C ABI) against the CPU oracle.

Bar: bit-exact for integer / index work (bad flags, merge order, n_cluster, optimal n_pcs and
level, TAD boundaries, diffT); stated tolerances for floating point (written next to each check).
"""
import json
import sys
import os

import numpy as np
import pytest

from oracle import tadpole_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- stage 1 -------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,order", [(64, "C"), (200, "C"), (200, "F"), (601, "C"), (2000, "F")])
def test_filter_flags_and_compaction(ctx, n, order):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(n, seed=n)
    m[3, 7] = np.nan                                  # NA -> 0 (R/TADpole.R:19)
    m[np.tril_indices(n, -1)] = -12345.0              # lower triangle must be ignored (forceSymmetric uplo='U')
    mm = np.asfortranarray(m) if order == "F" else m
    bad, rm, thr = ctx.filter(mm, bad_frac=0.01)
    sym = O.symmetrise_upper(m)
    obad, orm, othr = O.bad_columns(sym, 0.01)
    assert (bad == obad).all()
    # counts are integers, sums are exact in double: row means agree to the last bit or two
    np.testing.assert_allclose(rm, orm, rtol=4e-16, atol=0)
    assert thr == pytest.approx(othr, rel=1e-15)
    keep = np.flatnonzero(~bad)
    ctx.compact(keep)
    x = ctx.get_filtered(keep.size)
    assert (x == sym[np.ix_(keep, keep)]).all()


def test_filter_bad_frac_zero_and_float_data(ctx):
    rng = np.random.default_rng(5)
    m = rng.random((300, 300)) * 10
    m[17, :] = 0; m[:, 17] = 0
    bad, rm, thr = ctx.filter(m, bad_frac=0.0)
    assert bad.sum() == 1 and bad[17] and np.isnan(thr)
    bad, rm, thr = ctx.filter(m, bad_frac=0.05)
    sym = O.symmetrise_upper(m)
    obad, orm, othr = O.bad_columns(sym, 0.05)
    np.testing.assert_allclose(rm, orm, rtol=1e-14)
    assert (bad == obad).all()


# ---- stage 2 -------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [200, 601, 2000])
def test_correlation(ctx, synth_cache, n):
    c = synth_cache(n)
    ctx.set_filtered(c["lm"].mat)
    ctx.correlation()
    cor = ctx.get_correlation(c["lm"].mat.shape[0])
    # FP64 throughout; the one-pass covariance formula cancels (quirk Q8), so compare absolutely
    # on a quantity bounded by 1: |cor_gpu - cor_oracle| <= 1e-12
    assert np.abs(cor - c["cor"]).max() <= 1e-12
    assert (cor == cor.T).all()


def test_correlation_zero_variance_column(ctx):
    rng = np.random.default_rng(2)
    x = rng.poisson(5.0, (64, 64)).astype(float)
    x = np.triu(x) + np.triu(x, 1).T
    x[:, 10] = 3.0; x[10, :] = 3.0                    # constant column: sd = 0 -> NaN -> 0 (quirk Q9)
    ctx.set_filtered(x)
    ctx.correlation()
    cor = ctx.get_correlation(64)
    ref = O.sparse_cor(x)
    finite = np.isfinite(ref) & np.isfinite(cor)
    assert (np.isfinite(ref) == np.isfinite(cor)).all()
    assert np.abs(cor[finite] - ref[finite]).max() <= 1e-12


# ---- stage 3 -------------------------------------------------------------------------------------
def _align(scores, ref):
    sgn = np.sign((scores * ref).sum(axis=0))
    sgn[sgn == 0] = 1
    return scores * sgn


@pytest.mark.parametrize("n", [200, 601, 2000])
def test_pca_scores(ctx, synth_cache, n):
    c = synth_cache(n)
    nf = c["lm"].mat.shape[0]
    ctx.set_correlation(c["cor"])
    k = ctx.pca(200)
    assert k == c["k"]
    sc = _align(ctx.get_scores(nf, k), c["pcs"])
    scale = np.abs(c["pcs"]).max()
    # PCs agree up to sign; tolerance 1e-9 of the largest score (FP64, north_star).  The last
    # component of a full-rank request (k = nf) is numerical noise in both implementations.
    kk = k - 1 if k == nf else k
    assert np.abs(sc[:, :kk] - c["pcs"][:, :kk]).max() <= 1e-9 * scale
    # sign-invariant check on what the clustering consumes: pairwise squared distances
    idx = np.random.default_rng(0).integers(0, nf, (200, 2))
    for i in (1, 5, 50, kk):
        dg = ((sc[idx[:, 0], :i] - sc[idx[:, 1], :i]) ** 2).sum(1)
        do = ((c["pcs"][idx[:, 0], :i] - c["pcs"][idx[:, 1], :i]) ** 2).sum(1)
        np.testing.assert_allclose(dg, do, rtol=1e-9, atol=1e-12 * scale ** 2)


# ---- stages 4 + 5 --------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [200, 601])
def test_sweep_against_oracle_same_scores(ctx, synth_cache, n):
    """Feed the ORACLE's PC scores to the GPU sweep: merge order, seqdist, n_cluster and CH must
    match the Lance-Williams restatement candidate by candidate."""
    c = synth_cache(n)
    pcs, k = c["pcs"], c["k"]
    nf = pcs.shape[0]
    ctx.set_scores(pcs)
    ncl, sc = ctx.sweep(k, min_clusters=2)
    cands = sorted(set([1, 2, 3, 7, 20, 50, 100, k]))
    for i in cands:
        seq, order = ctx.dendro(i - 1, nf)
        oseq, oorder = O.coniss_lw(pcs[:, :i])
        assert (order == oorder).all(), f"merge order differs for candidate {i}"
        np.testing.assert_allclose(seq, oseq, rtol=1e-11)
        oscore, oncl, _ = O.candidate_scores(pcs, i, 2)
        assert ncl[i - 1] == oncl
        got = sc[i - 1, :oncl]
        assert (np.isnan(got) == np.isnan(oscore)).all()
        m = ~np.isnan(oscore)
        np.testing.assert_allclose(got[m], oscore[m], rtol=1e-9)     # CH tolerance 1e-9 (FP64)
        assert np.isnan(sc[i - 1, oncl:]).all()


def test_sweep_bitmap_neighbours(ctx, synth_cache, monkeypatch):
    """Above ~18k bins the live-boundary links are one bit per boundary in shared memory instead of two link arrays; the
    same code path forced at 601 bins must give the array path's result bit for bit, and the oracle's merge order."""
    c = synth_cache(601)
    pcs, k = c["pcs"], c["k"]
    nf = pcs.shape[0]
    ctx.set_scores(pcs)
    ncl_a, sc_a = ctx.sweep(k)
    ref = {i: ctx.dendro(i - 1, nf) for i in (1, 9, 77, k)}
    monkeypatch.setenv("TADPOLE_SWEEP_LINKS", "bitmap")
    ctx.set_scores(pcs)
    ncl_b, sc_b = ctx.sweep(k)
    assert np.array_equal(ncl_a, ncl_b) and np.array_equal(sc_a, sc_b, equal_nan=True)
    for i, (seq, order) in ref.items():
        seq_b, order_b = ctx.dendro(i - 1, nf)
        assert np.array_equal(seq, seq_b) and np.array_equal(order, order_b)
        assert np.array_equal(order_b, O.coniss_lw(pcs[:, :i])[1])


def test_sweep_candidate_sharding(ctx, synth_cache):
    c = synth_cache(200)
    pcs, k = c["pcs"], c["k"]
    ctx.set_scores(pcs)
    ncl_all, sc_all = ctx.sweep(k)
    ncl0, sc0 = ctx.sweep(k, cand_begin=0, cand_stride=2)
    ncl1, sc1 = ctx.sweep(k, cand_begin=1, cand_stride=2)
    assert (ncl0[0::2] == ncl_all[0::2]).all() and (ncl0[1::2] == 0).all()
    assert (ncl1[1::2] == ncl_all[1::2]).all()
    w = sc_all.shape[1]
    a = np.full_like(sc_all, np.nan); a[0::2, :sc0.shape[1]] = sc0[0::2]; a[1::2, :sc1.shape[1]] = sc1[1::2]
    assert np.array_equal(a, sc_all, equal_nan=True)


def test_sweep_ties_lowest_index_first(ctx):
    """Exact ties (duplicate rows) must resolve to the lowest boundary index, as the reference's
    strict '<' scan does."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal((60, 6))
    pcs = np.repeat(base, 2, axis=0)                # pairs of identical rows: 60 exact dSS = 0 ties
    pcs -= pcs.mean(0)
    ctx.set_scores(pcs)
    try:
        ctx.sweep(6)
    except Exception:
        pass                                        # broken stick may find no level; the dendrogram is still there
    for i in (1, 3, 6):
        seq, order = ctx.dendro(i - 1, pcs.shape[0])
        oseq, oorder = O.coniss_lw(pcs[:, :i])
        assert (order[:60] == np.arange(0, 120, 2)).all()      # the 60 zero-cost merges, in index order
        assert (order == oorder).all()


def test_sweep_structural_ties_follow_the_reference(ctx):
    """Bins with IDENTICAL score rows (every zero-variance bin maps to the same row of the correlation matrix, quirk Q9;
    the q arm keeps its all-zero bins, quirk Q3) make exactly tied increases: two single bins on either side of a cluster,
    runs of identical neighbours whose merged cluster ties again with the next one (zero ties that only appear AFTER a
    merge), identical bins far apart.  Lance-Williams arithmetic on the distance matrix (the reference) keeps those ties bit
    for bit and merges the lowest index; the product must produce the same merge order and the same heights."""
    rng = np.random.default_rng(7)
    n, k = 900, 24
    pcs = np.cumsum(rng.standard_normal((n, k)), axis=0) * (1.0 + np.arange(k)) ** -0.7
    v = rng.standard_normal(k) * 3.0                        # the row every "zero-variance" bin gets
    w = rng.standard_normal(k) * 0.5
    dup = [5, 40, 41, 42, 43, 100, 123, 124, 300, 301, 302, 500, 520, 521, 640, 641, 642, 643, 644, 777, 898]
    pcs[dup] = v
    pcs[[200, 230, 231, 260]] = w                          # a second family
    pcs[600:606] = pcs[600]                                 # a run of six identical ordinary rows
    pcs -= pcs.mean(0)
    ctx.set_scores(pcs)
    try:
        ctx.sweep(k)
    except Exception:
        pass                                                # (a candidate without a significant level: the dendrograms exist)
    for i in (1, 2, 7, 24):
        seq, order = ctx.dendro(i - 1, n)
        oseq, oorder = O.coniss_lw(pcs[:, :i])
        assert np.array_equal(order, oorder), f"merge order differs from the reference's at {i} PCs"
        np.testing.assert_allclose(seq, oseq, rtol=1e-9, atol=1e-18 * oseq.max())


def test_zero_variance_bins_tie_family(ctx):
    """The q arm of a centromere_search call keeps its all-zero bins (quirk Q3): their columns have zero variance, the whole
    row of the correlation matrix becomes 0 (0/0 = NaN -> 0, quirk Q9) and in exact arithmetic all of them get the same PC
    scores -- a family of exactly tied increases in every candidate's merge loop.  In floating point the scores of such
    bins agree to the last bits only (in the reference they come out of a BLAS product, here out of the subspace
    iteration), so the reference's own order among them is rounding noise; the product treats increases within 1e-11 as
    tied and merges the lowest index, which is what the reference's scan does on equal values.  Here: the product's scores
    of those bins agree to rounding; with those rows made bitwise equal (the exact-arithmetic picture) the reference's merge
    loop (C oracle) gives the same merge order, level counts and CH rows as the product for EVERY candidate, hence the
    same optimum."""
    from tadpole_b200.synth import synth_hic
    m = synth_hic(1300, seed=9, zero_frac=0.03, centromere=True)
    lm = O.load_mat_numeric(m, centromere_search=True)
    keep = (np.asarray(lm.q.names) - 1).astype(np.int32)
    nf = keep.size
    ctx.filter(m)
    ctx.compact(keep)
    ctx.correlation()
    cor = ctx.get_correlation(nf)
    zero = np.flatnonzero(np.all(cor == 0.0, axis=1))
    assert zero.size >= 8
    k = ctx.pca(200)
    s = ctx.get_scores(nf, k)
    assert np.abs(s[zero] - s[zero[0]]).max() <= 1e-13 * np.abs(s).max(), "zero-variance bins must share one row of scores"
    s[zero] = s[zero[0]]
    ctx.set_scores(s)
    ncl, sc = ctx.sweep(k)
    per = []
    for i in range(1, k + 1):
        score, n_cluster, oseq = O.candidate_scores_c(s, i, 2)
        seq, order = ctx.dendro(i - 1, nf)
        _, oorder = O.coniss_lw(s[:, :i])
        assert np.array_equal(order, oorder), f"merge order differs at {i} PCs"
        assert n_cluster == ncl[i - 1]
        np.testing.assert_allclose(sc[i - 1, :n_cluster][1:], score[1:], rtol=1e-9)
        per.append(score)
    _, opt_pcs, opt_k = O.reduce_scores(per)
    oc, ol = ctx.select(sc)
    assert (oc + 1, ol + 1) == (opt_pcs, opt_k)


def test_large_max_pcs(ctx, synth_cache):
    """prcomp(rank. = min(max_pcs, n)) accepts any max_pcs (R/TADpole.R:366-367).  Here: every max_pcs up to n for matrices
    up to 1024 bins (direct eigensolver), up to 672 above that, with an error that says so beyond."""
    from tadpole_b200 import TadpoleError
    from tadpole_b200.synth import synth_hic
    m = synth_hic(1300, seed=4)
    lm = O.load_mat_numeric(m)
    cor = O.sparse_cor(lm.mat)
    ref = O.prcomp_scores(cor, 600)
    ctx.set_correlation(cor)
    assert ctx.pca(600) == 600
    got = ctx.get_scores(cor.shape[0], 600)
    sgn = np.sign((got * ref).sum(axis=0))
    assert np.abs(got * sgn - ref).max() <= 1e-9 * np.abs(ref).max()
    ctx.set_correlation(cor)
    with pytest.raises(TadpoleError, match="max_pcs <= 672"):
        ctx.pca(1000)
    # at most 1024 bins: all of them
    c = synth_cache(601)
    ctx.set_correlation(c["cor"])
    nf = c["cor"].shape[0]
    assert ctx.pca(5000) == nf
    ref = O.prcomp_scores(c["cor"], nf)
    got = ctx.get_scores(nf, nf)
    keep = np.linalg.norm(ref, axis=0) > 1e-6 * np.linalg.norm(ref[:, 0])      # (the last components are rounding noise)
    sgn = np.sign((got * ref).sum(axis=0))
    assert np.abs(got * sgn - ref)[:, keep].max() <= 1e-9 * np.abs(ref).max()


# ---- full pipeline --------------------------------------------------------------------------------
@pytest.mark.parametrize("n,seed", [(200, 1), (200, 2), (601, 1), (2000, 1)])
def test_tadpole_end_to_end(ctx, n, seed):
    from tadpole_b200 import TADpole
    from tadpole_b200.synth import synth_hic
    m = synth_hic(n, seed=seed)
    tp = TADpole(m, ctx=ctx)
    ref = O.tadpole(m)
    assert tp.n_pcs == ref.n_pcs
    assert tp.optimal_n_clusters == ref.optimal_n_clusters
    assert tp.scores.shape == ref.scores.shape
    assert (np.isnan(tp.scores) == np.isnan(ref.scores)).all()
    msk = ~np.isnan(ref.scores)
    np.testing.assert_allclose(tp.scores[msk], ref.scores[msk], rtol=1e-8)
    assert sorted(tp.clusters) == sorted(str(k) for k in ref.clusters)
    for k, tab in ref.clusters.items():
        assert np.array_equal(tp.clusters[str(k)], tab), f"TAD boundaries differ at level {k}"
    # dendrogram of the optimal candidate: same merge order
    assert (np.argsort(tp.dendro.seqdist, kind="stable") == np.argsort(ref.seqdist, kind="stable")).all()


def test_tadpole_centromere(ctx):
    from tadpole_b200 import TADpole
    from tadpole_b200.synth import synth_hic
    m = synth_hic(700, seed=4, centromere=True)
    tp = TADpole(m, centromere_search=True, ctx=ctx)
    ref = O.tadpole(m, centromere_search=True)
    for arm in ("p", "q"):
        assert tp[arm].n_pcs == ref.arms[arm].n_pcs
        assert tp[arm].optimal_n_clusters == ref.arms[arm].optimal_n_clusters
        for k, tab in ref.arms[arm].clusters.items():
            assert np.array_equal(tp[arm].cluster[str(k)], tab)
    assert np.array_equal(tp.merging_arms, ref.merging_arms)


def test_golden_pipeline(ctx):
    """Committed fixture (tests/golden/make_golden.py): oracle outputs for a 160-bin matrix."""
    from tadpole_b200 import TADpole
    with open(os.path.join(GOLD, "pipeline_n160.json")) as fh:
        g = json.load(fh)
    m = np.array(g["matrix"], dtype=np.float64)
    tp = TADpole(m, ctx=ctx)
    assert tp.n_pcs == g["n_pcs"] and tp.optimal_n_clusters == g["optimal_n_clusters"]
    for k, tab in g["clusters"].items():
        assert np.array_equal(tp.clusters[k], np.array(tab))


# ---- stage 6 -------------------------------------------------------------------------------------
def test_difft_reference_fixture(ctx):
    """The reference's own fixture pair (inst/extdata/control.bed x case.bed); golden committed."""
    from tadpole_b200 import diffT
    with open(os.path.join(GOLD, "difft_control_case.json")) as fh:
        g = json.load(fh)
    out = diffT(np.array(g["control"]), np.array(g["case"]), ctx=ctx)
    assert out.shape == (194,)
    assert (out == np.array(g["normalised"])).all()          # integer arithmetic + one division: bit-exact
    assert round(out[22], 6) == 0.064716 and round(out[173], 6) == 0.881823


def test_difft_batch_random(ctx):
    from tadpole_b200.synth import synth_partition_pairs
    lx, ly = synth_partition_pairs(24, 1500, 60, seed=7)
    # non-contiguous labels and label-0 runs too
    lx[3, 100:140] = 5; ly[4, :] = 0; lx[5, :] = 1; ly[5, :] = 1
    out = ctx.difft_batch(lx, ly)
    for p in range(lx.shape[0]):
        ref = O.difft_from_labels(lx[p], ly[p])
        assert (out[p] == ref).all(), p


def test_difft_many_distinct_labels_uses_global_table(ctx):
    L = 6000
    lx = np.arange(1, L + 1, dtype=np.int32)[None, :]          # every bin its own TAD: 3L keys
    ly = (np.arange(L, dtype=np.int32) // 2 + 1)[None, :]
    out = ctx.difft_batch(lx, ly)
    assert (out[0] == O.difft_from_labels_c(lx[0], ly[0])).all()


def test_difft_full_size_properties(ctx):
    """BASELINE config 5 size (L = 15000): size-independent properties + sampled oracle pairs."""
    from tadpole_b200.synth import synth_partition_pairs
    lx, ly = synth_partition_pairs(64, 15000, 500, seed=11)
    out = ctx.difft_batch(lx, ly)
    assert (np.diff(out, axis=1) >= 0).all() and (out[:, -1] == 1.0).all()
    same = ctx.difft_batch(lx, lx)                             # identical calls: only label-0 bins differ... none do
    assert (same == 0).all()
    sym = ctx.difft_batch(ly, lx)
    assert (sym == out).all()                                  # symmetric in its arguments
    for p in (0, 63):
        assert (out[p] == O.difft_from_labels_c(lx[p], ly[p])).all()


# ---- several calls in flight on one GPU ------------------------------------------------------------------
def test_batch_of_calls_equals_sequential_calls():
    """TADpole_batch runs independent calls concurrently on one GPU (one context / stream / host thread each);
    every result must be bit-identical to the call done alone."""
    from tadpole_b200 import Context, ContextPool, TADpole, TADpole_batch, api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    mats = [synth_hic(n, seed=40 + i) for i, n in enumerate((300, 1200, 600, 1200, 300, 900))]
    pool = ContextPool(0, streams=3)
    got = TADpole_batch(mats, pool=pool)
    pool.close()
    solo = Context(0)
    for m, g in zip(mats, got):
        r = TADpole(m, ctx=solo)
        assert (r.n_pcs, r.optimal_n_clusters) == (g.n_pcs, g.optimal_n_clusters)
        assert np.array_equal(r.dendro.seqdist, g.dendro.seqdist)
        assert np.array_equal(r.scores, g.scores, equal_nan=True)
        assert r.clusters.keys() == g.clusters.keys()
        for key in r.clusters:
            assert np.array_equal(r.clusters[key], g.clusters[key])
    solo.close()
