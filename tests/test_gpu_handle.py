"""GPU tests of the SURVEY 8(f) rows built on top of the path: the device-resident pipeline handle (tp_recall,
all-candidate dendrograms) and the opt-in centromere repairs.  Checker: the oracle run afresh with the same arguments."""
import numpy as np
import pytest

from oracle import tadpole_oracle as O

pytestmark = pytest.mark.gpu


def same_result(tp, ref):
    assert tp.n_pcs == ref.n_pcs and tp.optimal_n_clusters == ref.optimal_n_clusters
    assert tp.scores.shape == ref.scores.shape and (np.isnan(tp.scores) == np.isnan(ref.scores)).all()
    msk = ~np.isnan(ref.scores)
    np.testing.assert_allclose(tp.scores[msk], ref.scores[msk], rtol=1e-8)       # CH: 1e-8 relative end to end
    assert sorted(tp.clusters) == sorted(str(k) for k in ref.clusters)
    for k, tab in ref.clusters.items():
        assert np.array_equal(tp.clusters[str(k)], tab), f"TAD boundaries differ at level {k}"


@pytest.mark.parametrize("n", [300, 1200])
def test_recall_equals_fresh_call(ctx, n):
    from tadpole_b200 import TADpole, api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    m = synth_hic(n, seed=6)
    tp = TADpole(m, max_pcs=60, ctx=ctx)
    same_result(tp, O.tadpole(m, max_pcs=60))
    launches0 = ctx.launches
    # another min_clusters, then fewer PCs: only the sweep is repeated, results equal those of a fresh call
    tp2 = tp.recall(max_pcs=60, min_clusters=4)
    same_result(tp2, O.tadpole(m, max_pcs=60, min_clusters=4))
    tp3 = tp2.recall(max_pcs=25, min_clusters=2)
    same_result(tp3, O.tadpole(m, max_pcs=25))
    assert ctx.launches - launches0 < 2 * 12                    # two recalls: a handful of launches each, no PCA
    fresh = TADpole(m, max_pcs=25, ctx=ctx)
    assert fresh.n_pcs == tp3.n_pcs and fresh.optimal_n_clusters == tp3.optimal_n_clusters
    assert all(np.array_equal(fresh.clusters[k], tp3.clusters[k]) for k in fresh.clusters)
    with pytest.raises(RuntimeError, match="used for another matrix"):
        tp3.recall(max_pcs=10)                                  # `fresh` replaced the resident state
    with pytest.raises(Exception, match="holds 25"):
        fresh.recall(max_pcs=40)


def test_every_candidate_dendrogram_is_resident(ctx, synth_cache):
    from tadpole_b200 import TADpole, api
    api.QUIET = True
    c = synth_cache(200)
    tp = TADpole(c["mat"], ctx=ctx)
    for n_pcs in (1, 2, 17, tp.n_pcs, c["k"]):
        d = tp.dendro_for(n_pcs)
        oseq, _ = O.coniss_lw(c["pcs"][:, :n_pcs])
        # same merge order; heights within the seqdist tolerance of the parity tests
        assert (np.argsort(d.seqdist, kind="stable") == np.argsort(oseq, kind="stable")).all()
        np.testing.assert_allclose(d.seqdist, oseq, rtol=1e-9, atol=1e-11 * oseq.max())
        assert d.merge.shape == (oseq.size, 2) and (d.labels == c["lm"].names).all()
    assert np.array_equal(tp.dendro_for(tp.n_pcs).seqdist, tp.dendro.seqdist)


def test_centromere_fix_mode(ctx):
    from tadpole_b200 import TADpole, load_mat, api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    m = synth_hic(700, seed=4, centromere=True)
    # a bad column inside the q arm: the reference removes the wrong row (quirk Q3); the fixed mode the right one
    lm0 = O.load_mat_numeric(m, centromere_search=True)
    q0 = int(lm0.centromere[-1])
    m[q0 + 40, :] = 0; m[:, q0 + 40] = 0
    tp = TADpole(m, centromere_search=True, centromere_fix=True, ctx=ctx)
    ref = O.tadpole(m, centromere_search=True, fix_q_arm=True)
    for arm in ("p", "q"):
        assert tp[arm].n_pcs == ref.arms[arm].n_pcs and tp[arm].optimal_n_clusters == ref.arms[arm].optimal_n_clusters
        for k, tab in ref.arms[arm].clusters.items():
            assert np.array_equal(tp[arm].clusters[str(k)], tab)
        rs = ref.arms[arm].scores
        assert tp[arm].scores.shape == rs.shape
        np.testing.assert_allclose(tp[arm].scores[~np.isnan(rs)], rs[~np.isnan(rs)], rtol=1e-8)
    assert np.array_equal(tp.merging_arms, ref.merging_arms)
    la = load_mat(m, centromere_search=True, centromere_fix=True, ctx=ctx)
    assert q0 + 41 not in la.q.names and np.array_equal(la.q.names, O.load_mat_numeric(m, centromere_search=True, fix_q_arm=True).q.names)
    # the default stays bug-compatible
    tpc = TADpole(m, centromere_search=True, ctx=ctx)
    refc = O.tadpole(m, centromere_search=True)
    assert np.array_equal(tpc.merging_arms, refc.merging_arms) and "scores" not in tpc.q
    # no split possible (no centromere): the reference errors (Q4), the fixed mode processes the chromosome whole
    m2 = synth_hic(300, seed=8, zero_frac=0.0)
    with pytest.raises(ValueError):
        TADpole(m2, centromere_search=True, bad_frac=0.0, ctx=ctx)
    whole = TADpole(m2, centromere_search=True, bad_frac=0.0, centromere_fix=True, ctx=ctx)
    same_result(whole, O.tadpole(m2, bad_frac=0.0))
    # ... or the longest bad stretch touches an end
    m3 = synth_hic(300, seed=8, zero_frac=0.0)
    m3[:4, :] = 0; m3[:, :4] = 0
    with pytest.raises(ValueError):
        TADpole(m3, centromere_search=True, bad_frac=0.0, ctx=ctx)
    same_result(TADpole(m3, centromere_search=True, bad_frac=0.0, centromere_fix=True, ctx=ctx), O.tadpole(m3, bad_frac=0.0))


def test_tadpole_tune_environment(monkeypatch):
    """TADPOLE_TUNE applies tp_ctx_set at context creation; unknown keys and malformed items are errors."""
    from tadpole_b200 import Context
    monkeypatch.setenv("TADPOLE_TUNE", "pca_block=288,iop_final_min_n=4096")
    c = Context(0)
    c.close()
    for bad in ("no_such_key=1", "pca_block", "pca_block=abc", "=3"):
        monkeypatch.setenv("TADPOLE_TUNE", bad)
        with pytest.raises(Exception, match="TADPOLE_TUNE|unknown key"):
            Context(0)


def test_blocking_host_wait_gives_the_same_result(ctx):
    """sync_blocking only changes HOW the host thread waits for the stream (sleeping on a blocking-sync event instead of
    spinning in cudaStreamSynchronize): the call must return the same bits."""
    from tadpole_b200.synth import synth_hic
    m = synth_hic(700, seed=4)
    a = ctx.call(m, max_pcs=80)
    ctx.set("sync_blocking", 1)
    try:
        b = ctx.call(m, max_pcs=80)
    finally:
        ctx.set("sync_blocking", 0)
    assert (a["n_pcs"], a["n_clusters"], a["nf"]) == (b["n_pcs"], b["n_clusters"], b["nf"])
    assert np.array_equal(a["seqdist"], b["seqdist"])
    assert np.array_equal(a["scores"], b["scores"], equal_nan=True)
