"""GPU tests of the R binding: the .Call routines of r-package/TADpoleB200/src/r_shim.c, driven through the stand-in
R runtime (tests/mock_r; R is not installed), must hand R the same values the ctypes binding returns -- the path
R code -> .Call -> C ABI -> CUDA is the product's drop-in boundary."""
import os
import struct
import sys

import numpy as np
import pytest

from oracle import tadpole_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    from mock_r.driver import MockR
    r = MockR()
    yield r
    r.reset()                       # runs the external-pointer finalizer: tp_ctx_destroy


@pytest.fixture(scope="module")
def rctx(R):
    return R.call("C_tp_ctx", 0, raw=True)


def test_filter_and_call_arm_through_the_shim(R, rctx, ctx):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(400, seed=12)
    bad = R.call("C_tp_filter", rctx, m, 0.01)               # an R matrix: column-major doubles in host memory
    want_bad, _, _ = ctx.filter(m, bad_frac=0.01)
    assert bad.dtype == bool and np.array_equal(bad, want_bad)
    keep = np.flatnonzero(~bad)
    n_pcs, n_cl, seq, scores, gen = R.call("C_tp_call_arm", rctx, keep, 200, 2)
    assert gen[0] == R.call("C_tp_generation", rctx)[0]
    want = ctx.call_arm(keep.astype(np.int32))
    assert n_pcs[0] == want["n_pcs"] and n_cl[0] == want["n_clusters"]
    assert np.array_equal(seq, want["seqdist"])
    assert scores.shape == want["scores"].shape and np.array_equal(np.isnan(scores), np.isnan(want["scores"]))
    assert np.array_equal(scores[~np.isnan(scores)], want["scores"][~np.isnan(scores)])
    # padding is R's NA_real_ (NaN with payload 1954), not a plain NaN
    na = scores[np.isnan(scores)]
    assert na.size and all(struct.unpack("Q", struct.pack("d", v))[0] == 0x7FF00000000007A2 for v in na[:50])
    # and equals the oracle
    ref = O.tadpole(m)
    assert n_pcs[0] == ref.n_pcs and n_cl[0] == ref.optimal_n_clusters
    # per-level tables and the optimal labels as the R code asks for them
    levels = np.flatnonzero(~np.isnan(scores[n_pcs[0] - 1])) + 1
    tabs = R.call("C_tp_levels", seq, levels, keep + 1, np.flatnonzero(bad) + 1, False)
    for lv, t in zip(levels, tabs):
        assert np.array_equal(t, ref.clusters[int(lv)])
    labels = R.call("C_tp_labels", seq, int(n_cl[0]), keep + 1, np.flatnonzero(bad) + 1, False)
    assert labels.size == m.shape[0] and labels.max() <= n_cl[0]
    # recall and any candidate's dendrogram from the resident state
    r2 = R.call("C_tp_recall", rctx, float(gen[0]), 50, 3)
    ref2 = O.tadpole(m, max_pcs=50, min_clusters=3)
    assert r2[0][0] == ref2.n_pcs and r2[1][0] == ref2.optimal_n_clusters
    assert r2[4][0] != gen[0]                                # the sweep state was replaced: a new generation
    d7 = R.call("C_tp_dendro", rctx, float(r2[4][0]), 7)
    oseq, _ = O.coniss_lw(ref.pcs[:, :7])
    assert (np.argsort(d7, kind="stable") == np.argsort(oseq, kind="stable")).all()
    # dendro$merge through the shim = the oracle's literal .find.groups loop
    merge = R.call("C_tp_find_groups", seq)
    assert merge.shape == (seq.size, 2) and np.array_equal(merge, O.find_groups(seq)[0])
    # load_mat()'s return value: mat[keep, keep] of the symmetrised matrix
    filt = R.call("C_tp_get_filtered", rctx, keep)
    assert np.array_equal(filt, O.symmetrise_upper(m)[np.ix_(keep, keep)])


def test_stale_handle_is_refused_and_buffers_follow_the_context(R, rctx):
    """a <- TADpole(A); b <- TADpole(B); tadpole_recall(a): the context now holds B (more good bins than A).  The shim must
    refuse A's handle instead of packing B's sweep into A-sized buffers."""
    from mock_r.driver import RError
    from tadpole_b200.synth import synth_hic
    a, b = synth_hic(300, seed=3), synth_hic(420, seed=4)
    bad_a = R.call("C_tp_filter", rctx, a, 0.01)
    res_a = R.call("C_tp_call_arm", rctx, np.flatnonzero(~bad_a), 200, 2)
    bad_b = R.call("C_tp_filter", rctx, b, 0.01)
    res_b = R.call("C_tp_call_arm", rctx, np.flatnonzero(~bad_b), 200, 2)
    with pytest.raises(RError, match="used for another matrix"):
        R.call("C_tp_recall", rctx, float(res_a[4][0]), 100, 2)
    with pytest.raises(RError, match="used for another matrix"):
        R.call("C_tp_dendro", rctx, float(res_a[4][0]), 5)
    # B's own handle works, and every buffer has B's sizes
    r = R.call("C_tp_recall", rctx, float(res_b[4][0]), 100, 2)
    assert r[2].size == int((~bad_b).sum()) - 1 and r[3].shape[0] == 100
    ref = O.tadpole(b, max_pcs=100)
    assert r[0][0] == ref.n_pcs and r[1][0] == ref.optimal_n_clusters
    with pytest.raises(RError, match="n_pcs must be between"):
        R.call("C_tp_dendro", rctx, float(r[4][0]), 101)


def test_arms_and_batch_through_the_shim(R, rctx):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(700, seed=5, centromere=True)
    ref = O.tadpole(m, centromere_search=True)
    bad = R.call("C_tp_filter", rctx, m, 0.01)
    lm = O.load_mat_numeric(m, centromere_search=True)
    kp, kq = np.asarray(lm.p.names) - 1, np.asarray(lm.q.names) - 1
    p, q = R.call("C_tp_call_arms", rctx, kp, kq, 200, 2)
    for got, arm in ((p, ref.arms["p"]), (q, ref.arms["q"])):
        assert got[0][0] == arm.n_pcs and got[1][0] == arm.optimal_n_clusters
    # a batch: results in input order, one failing matrix does not stop the others
    mats = [synth_hic(260, seed=s) for s in (1, 2, 3)] + [np.zeros((50, 50))]
    out = R.call("C_tp_call_batch", rctx, mats, 200, 2, 0.01, 3)
    assert len(out) == 4 and isinstance(out[3][0], str)
    for mm, item in zip(mats[:3], out[:3]):
        o = O.tadpole(mm)
        badv, npcs, ncl, seq, scores, levels, tables = item
        assert npcs[0] == o.n_pcs and ncl[0] == o.optimal_n_clusters and badv.dtype == bool
        assert sorted(int(l) for l in levels) == sorted(o.clusters)
        for lv, t in zip(levels, tables):
            assert np.array_equal(t, o.clusters[int(lv)])


def test_ingest_through_the_shim(R, rctx, tmp_path):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(300, seed=2)
    m[5, 9] = 0.1 + 0.2
    path = tmp_path / "m.tsv"
    path.write_text(O.matrix_to_text(m))
    n = R.call("C_tp_ingest", rctx, str(path))
    assert n[0] == 300
    back = R.call("C_tp_ingested_matrix", rctx)
    assert np.array_equal(back, m)                           # m[i, j] is field (i, j) of the file
    bad = R.call("C_tp_filter", rctx, None, 0.01)            # NULL: the ingested matrix
    obad, _, _ = O.bad_columns(O.symmetrise_upper(m), 0.01)
    assert np.array_equal(bad, obad)
    from mock_r.driver import RError
    with pytest.raises(RError, match="cannot open"):
        R.call("C_tp_ingest", rctx, "/nonexistent.tsv")
    with pytest.raises(RError, match="square numeric matrix"):
        R.call("C_tp_filter", rctx, np.zeros((3, 4)), 0.01)


def test_sparse_input_through_the_shim(R, rctx, tmp_path):
    """sparse_counts() / a Matrix::sparseMatrix in R/tadpole.R: pixels in, the same bad columns and call as the matrix."""
    from tadpole_b200.synth import synth_hic
    m = synth_hic(300, seed=6)
    b1, b2, v = O.dense_to_coo(m)
    got = R.call("C_tp_ingest_coo", rctx, b1 + 1, b2 + 1, v, 300, 1)       # one-based, as R users would pass them
    assert got[0] == 300 and got[1] == 0
    assert np.array_equal(R.call("C_tp_ingested_matrix", rctx), np.triu(m))
    bad = R.call("C_tp_filter", rctx, None, 0.01)
    obad, _, _ = O.bad_columns(O.symmetrise_upper(m), 0.01)
    assert np.array_equal(bad, obad)
    path = tmp_path / "pixels.tsv"
    path.write_text("".join(f"{a}\t{b}\t{int(c)}\n" for a, b, c in zip(b1.tolist(), b2.tolist(), v.tolist())))
    got = R.call("C_tp_ingest_coo_file", rctx, str(path), 0, 0)
    assert got[0] == 300 and got[1] == b1.size and got[2] == 0
    assert np.array_equal(R.call("C_tp_filter", rctx, None, 0.01), obad)
    from mock_r.driver import RError
    with pytest.raises(RError, match="outside the 300 bins"):
        R.call("C_tp_ingest_coo", rctx, b1, b2, v, 300, 1)                 # bin 0 with index_base 1
    with pytest.raises(RError, match="equal length"):
        R.call("C_tp_ingest_coo", rctx, b1, b2[:-1], v, 300, 0)


def test_difft_and_null_through_the_shim(R, rctx, ctx):
    import json
    with open(os.path.join(ROOT, "tests", "golden", "difft_control_case.json")) as fh:
        g = json.load(fh)
    control, case = np.array(g["control"]), np.array(g["case"])
    tx, ty = O.difft_labels(control, case)
    out = R.call("C_tp_difft", rctx, tx, ty, int(tx.size), 1)
    assert (out == np.array(g["normalised"])).all()
    borders, totals, curves = R.call("C_tp_difft_null", rctx, tx, 0, 0, int(control.shape[0]), np.zeros(0, np.int32), 42.0, 16)
    want = ctx.difft_null(tx, control.shape[0], 16, seed=42)
    assert borders.shape == (control.shape[0] - 1, 16) and np.array_equal(borders.T, want["borders"])
    assert np.array_equal(totals, want["totals"]) and np.array_equal(curves.T, want["curves"])
    for i in (0, 15):
        rb = O.random_bed(control, seed=42, perm=i)
        assert np.array_equal(borders[:, i] + control[0, 0], rb[1:, 0] + 1)
        assert (curves[:, i] == O.difft(control, rb)).all()
