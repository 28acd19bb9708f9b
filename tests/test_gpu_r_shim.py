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
    n_pcs, n_cl, seq, scores = R.call("C_tp_call_arm", rctx, keep, 200, 2)
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
    r2 = R.call("C_tp_recall", rctx, int(keep.size), 50, 3)
    ref2 = O.tadpole(m, max_pcs=50, min_clusters=3)
    assert r2[0][0] == ref2.n_pcs and r2[1][0] == ref2.optimal_n_clusters
    d7 = R.call("C_tp_dendro", rctx, int(keep.size), 7)
    oseq, _ = O.coniss_lw(ref.pcs[:, :7])
    assert (np.argsort(d7, kind="stable") == np.argsort(oseq, kind="stable")).all()


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
