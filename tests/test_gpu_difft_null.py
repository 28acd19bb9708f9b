"""GPU tests of the diffT null-distribution workflow (tp_difft_null): random_bed (R/DiffT.R:61-73) drawn on the
device, each partition scored with diffT (R/DiffT.R:19-50).  Integer work: bit-exact against the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import tadpole_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def control_bed():
    with open(os.path.join(GOLD, "difft_control_case.json")) as fh:
        g = json.load(fh)
    return np.array(g["control"]), np.array(g["case"])


def test_draws_equal_the_oracle_generator(ctx):
    from tadpole_b200 import random_bed_batch
    control, _ = control_bed()
    beds = random_bed_batch(control, 50, seed=1234, ctx=ctx)
    assert beds.shape == (50, control.shape[0], 2)
    for i in range(50):
        assert np.array_equal(beds[i], O.random_bed(control, seed=1234, perm=i)), i
    # bad columns: positional removal from start:end, then the first kept bin is never a border
    bad = [1, 2, 50, 51, 52, 120]
    beds = random_bed_batch(control, 20, bad_columns=bad, seed=7, ctx=ctx)
    for i in range(20):
        assert np.array_equal(beds[i], O.random_bed(control, bad_columns=bad, seed=7, perm=i)), i


def test_partition_structure_and_bad_columns(ctx):
    bed = np.array([[11, 40], [41, 90], [91, 95], [96, 300], [301, 410]])
    bad = np.array([1, 5, 6, 7, 200, 201, 399, 400])
    from tadpole_b200 import random_bed_batch
    beds = random_bed_batch(bed, 400, bad_columns=bad, seed=99, ctx=ctx)
    start, end = 11, 410
    borders = beds[:, 1:, 0] + 1                                   # start column is borders - 1
    assert (np.diff(borders, axis=1) > 0).all()                    # sorted, distinct
    assert (borders > start + 1).all() and (borders <= end).all()   # position 1 is bad, position 2 is bins[1]: never drawn
    assert not np.isin(borders - start + 1, bad).any()
    assert (beds[:, 0, 0] == start).all() and (beds[:, -1, 1] == end).all()
    assert (beds[:, 1:, 0] == beds[:, :-1, 1] + 1).all()            # rows tile the extent
    # different seeds give different draws, the same seed the same
    again = random_bed_batch(bed, 400, bad_columns=bad, seed=99, ctx=ctx)
    other = random_bed_batch(bed, 400, bad_columns=bad, seed=100, ctx=ctx)
    assert np.array_equal(beds, again) and not np.array_equal(beds, other)


def test_uniformity_of_border_positions(ctx):
    from tadpole_b200 import random_bed_batch
    bed = np.array([[1, 20], [21, 40], [41, 60], [61, 80], [81, 101]])
    n = 20000
    beds = random_bed_batch(bed, n, seed=5, ctx=ctx)
    borders = (beds[:, 1:, 0] + 1).ravel()
    counts = np.bincount(borders, minlength=102)[2:102]             # candidates: bins 2..101
    expect = n * 4 / 100
    chi2 = ((counts - expect) ** 2 / expect).sum()
    assert counts.sum() == 4 * n and chi2 < 160                    # 99 dof: mean 99, 160 is beyond the 99.99th percentile
    # order statistics of a uniform subset: E[smallest of 4 out of {2..101}] = 1 + 101 / 5 = 21.2 (sd of the mean ~0.11)
    first = beds[:, 1, 0] + 1
    assert abs(first.mean() - 21.2) < 0.6


def test_null_curves_and_totals_equal_oracle(ctx):
    from tadpole_b200 import diffT_null, diffT
    control, case = control_bed()
    res = diffT_null(control, case, nperm=40, seed=3, ctx=ctx)
    assert res.curves.shape[0] == 40 and res.beds.shape == (40, case.shape[0], 2)
    for i in range(40):
        rb = O.random_bed(case, seed=3, perm=i)
        assert np.array_equal(res.beds[i], rb)
        raw = O.difft(control, rb, raw=True)
        want = O.difft(control, rb)
        assert (res.curves[i] == want).all(), i
        assert res.totals[i] == raw[-1], i
        assert (diffT(control, rb, ctx=ctx) == want).all()
    # observed score against the null: the control/case pair of the reference's fixture
    obs = O.difft(control, case, raw=True)[-1]
    assert obs == 1777 and np.isfinite(res.totals).all()


def test_null_with_bad_columns_labels(ctx):
    control, _ = control_bed()
    size = int(control[-1, 1] - control[0, 0] + 1)
    tx = O.bin_index(control, size).astype(np.int32)
    bad = [3, 4, 100]
    r = ctx.difft_null(tx, control.shape[0], 25, bad_positions=bad, seed=11, want_labels=True)
    for i in range(25):
        rb = O.random_bed(control, bad_columns=bad, seed=11, perm=i)
        assert np.array_equal(r["labels"][i], O.bin_index(rb, size)), i
        assert (r["curves"][i] == O.difft_from_labels(tx, O.bin_index(rb, size))).all()


def test_null_errors(ctx):
    from tadpole_b200 import TadpoleError
    with pytest.raises(TadpoleError, match="sample larger than the population"):
        ctx.difft_null(np.ones(5, np.int32), 6, 3)
    r = ctx.difft_null(np.ones(9, np.int32), 1, 4)                  # one TAD: nothing to draw
    assert r["borders"].shape == (4, 0) and (r["totals"] == 0).all()


def test_config5_scale_null(ctx):
    """BASELINE configs[4] as it is used in practice: 1000 random partitions of a 15 000-bin call, drawn and scored on the GPU."""
    L, T = 15000, 500
    rng = np.random.default_rng(0)
    cuts = np.sort(rng.choice(np.arange(2, L + 1), T - 1, replace=False))
    bed = np.stack([np.concatenate(([1], cuts)), np.concatenate((cuts - 1, [L]))], axis=1)
    from tadpole_b200 import diffT_null
    res = diffT_null(bed, nperm=1000, seed=1, ctx=ctx)
    assert res.curves.shape == (1000, L) and (res.curves[:, -1] == 1.0).all()
    assert (np.diff(res.curves, axis=1) >= 0).all()
    for i in (0, 499, 999):                                          # the O(L^2) oracle in C on three of them
        rb = O.random_bed(bed, seed=1, perm=i)
        assert np.array_equal(res.beds[i], rb)
        tx, ty = O.difft_labels(bed, rb)
        want = O.difft_from_labels_c(tx, ty)
        assert (res.curves[i] == want).all()
