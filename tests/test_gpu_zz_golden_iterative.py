"""GPU parity against the committed golden fixture of a size that takes the subspace iteration and the tcgen05 int8 kernels
(tests/golden/pipeline_n1100.json, made by tests/golden/make_golden.py from the CPU oracle).  Added at the very end of round 1
without a GPU run of its own (the round's GPU budget was spent), hence a file of its own that sorts last."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_pipeline_iterative_path(ctx):
    """Committed fixture for a size that takes the subspace iteration and the tcgen05 int8 kernels (Nf >= 1024):
    tests/golden/pipeline_n1100.json = the oracle's TADpole(max_pcs = 40) result for golden_int_matrix(1100, 8)."""
    import hashlib
    from tadpole_b200 import TADpole
    with open(os.path.join(GOLD, "pipeline_n1100.json")) as fh:
        g = json.load(fh)
    sys.path.insert(0, GOLD)
    from intgen import golden_int_matrix
    m = golden_int_matrix(g["n"], g["seed"])                # integer-only arithmetic: the same bits on any machine
    assert hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest() == g["matrix_sha256"], "the generator drifted"
    tp = TADpole(m, max_pcs=g["max_pcs"], ctx=ctx)
    assert tp.n_pcs == g["n_pcs"] and tp.optimal_n_clusters == g["optimal_n_clusters"]
    assert sorted(tp.clusters) == sorted(g["clusters"])
    for k, tab in g["clusters"].items():
        assert np.array_equal(tp.clusters[k], np.array(tab)), f"TAD boundaries differ at level {k}"
    ref = np.array([[np.nan if x is None else x for x in row] for row in g["scores"]])
    assert tp.scores.shape == ref.shape and (np.isnan(tp.scores) == np.isnan(ref)).all()
    msk = ~np.isnan(ref)
    np.testing.assert_allclose(tp.scores[msk], ref[msk], rtol=1e-8)              # CH: 1e-8 relative end to end
    # dendrogram of the optimal candidate: same merge order; heights follow the PC scores (1e-9 of the largest score), so
    # the smallest of them are only loosely pinned end to end (the sweep alone is held to 1e-11 in test_sweep_* above)
    gs = np.array(g["seqdist"])
    assert (np.argsort(tp.dendro.seqdist, kind="stable") == np.argsort(gs, kind="stable")).all()
    np.testing.assert_allclose(tp.dendro.seqdist, gs, rtol=1e-3, atol=0)
