"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, and
the pure-host entry points (tp_select, tp_assemble) and Python wrappers agree with the oracle.
No CUDA compute is called here."""
import os
import re

import numpy as np
import pytest

from oracle import tadpole_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from tadpole_b200 import _lib
    return _lib


def test_library_exports_every_header_symbol(lib):
    with open(os.path.join(ROOT, "include", "tadpole_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(tp_[a-z_0-9]+)\s*\(", text))
    declared.discard("tp_ctx")
    l = lib.load()
    missing = [s for s in sorted(declared) if not hasattr(l, s)]
    assert not missing, missing
    assert declared == set(lib.EXPORTED)
    assert l.tp_version() >= 100


def test_no_gpu_is_a_loud_error(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.TadpoleError) as e:
        lib.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tadpole_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_tp_select_matches_oracle(lib):
    import ctypes
    rng = np.random.default_rng(0)
    l = lib.load()
    for _ in range(20):
        k, w = int(rng.integers(1, 12)), int(rng.integers(2, 9))
        rows = []
        for r in range(k):
            ncl = int(rng.integers(2, w + 1))
            row = np.full(ncl, np.nan)
            row[1:] = rng.integers(1, 6, ncl - 1).astype(float)     # small ints: plenty of ties
            rows.append(row)
        scores, opcs, ok = O.reduce_scores(rows)
        oc, ol = ctypes.c_int(), ctypes.c_int()
        sc = np.ascontiguousarray(scores)
        rc = l.tp_select(sc.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), sc.shape[0], sc.shape[1], sc.shape[1],
                         ctypes.byref(oc), ctypes.byref(ol))
        assert rc == 0 and (oc.value + 1, ol.value + 1) == (opcs, ok)


def test_tp_assemble_matches_oracle(lib):
    rng = np.random.default_rng(1)
    for trial in range(30):
        n = int(rng.integers(8, 60))
        bad_mask = rng.random(n) < 0.2
        bad_mask[rng.integers(0, n)] = False
        names = np.flatnonzero(~bad_mask) + 1
        nf = names.size
        if nf < 3:
            continue
        seq = rng.integers(1, 5, nf - 1).astype(float).cumsum()[rng.permutation(nf - 1)]
        seq[rng.integers(0, nf - 1)] = seq[0]                       # a tie
        for k in (1, 2, min(5, nf)):
            good = O.cutree(seq, k)
            for bad in (np.flatnonzero(bad_mask) + 1, None):
                tab, labels = lib.assemble(seq, k, names, bad)
                if bad is None:
                    ref_fixed = good
                    _, lens = O._rle(good)
                    eb = np.cumsum(lens)
                    ref_tab = np.stack([np.concatenate(([1], eb[:-1] + 1)), eb], axis=1)
                else:
                    ref_fixed = O.fixed_labels(good, names, bad)
                    ref_tab = O.coords_from_labels(ref_fixed)
                assert labels[: ref_fixed.size].tolist() == ref_fixed.tolist()
                assert tab.tolist() == ref_tab.tolist()


def test_assemble_q_arm_quirk_duplicate_names(lib):
    """Quirk Q3: bad columns that were never removed appear both as good rows and as bad names."""
    names = np.array([11, 12, 13, 14, 15, 16])
    bad = np.array([13, 15])                                         # still present among names
    seq = np.array([1.0, 5.0, 2.0, 9.0, 3.0])
    good = O.cutree(seq, 3)
    ref = O.fixed_labels(good, names, bad)
    tab, labels = lib.assemble(seq, 3, names, bad)
    assert labels[: ref.size].tolist() == ref.tolist() and ref.size == 8
    assert tab.tolist() == O.coords_from_labels(ref).tolist()


def test_python_host_helpers_match_oracle():
    from tadpole_b200 import api, hclust
    rng = np.random.default_rng(2)
    seq = rng.random(40)
    assert hclust.find_groups(seq).tolist() == O.find_groups(seq)[0].tolist()
    # the library's union-find against the oracle's literal repeated-which.min loop, ties and long chains included
    for n1, levels in ((1, 3), (2, 2), (257, 4), (1500, 50)):
        s2 = np.round(rng.random(n1) * levels)
        assert hclust.find_groups(s2).tolist() == O.find_groups(s2)[0].tolist(), n1
    s3 = np.cumsum(rng.random(300))                       # one chain growing to the right, then to the left
    assert hclust.find_groups(s3).tolist() == O.find_groups(s3)[0].tolist()
    assert hclust.find_groups(s3[::-1]).tolist() == O.find_groups(s3[::-1])[0].tolist()
    for k in (1, 2, 7, 41):
        assert hclust.cutree(seq, k).tolist() == O.cutree(seq, k).tolist()
    bed = np.array([[5, 9], [10, 20], [18, 25]])
    assert api.bin_index(bed, 21).tolist() == O.bin_index(bed, 21).tolist()
    bx, by = np.array([[3, 8], [9, 15]]), np.array([[1, 6], [7, 12]])
    tx, ty = api._difft_labels(bx, by)
    ox, oy = O.difft_labels(bx, by)
    assert tx.tolist() == ox.tolist() and ty.tolist() == oy.tolist()
    with pytest.raises(ValueError):
        api._difft_labels(bx, by[:1])
    rb = api.random_bed(np.array([[1, 10], [11, 30], [31, 40]]), rng=np.random.default_rng(0))
    assert rb.shape == (3, 2) and rb[0, 0] == 1 and rb[-1, 1] == 40


def test_centromere_split_host_logic_matches_oracle():
    from tadpole_b200 import api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    m = synth_hic(400, seed=2, centromere=True)
    sym = O.symmetrise_upper(m)
    bad, _, _ = O.bad_columns(sym, 0.01)
    (kp, bp), (kq, bq), cen = api._split_centromere(bad)
    lm = O.load_mat_numeric(m, centromere_search=True)
    assert (kp + 1).tolist() == lm.p.names.tolist() and (kq + 1).tolist() == lm.q.names.tolist()
    assert cen.tolist() == lm.centromere.tolist()
    assert (bq is None) == (lm.q.bad_columns is None)


def test_assemble_levels_matches_per_level_calls():
    """tp_assemble_levels (one ranking for all levels) gives the tables of tp_assemble level by level,
    with bad bins re-inserted, without, and with bad_columns = NULL."""
    from tadpole_b200 import _lib
    rng = np.random.default_rng(11)
    nf = 300
    seq = rng.random(nf - 1)
    seq[40] = seq[41] = seq[42]                       # ties: first index merges first
    allbins = np.arange(1, nf + 40 + 1)
    bad = np.sort(rng.choice(allbins, 40, replace=False)).astype(np.int32)
    names = np.setdiff1d(allbins, bad).astype(np.int32)
    levels = np.array([1, 2, 3, 17, 120, nf], dtype=np.int32)
    for b in (bad, np.zeros(0, np.int32), None):
        got = _lib.assemble_levels(seq, levels, names, b)
        for k in levels:
            ref, _ = _lib.assemble(seq, int(k), names, b)
            assert np.array_equal(got[int(k)], ref), (k, None if b is None else len(b))
    # levels in any order, with repeats: every level's rows land in its own slot of the output
    import ctypes
    lib = _lib.load()
    lv = np.array([120, 2, nf, 2, 17, 1, 120], dtype=np.int32)
    cap = int(lv.sum()) + lv.size * (bad.size + 2)
    start, end = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    off = np.zeros(lv.size + 1, np.int32)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    assert lib.tp_assemble_levels(seq.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nf, ip(lv), lv.size, ip(names), ip(bad),
                                  bad.size, ip(start), ip(end), ip(off)) == 0
    assert off.tolist() == [0] + np.cumsum(lv).tolist()
    for i, k in enumerate(lv):
        ref, _ = _lib.assemble(seq, int(k), names, bad)
        assert np.array_equal(np.stack([start[off[i]:off[i + 1]], end[off[i]:off[i + 1]]], axis=1), ref), (i, k)


@pytest.mark.parametrize("n,nranks", [(300, 2), (500, 4), (700, 8), (513, 3), (640, 5), (129, 2), (260, 4)])
def test_symshard_rule_computes_every_block_pair_once(n, nranks):
    """The rule by which the ranks of a sharded call divide a symmetric product (csrc/common.cuh SymShard), through the
    host-only hooks: replaying the tile launches of every rank, each element is written by its row owner or is the mirror
    of an element written by ITS row owner; inside a diagonal block the upper triangle is computed and mirrored on store."""
    from tadpole_b200 import _lib
    lib = _lib.load()
    rpr = -(-(-(-n // nranks)) // 64) * 64
    who = np.full((n, n), -1)
    tiles = np.zeros(nranks, dtype=int)
    for rank in range(nranks):
        r0 = min(rank * rpr, n); r1 = min(r0 + rpr, n)
        for m0 in range(r0, r1, 128):
            r_hi = min(m0 + 127, r1 - 1)
            for n0 in range(0, n, 64):
                c_hi = min(n0 + 63, n - 1)
                if not lib.tp_test_ss_tile(m0, r_hi, n0, c_hi, nranks, rpr):
                    continue
                tiles[rank] += 1
                who[m0:r_hi + 1, n0:c_hi + 1] = rank
                if n0 // rpr == rank:                     # mirror on store inside the owner's diagonal block
                    who[n0:c_hi + 1, m0:r_hi + 1] = np.where(who[n0:c_hi + 1, m0:r_hi + 1] < 0, rank, who[n0:c_hi + 1, m0:r_hi + 1])
    own = np.arange(n) // rpr
    need = np.array([[lib.tp_test_ss_need(i, j, nranks, rpr) for j in range(n)] for i in range(n)], dtype=bool)
    diag = own[:, None] == own[None, :]
    assert (who[diag] == np.broadcast_to(own[:, None], (n, n))[diag]).all()
    off = ~diag
    assert (need ^ need.T)[off].all()                                   # exactly one of (i, j), (j, i) is computed
    assert (who[off & need] == np.broadcast_to(own[:, None], (n, n))[off & need]).all()
    if n >= 128 * nranks:                                               # and the work is balanced
        assert tiles.max() <= 1.6 * max(tiles.min(), 1) + 2


def test_assemble_levels_random_against_per_level_reference():
    """tp_assemble_levels builds every level's table from cluster intervals; tp_assemble materialises labels, applies
    fix_values and run-length encodes (the reference's own order of steps, R/TADpole.R:470-497,503-510).  They must agree
    on every level for arbitrary bad-bin patterns, including bad names that duplicate good ones (q-arm quirk Q3), bad bins
    at both ends and adjacent bad runs."""
    from tadpole_b200 import _lib
    rng = np.random.default_rng(2024)
    for trial in range(150):
        nf = int(rng.integers(2, 50))
        seq = rng.random(nf - 1)
        if nf > 4 and trial % 3 == 0:
            seq[1] = seq[2]                                   # a tie
        total = nf + int(rng.integers(0, 25))
        allbins = np.arange(1, total + 1)
        nb = total - nf
        bad = np.sort(rng.choice(allbins, nb, replace=False)).astype(np.int32)
        names = np.setdiff1d(allbins, bad).astype(np.int32)
        if trial % 5 == 0 and nb:                             # quirk Q3: bad names that were never removed from the arm
            bad = np.sort(np.concatenate([bad, rng.choice(names, min(3, nf), replace=False)])).astype(np.int32)
        levels = np.arange(1, nf + 1, dtype=np.int32)
        for b in (bad, None):
            got = _lib.assemble_levels(seq, levels, names, b)
            for k in levels:
                ref, _ = _lib.assemble(seq, int(k), names, b)
                assert np.array_equal(got[int(k)], ref), (trial, nf, k, None if b is None else b.tolist())
