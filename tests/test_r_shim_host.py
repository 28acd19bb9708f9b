"""CPU tests of the R binding: r-package/TADpoleB200/src/r_shim.c compiles (against the stand-in R headers in
tests/mock_r, R itself is not installed), registers every .Call routine the R code uses, keeps PROTECT balanced,
turns library errors into R errors, and its host-only routine agrees with the ctypes binding."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.fixture(scope="module")
def R():
    import __graft_entry__ as g
    g.build()
    from mock_r.driver import MockR
    r = MockR()
    yield r
    r.reset()


def test_every_routine_the_r_code_calls_is_registered(R):
    with open(os.path.join(ROOT, "r-package", "TADpoleB200", "R", "tadpole.R")) as fh:
        used = set(re.findall(r"\.Call\((C_tp_[a-z_]+)", fh.read()))
    with open(os.path.join(ROOT, "r-package", "TADpoleB200", "src", "r_shim.c")) as fh:
        src = fh.read()
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(C_tp_[a-z_]+)", \(DL_FUNC\)&\1, (\d+)\}', src))
    assert used and used <= set(table), used - set(table)
    for name, nargs in table.items():
        assert R.registered(name) == nargs
        # the C definition takes as many SEXPs as the table says
        sig = re.search(r"SEXP %s\(([^)]*)\)" % name, src).group(1)
        assert (0 if sig.strip() == "void" else sig.count("SEXP")) == nargs, name


def test_library_errors_become_r_errors(R):
    from mock_r.driver import RError
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RError, match="no CPU fallback"):
            R.call("C_tp_ctx", 0)
    with pytest.raises(RError, match="takes 5 arguments"):
        R.call("C_tp_levels", 1, 2)
    with pytest.raises(RError, match="no such routine"):
        R.call("C_tp_nothing")


def test_levels_routine_equals_ctypes_binding(R):
    from tadpole_b200 import _lib
    rng = np.random.default_rng(0)
    nf = 120
    seq = rng.random(nf - 1) * 10
    names = np.delete(np.arange(1, nf + 6), [3, 50, 51, 90, 124])
    bad = np.array([4, 51, 52, 91, 125])
    levels = np.array([2, 5, 17, 40])
    want = _lib.assemble_levels(seq, levels, names, bad)
    got = R.call("C_tp_levels", seq, levels, names, bad, False)
    assert len(got) == 4
    for lv, tab in zip(levels, got):
        assert np.array_equal(tab, want[int(lv)])
    # attr(mat, 'bad_columns') is NULL
    want = _lib.assemble_levels(seq, levels, np.arange(1, nf + 1), None)
    got = R.call("C_tp_levels", seq, levels, np.arange(1, nf + 1), np.zeros(0, np.int32), True)
    assert all(np.array_equal(t, want[int(lv)]) for lv, t in zip(levels, got))
