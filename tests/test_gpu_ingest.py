"""GPU tests of the input side (csrc/ingest.cu) through the C ABI: the matrix parsed on the device from the file's
text must equal, bit for bit, what the reference's read.big.matrix(type='double', sep='\\t') yields
(R/TADpole.R:17), restated by oracle.read_matrix_text (Python float() = correctly rounded strtod)."""
import os

import numpy as np
import pytest

from oracle import tadpole_oracle as O

pytestmark = pytest.mark.gpu


def same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint64)[~np.isnan(a)], b.view(np.uint64)[~np.isnan(b)]) \
        and np.array_equal(np.isnan(a), np.isnan(b))


def ingest_text(ctx, text, sep="\t"):
    _, n = ctx.ingest_tsv(text.encode() if isinstance(text, str) else text, sep=sep)
    return ctx.get_ingested(n)


@pytest.mark.parametrize("n", [2, 3, 17, 200, 601, 3000])
def test_count_matrix(ctx, n):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(n, seed=n) if n >= 64 else np.random.default_rng(n).poisson(40.0, (n, n)).astype(float)
    text = O.matrix_to_text(m)
    got = ingest_text(ctx, text)
    assert same(got, O.read_matrix_text(text)) and same(got, m)
    assert ctx.ingest_stats()["host_fields"] == 0


def test_float_matrix_all_spellings(ctx):
    rng = np.random.default_rng(11)
    n = 257
    m = rng.random((n, n)) * 10.0 ** rng.integers(-8, 9, (n, n))
    m[rng.random((n, n)) < 0.05] = np.nan
    m[rng.random((n, n)) < 0.05] *= -1
    rows = []
    for i in range(n):
        f = []
        for j in range(n):
            v = m[i, j]
            c = (i * n + j) % 7
            if v != v:
                f.append(["NA", "NaN", "", "nan"][(i + j) % 4])
            elif c == 0:
                f.append(repr(float(v)))
            elif c == 1:
                f.append("%.17g" % v)
            elif c == 2:
                f.append("%.6f" % v)
            elif c == 3:
                f.append("%.10E" % v)
            elif c == 4:
                f.append(" %.3e " % v)
            elif c == 5:
                f.append("%.25f" % v)               # > 19 significant digits
            else:
                f.append("+%d" % int(abs(v)))
        rows.append("\t".join(f))
    text = "\r\n".join(rows)                          # CRLF line ends, no newline at the end
    got = ingest_text(ctx, text)
    assert same(got, O.read_matrix_text(text))
    text2 = "\n".join(rows) + "\n\n\n"                 # blank lines at the end are not rows
    assert same(ingest_text(ctx, text2), got)


def test_fields_left_to_the_host(ctx):
    # 40+ digit fields sitting exactly on a rounding boundary, and spellings only strtod knows
    half = "1.00000000000000011102230246251565404236316680908203125"          # 1 + 2^-53: ties to even -> 1.0
    above = "1.000000000000000111022302462515654042363166809082031250000001"
    rows = [[half, above, "0x1p-3"], ["9007199254740993", "Infinity", "1e5"], ["1", "2", "3"]]
    text = "\n".join("\t".join(r) for r in rows) + "\n"
    got = ingest_text(ctx, text)
    want = np.array([[1.0, np.nextafter(1.0, 2.0), 0.125], [9007199254740992.0, np.inf, 1e5], [1, 2, 3]])
    assert same(got, want)
    assert ctx.ingest_stats()["host_fields"] >= 3


def test_errors(ctx):
    from tadpole_b200 import TadpoleError
    with pytest.raises(TadpoleError, match="row 2 has 2 fields"):
        ingest_text(ctx, "1\t2\t3\n4\t5\n6\t7\t8\n")
    with pytest.raises(TadpoleError, match="not a number"):
        ingest_text(ctx, "1\t2\nx1\t3\n")
    with pytest.raises(TadpoleError, match="at least 2 rows"):
        ingest_text(ctx, "1\t2\n")
    with pytest.raises(TadpoleError, match="no data"):
        ingest_text(ctx, "\n\n")
    with pytest.raises(TadpoleError, match="cannot open"):
        ctx.ingest_tsv("/nonexistent/matrix.tsv")
    # a trailing separator is one more (empty) field
    with pytest.raises(TadpoleError, match="row 1 has 3 fields"):
        ingest_text(ctx, "1\t2\t\n3\t4\t\n")


def test_other_separator_and_long_rows(ctx):
    rng = np.random.default_rng(2)
    n = 1500                                            # rows of ~27 KB: several 4 KB tiles per row
    m = rng.random((n, n))
    text = O.matrix_to_text(m, sep=" ")
    assert same(ingest_text(ctx, text, sep=" "), m)


def test_file_larger_than_one_staging_chunk(ctx, tmp_path):
    from tadpole_b200.synth import synth_hic
    n = 2800
    m = synth_hic(n, seed=3) * 1000.0 + 0.5             # ~7 bytes per field: ~40 MB, two 32 MB pinned chunks
    path = tmp_path / "big.tsv"
    text = O.matrix_to_text(m)
    path.write_text(text)
    assert os.path.getsize(path) > (33 << 20)
    _, nn = ctx.ingest_tsv(str(path))
    assert nn == n and same(ctx.get_ingested(n), m)
    st = ctx.ingest_stats()
    assert st["text_bytes"] == len(text) - 1 and st["host_fields"] == 0


def test_tadpole_from_file_equals_tadpole_from_array(ctx, tmp_path):
    from tadpole_b200 import TADpole, load_mat, api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    m = synth_hic(400, seed=9)
    path = tmp_path / "m.tsv"
    path.write_text(O.matrix_to_text(m))
    a = TADpole(str(path), ctx=ctx)
    b = TADpole(m, ctx=ctx)
    assert a.n_pcs == b.n_pcs and a.optimal_n_clusters == b.optimal_n_clusters
    assert np.array_equal(a.scores, b.scores, equal_nan=True) and np.array_equal(a.dendro.seqdist, b.dendro.seqdist)
    assert a.clusters.keys() == b.clusters.keys() and all(np.array_equal(a.clusters[k], b.clusters[k]) for k in a.clusters)
    la, lb = load_mat(str(path), ctx=ctx), load_mat(m, ctx=ctx)
    assert np.array_equal(la.bad_columns, lb.bad_columns) and np.array_equal(la.to_numpy(), lb.to_numpy())
    # the oracle on the oracle's reading of the same file
    ref = O.tadpole(O.read_matrix_text(path.read_text()))
    assert a.n_pcs == ref.n_pcs and a.optimal_n_clusters == ref.optimal_n_clusters
    assert all(np.array_equal(a.clusters[str(k)], t) for k, t in ref.clusters.items())
