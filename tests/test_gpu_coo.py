"""GPU tests of the sparse input (csrc/ingest.cu: tp_ingest_coo, tp_ingest_coo_file) through the C ABI: the dense matrix
built in HBM from upper-triangle pixels must equal, bit for bit, the matrix the reference would have read from the
equivalent dense file (R/TADpole.R:17,20), restated by oracle.coo_to_dense / oracle.read_coo_text; and a TADpole() call
on the pixels must return what the call on the dense matrix returns."""
import numpy as np
import pytest

from oracle import tadpole_oracle as O

pytestmark = pytest.mark.gpu


def same(a, b):
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and \
        np.array_equal(a.view(np.uint64)[~np.isnan(a)], b.view(np.uint64)[~np.isnan(b)])


def pixels_text(b1, b2, v, sep="\t", header=False):
    def one(x):
        return "NA" if x != x else (str(int(x)) if float(x).is_integer() else repr(float(x)))
    body = "".join(f"{a}{sep}{b}{sep}{one(c)}\n" for a, b, c in zip(b1.tolist(), b2.tolist(), v.tolist()))
    return (f"bin1_id{sep}bin2_id{sep}count\n" if header else "") + body


@pytest.mark.parametrize("n", [2, 17, 200, 1500])
def test_pixels_from_arrays(ctx, n):
    from tadpole_b200.synth import synth_hic
    m = synth_hic(n, seed=n) if n >= 64 else np.random.default_rng(n).poisson(1.0, (n, n)).astype(float)
    b1, b2, v = O.dense_to_coo(m)
    perm = np.random.default_rng(1).permutation(b1.size)            # pixel order does not matter
    ptr, nn = ctx.ingest_coo(b1[perm], b2[perm], v[perm], n)
    want, below = O.coo_to_dense(b1, b2, v, n)
    assert nn == n and ctx.last_coo_below == below == 0
    assert same(ctx.get_ingested(n), want) and same(want, np.triu(m))


def test_one_based_lower_triangle_duplicates_and_nan(ctx):
    n = 50
    rng = np.random.default_rng(5)
    b1 = rng.integers(1, n + 1, 4000).astype(np.int32)
    b2 = rng.integers(1, n + 1, 4000).astype(np.int32)                  # both triangles, many repeated cells
    v = rng.poisson(7.0, 4000).astype(float)
    v[::97] = np.nan
    ptr, nn = ctx.ingest_coo(b1, b2, v, n, index_base=1)
    want, below = O.coo_to_dense(b1, b2, v, n, index_base=1)
    assert below > 0 and ctx.last_coo_below == below
    assert same(ctx.get_ingested(n), want)


def test_more_pixels_than_one_staging_chunk(ctx):
    n = 3000
    rng = np.random.default_rng(9)
    nnz = (2 << 20) * 2 + 12345                                         # 2 full 32 MB chunks and a ragged one
    b1 = rng.integers(0, n, nnz).astype(np.int32)
    b2 = rng.integers(0, n, nnz).astype(np.int32)
    lo, hi = np.minimum(b1, b2), np.maximum(b1, b2)
    v = rng.integers(1, 50, nnz).astype(float)                          # integer counts: sums are exact in any order
    ctx.ingest_coo(lo, hi, v, n)
    want, _ = O.coo_to_dense(lo, hi, v, n)
    assert same(ctx.get_ingested(n), want)


@pytest.mark.parametrize("header", [False, True])
def test_pixels_from_text_file(ctx, tmp_path, header):
    from tadpole_b200.synth import synth_hic
    n = 400
    m = synth_hic(n, seed=4)
    m[3, 9] = 0.1; m[5, 5] = 1234567.891; m[7, 300] = 1e-7; m[0, 1] = np.nan; m[2, 2] = 2.5e10
    b1, b2, v = O.dense_to_coo(m)
    text = pixels_text(b1, b2, v, header=header)
    p = tmp_path / "pixels.tsv"
    p.write_text(text)
    ptr, nn = ctx.ingest_coo(path=str(p), n=n)
    r1, r2, rv = O.read_coo_text(text)
    want, _ = O.coo_to_dense(r1, r2, rv, n)
    assert nn == n and ctx.last_coo_nnz == b1.size
    assert same(ctx.get_ingested(n), want) and same(want, np.triu(m))
    # n left to the file: largest bin + 1
    m2 = m.copy(); m2[:, n - 7:] = 0; m2[n - 7:, :] = 0
    b1, b2, v = O.dense_to_coo(m2)
    p.write_text(pixels_text(b1, b2, v, sep=" ", header=header))
    ptr, nn = ctx.ingest_coo(path=str(p), sep=" ")
    assert nn == n - 7 and same(ctx.get_ingested(nn), np.triu(m2)[:nn, :nn])


def test_counts_left_to_the_host(ctx, tmp_path):
    p = tmp_path / "p.tsv"
    p.write_text("0\t0\t0.1000000000000000055511151231257827\n0\t1\t1.7976931348623157e308\n1\t1\t0x10\n1\t2\t7\n")
    ptr, nn = ctx.ingest_coo(path=str(p))
    got = ctx.get_ingested(nn)
    assert nn == 3 and got[0, 0] == 0.1 and got[0, 1] == 1.7976931348623157e308 and got[1, 1] == 16.0 and got[1, 2] == 7.0
    assert ctx.ingest_stats()["host_fields"] >= 1


def test_errors(ctx, tmp_path):
    from tadpole_b200 import TadpoleError
    with pytest.raises(TadpoleError, match="entry 2 is .7, 1.: outside the 5 bins"):
        ctx.ingest_coo([0, 7, 1], [1, 1, 1], [1.0, 2.0, 3.0], 5)
    with pytest.raises(TadpoleError, match="outside"):
        ctx.ingest_coo([0], [1], [1.0], 5, index_base=1)
    with pytest.raises(TadpoleError, match="2..200000"):
        ctx.ingest_coo([0], [0], [1.0], 1)
    p = tmp_path / "p.tsv"
    p.write_text("0\t1\t3\n1\t2\n2\t2\t1\n")
    with pytest.raises(TadpoleError, match="row 2 is not"):
        ctx.ingest_coo(path=str(p))
    p.write_text("0\t1\t3\n1\t-2\t4\n")
    with pytest.raises(TadpoleError, match="row 2 is not"):
        ctx.ingest_coo(path=str(p))
    p.write_text("0\t1\t3\n1\t2\tx\n")
    with pytest.raises(TadpoleError, match="not a number"):
        ctx.ingest_coo(path=str(p))
    p.write_text("0\t1\t3\n1\t9\t4\n")
    with pytest.raises(TadpoleError, match="row 2 names a bin outside the 5 bins"):
        ctx.ingest_coo(path=str(p), n=5)
    with pytest.raises(TadpoleError, match="cannot open"):
        ctx.ingest_coo(path=str(tmp_path / "missing.tsv"))


def test_tadpole_from_pixels_equals_tadpole_from_matrix(ctx, tmp_path):
    from tadpole_b200 import SparseCounts, TADpole, load_mat
    from tadpole_b200.synth import synth_hic
    n = 1200
    m = synth_hic(n, seed=21)
    ref = TADpole(m, max_pcs=40, ctx=ctx)
    b1, b2, v = O.dense_to_coo(m)
    for src in (SparseCounts(b1, b2, v, n),):
        got = TADpole(src, max_pcs=40, ctx=ctx)
        assert got.n_pcs == ref.n_pcs and got.optimal_n_clusters == ref.optimal_n_clusters
        assert np.array_equal(got.dendro.seqdist, ref.dendro.seqdist)
        assert all(np.array_equal(got.clusters[k], ref.clusters[k]) for k in ref.clusters)
    p = tmp_path / "pixels.tsv"
    p.write_text(pixels_text(b1, b2, v))
    got = TADpole(SparseCounts(path=str(p), n_bins=n), max_pcs=40, ctx=ctx)
    assert got.n_pcs == ref.n_pcs and np.array_equal(got.dendro.seqdist, ref.dendro.seqdist)
    lm = load_mat(SparseCounts(b1, b2, v, n), ctx=ctx)
    assert np.array_equal(lm.keep, load_mat(m, ctx=ctx).keep)
    import scipy.sparse as sp
    got = TADpole(SparseCounts.from_scipy(sp.coo_matrix(np.triu(m))), max_pcs=40, ctx=ctx)
    assert np.array_equal(got.dendro.seqdist, ref.dendro.seqdist)


def test_pixel_arrays_large_enough_for_the_staging_lanes(ctx):
    """Arrays of 64 MB and more go up through the pinned staging lanes (filter.cu, tp_upload_range), three uploads back
    to back re-using the lanes' buffers."""
    n = 6000
    rng = np.random.default_rng(13)
    nnz = 9_000_000                                                    # 72 MB of counts, 36 MB per bin array
    b1 = rng.integers(0, n, nnz).astype(np.int32)
    b2 = rng.integers(0, n, nnz).astype(np.int32)
    lo, hi = np.minimum(b1, b2), np.maximum(b1, b2)
    v = rng.integers(1, 9, nnz).astype(float)
    want = np.zeros((n, n))
    np.add.at(want, (lo, hi), v)
    for _ in range(2):
        ctx.ingest_coo(lo, hi, v, n)
        assert same(ctx.get_ingested(n), want)
    v2 = np.concatenate([v, v])                                        # 144 MB / 72 MB / 72 MB: every array through the lanes
    ctx.ingest_coo(np.concatenate([lo, lo]), np.concatenate([hi, hi]), v2, n)
    assert same(ctx.get_ingested(n), 2 * want)
