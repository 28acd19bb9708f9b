"""The pieces one single-threaded host session (R behind .Call) needs to use a whole box, through the C ABI:
tp_call_batch (library-owned threads keep calls in flight), tp_ctx_create_multi (one context over several GPUs: the
sharded path a torchrun job runs, driven from ONE thread), tp_call_arms (arms on disjoint halves of the devices).
The multi-device tests need >= 2 visible GPUs (`gpurun --gpus 2`); on one GPU they are skipped."""
import numpy as np
import pytest

from oracle import tadpole_oracle as O

pytestmark = pytest.mark.gpu


def _same_call(a, b):
    assert a["n_pcs"] == b["n_pcs"] and a["n_clusters"] == b["n_clusters"] and a["nf"] == b["nf"]
    assert np.array_equal(a["bad"], b["bad"]) if "bad" in a and "bad" in b else True
    assert np.array_equal(a["seqdist"], b["seqdist"])
    assert np.array_equal(a["scores"], b["scores"], equal_nan=True)


def test_batch_entry_equals_one_call_at_a_time(ctx):
    from tadpole_b200.synth import synth_hic
    mats = [synth_hic(n, seed=s) for n, s in ((300, 1), (420, 2), (1100, 3), (300, 4), (512, 5), (640, 6), (300, 7))]
    got = ctx.call_batch(mats, inflight=3)
    assert len(got) == len(mats)
    for m, g in zip(mats, got):
        want = ctx.call(m)
        _same_call(g, want)
        assert g["device_ms"] > 0
    # the per-level tables built in the library's threads are the reference's (oracle) tables
    ref = O.tadpole(mats[1])
    assert sorted(got[1]["tables"]) == sorted(ref.clusters)
    for lv, tab in ref.clusters.items():
        assert np.array_equal(got[1]["tables"][lv], tab)
    # column-major (R) inputs, and a failing matrix in the middle does not stop the others
    mixed = [np.asfortranarray(mats[0]), np.zeros((40, 40)), np.asfortranarray(mats[3])]
    out = ctx.call_batch(mixed, inflight=2)
    _same_call(out[0], got[0]); _same_call(out[2], got[3])
    assert isinstance(out[1], Exception)
    assert ctx.call_batch([], inflight=2) == []


def test_batch_of_slightly_different_sizes_many_in_flight(ctx):
    """Eight host threads launching the same kernels with slightly different shared-memory sizes (bins 560..606): the
    opt-in shared-memory limit of a kernel is process-wide state, and 'set my size, launch' used to race with another
    thread's smaller size (cudaErrorInvalidValue in bench.py's end-to-end pass)."""
    from tadpole_b200.synth import synth_hic
    mats = [synth_hic(560 + 2 * i, seed=60 + i) for i in range(24)]
    for rep in range(2):
        got = ctx.call_batch(mats, inflight=8, tables=False)
        assert not any(isinstance(g, Exception) for g in got), [str(g) for g in got if isinstance(g, Exception)][:1]
    for i in (0, 11, 23):
        _same_call(got[i], ctx.call(mats[i]))


def test_batch_on_device_inputs(ctx):
    import torch
    from tadpole_b200.synth import synth_hic
    mats = [synth_hic(400, seed=s) for s in (11, 12, 13)]
    dev = [torch.as_tensor(m, device="cuda") for m in mats]
    torch.cuda.synchronize()
    got = ctx.call_batch(None, device_ptrs=[d.data_ptr() for d in dev], n=400, inflight=3, tables=False)
    for m, g in zip(mats, got):
        _same_call(g, ctx.call(m))


def test_TADpole_batch_objects(ctx):
    from tadpole_b200 import TADpole, TADpole_batch, api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    mats = [synth_hic(350, seed=s) for s in (21, 22, 23, 24)]
    got = TADpole_batch(mats, ctx=ctx, streams=2)
    for m, g in zip(mats, got):
        w = TADpole(m, ctx=ctx)
        assert g.n_pcs == w.n_pcs and g.optimal_n_clusters == w.optimal_n_clusters
        assert sorted(g.clusters) == sorted(w.clusters)
        assert all(np.array_equal(g.clusters[k], w.clusters[k]) for k in w.clusters)
        assert np.array_equal(g.dendro.merge, w.dendro.merge)


needs2 = pytest.mark.skipif("__import__('tadpole_b200')._lib.device_count() < 2", reason="needs >= 2 GPUs")


@pytest.fixture(scope="module")
def mctx():
    from tadpole_b200 import _lib
    from tadpole_b200 import Context
    nd = min(_lib.device_count(), 8)
    c = Context(list(range(nd)))
    yield c
    c.close()


@needs2
def test_multi_device_context_equals_single_gpu(ctx, mctx):
    """one call spread over the GPUs of one process: row-sharded symmetric products, all-gathered operator applications,
    rank-interleaved sweep -- bit-identical to the one-GPU call"""
    from tadpole_b200.synth import synth_hic
    for n, min_n in ((1500, 1024), (4500, 4096)):
        m = synth_hic(n, seed=n)
        ctx.set("dist_min_n", min_n); mctx.set("dist_min_n", min_n)
        want, got = ctx.call(m), mctx.call(m)
        _same_call(got, want)
        # any candidate's dendrogram, whichever device ran it
        for cand in (0, 1, 7, want["k"] - 1):
            a, ao = ctx.dendro(cand, want["nf"])
            b, bo = mctx.dendro(cand, want["nf"])
            assert np.array_equal(a, b) and np.array_equal(ao, bo), cand
        # recall on the sharded context
        r1, r2 = ctx.recall(want["nf"], max_pcs=60, min_clusters=3), mctx.recall(want["nf"], max_pcs=60, min_clusters=3)
        _same_call(r2, r1)
    ctx.set("dist_min_n", 4096); mctx.set("dist_min_n", 4096)
    # fewer candidates than devices
    m = synth_hic(300, seed=9)
    _same_call(mctx.call(m, max_pcs=1), ctx.call(m, max_pcs=1))
    # a matrix that is already on the first device (ingest / device pointer): the other devices receive it over NVLink
    import torch
    d = torch.as_tensor(synth_hic(1500, seed=31), device=f"cuda:{mctx.device}")
    torch.cuda.synchronize()
    mctx.set("dist_min_n", 1024)
    _same_call(mctx.call(device_ptr=d.data_ptr(), n=1500, colmajor=0), ctx.call(d.cpu().numpy()))
    mctx.set("dist_min_n", 4096)


@needs2
def test_arms_on_disjoint_device_halves(ctx, mctx):
    from tadpole_b200 import TADpole, api
    from tadpole_b200.synth import synth_hic
    api.QUIET = True
    m = synth_hic(1400, seed=77, centromere=True)
    one = TADpole(m, centromere_search=True, ctx=ctx)
    many = TADpole(m, centromere_search=True, ctx=mctx)
    ref = O.tadpole(m, centromere_search=True)
    assert np.array_equal(one.merging_arms, many.merging_arms) and np.array_equal(many.merging_arms, ref.merging_arms)
    for arm in ("p", "q"):
        a, b = one[arm], many[arm]
        assert a.n_pcs == b.n_pcs and a.optimal_n_clusters == b.optimal_n_clusters
        assert np.array_equal(a.dendro.seqdist, b.dendro.seqdist)
        assert all(np.array_equal(a.cluster[k], b.cluster[k]) for k in a.cluster)


@needs2
def test_batch_over_all_devices(ctx, mctx):
    from tadpole_b200.synth import synth_hic
    mats = [synth_hic(300 + 20 * s, seed=40 + s) for s in range(9)]
    got = mctx.call_batch(mats, inflight=2)
    for m, g in zip(mats, got):
        _same_call(g, ctx.call(m))
