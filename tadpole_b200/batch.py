"""Several independent TADpole() calls in flight on the GPUs of one box.

The reference processes one chromosome per call; genome-wide use loops over chromosomes (and the foreach workers of
R/TADpole.R:103-104 only parallelise inside one call).  On a B200 a single 2 000-bin call keeps most of the 148 SMs
idle, so independent calls overlap almost perfectly when each has its own context (own stream, own device buffers).
The overlap is the library's job (tp_call_batch, csrc/group.cu): library-owned host threads keep `streams` calls in
flight per device, over every device of the context, and also build the per-level start / end tables; the caller --
one Python thread here, one R thread behind .Call -- just waits.  Results are the same objects, in the same order, as
calling TADpole() on each matrix in turn.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .api import Tadpole, _messages_optimal, get_context, message, read_matrix
from .hclust import Dendro

__all__ = ["ContextPool", "TADpole_batch"]


class ContextPool:
    """A context whose batch calls keep `streams` calls in flight per device (the pool of per-call contexts lives in the
    library and is reused from batch to batch).  device: one index or a list of devices."""

    def __init__(self, device=0, streams=4):
        self.device = device
        self.streams = int(streams)
        self.ctx = _lib.Context(device)

    def close(self):
        self.ctx.close()


def _tadpole_from_batch_item(res):
    bad = res["bad"]
    message(f"{int(bad.sum())} bad columns found at position(s):")
    message(" ".join(str(i) for i in np.flatnonzero(bad) + 1))
    _messages_optimal(res)
    tp = Tadpole()
    tp.n_pcs = res["n_pcs"]
    tp.optimal_n_clusters = res["n_clusters"]
    tp.dendro = Dendro(res["seqdist"], labels=(np.flatnonzero(~bad) + 1).astype(np.int32))
    tp.clusters = {str(k): t for k, t in res["tables"].items()}
    tp.scores = res["scores"]
    return tp                          # no device handle: the per-call contexts are reused by the next call


def TADpole_batch(mat_files, max_pcs=200, min_clusters=2, bad_frac=0.01, centromere_search=False, pool=None,
                  device=0, streams=4, ctx=None):
    """TADpole() on every matrix of `mat_files` (paths or arrays), `streams` calls in flight per GPU."""
    if centromere_search:
        from .api import TADpole
        c = ctx or (pool.ctx if pool else get_context(device if isinstance(device, int) else device[0]))
        return [TADpole(m, max_pcs=max_pcs, min_clusters=min_clusters, bad_frac=bad_frac, centromere_search=True, ctx=c)
                for m in mat_files]
    own = pool is None and ctx is None
    if ctx is None:
        pool = pool or ContextPool(device, streams)
        ctx, streams = pool.ctx, pool.streams
    try:
        mats = [read_matrix(m, ctx) for m in mat_files]
        res = ctx.call_batch(mats, max_pcs=max_pcs, min_clusters=min_clusters, bad_frac=bad_frac, inflight=streams)
        for r in res:
            if isinstance(r, Exception):
                raise r
        return [_tadpole_from_batch_item(r) for r in res]
    finally:
        if own:
            pool.close()
