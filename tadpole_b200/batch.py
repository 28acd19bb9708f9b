"""Several independent TADpole() calls in flight on one GPU.

The reference processes one chromosome per call; genome-wide use loops over chromosomes (and the foreach workers of
R/TADpole.R:103-104 only parallelise inside one call).  On a B200 a single 2 000-bin call keeps most of the 148 SMs
idle: its b x b eigen / Cholesky kernels run on one 8-CTA cluster and the CONISS sweep on 200 warps.  Independent calls
therefore overlap almost perfectly when each has its own context (own stream, own device buffers): this module runs a
list of matrices through a small pool of contexts, one host thread per context (ctypes releases the GIL while the
library runs).  Results are the same objects, in the same order, as calling TADpole() on each matrix in turn.
"""
from __future__ import annotations

import threading

from . import _lib
from .api import TADpole

__all__ = ["ContextPool", "TADpole_batch"]


class ContextPool:
    """`streams` contexts on one device, created once and reused."""

    def __init__(self, device=0, streams=4):
        self.device = device
        self.contexts = [_lib.Context(device) for _ in range(int(streams))]

    def __len__(self):
        return len(self.contexts)

    def close(self):
        for c in self.contexts:
            c.close()
        self.contexts = []

    def map(self, fn, items):
        """fn(ctx, item) for every item, at most one item per context at a time; results in input order."""
        items = list(items)
        out = [None] * len(items)
        errors = []
        nxt = [0]
        lock = threading.Lock()

        def worker(ctx):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(items) or errors:
                    return
                try:
                    out[i] = fn(ctx, items[i])
                except BaseException as exc:      # surfaced in the caller's thread
                    errors.append(exc)
                    return

        threads = [threading.Thread(target=worker, args=(c,)) for c in self.contexts[: max(1, min(len(self.contexts), len(items)))]]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return out


def TADpole_batch(mat_files, max_pcs=200, min_clusters=2, bad_frac=0.01, centromere_search=False, pool=None,
                  device=0, streams=4):
    """TADpole() on every matrix of `mat_files` (paths or arrays), `streams` calls in flight on one GPU."""
    own = pool is None
    pool = pool or ContextPool(device, streams)
    try:
        return pool.map(lambda ctx, m: TADpole(m, max_pcs=max_pcs, min_clusters=min_clusters, bad_frac=bad_frac,
                                               centromere_search=centromere_search, ctx=ctx), mat_files)
    finally:
        if own:
            pool.close()
