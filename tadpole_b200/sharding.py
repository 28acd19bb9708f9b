"""Multi-GPU plumbing: one process per GPU (torch.distributed), candidates of one sweep sharded
rank-interleaved (cost grows with the number of PCs, so interleaving balances the ranks), as
the reference's foreach/%dopar% shards them over fork workers (R/TADpole.R:103-104).  The only
exchange is the tiny per-candidate score rows, combined with one all-reduce; the CONISS data path
itself has no collective.  Independent calls (chromosome batches, arms) need no exchange at all.
"""
from __future__ import annotations

import numpy as np

__all__ = ["candidate_range", "sharded_sweep", "select"]


def candidate_range(rank, world, k):
    """(begin, stop, stride) of the 0-based candidates owned by `rank`."""
    return rank, k, world


def sharded_sweep(sweep_fn, k, rank, world, group=None):
    """sweep_fn(cand_begin, cand_stride) -> (n_cluster[k] with 0 for foreign rows, scores[k, w] NaN
    padded).  Returns the full (n_cluster[k], scores[k, maxlev]) on every rank."""
    import torch
    import torch.distributed as dist
    begin, _, stride = candidate_range(rank, world, k)
    ncl, sc = sweep_fn(begin, stride)
    if world == 1:
        return ncl, sc
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t_ncl = torch.as_tensor(np.asarray(ncl, dtype=np.int64), device=dev)
    dist.all_reduce(t_ncl, op=dist.ReduceOp.SUM, group=group)          # rows are disjoint: sum = union
    ncl_all = t_ncl.cpu().numpy()
    width = int(ncl_all.max())
    # NaN marks "no score"; exchange (value, mask) so that the sum over ranks is exact
    val = np.zeros((k, width))
    msk = np.zeros((k, width))
    w = min(width, sc.shape[1])
    own = ~np.isnan(sc[:, :w])
    val[:, :w][own] = sc[:, :w][own]
    msk[:, :w][own] = 1.0
    t = torch.as_tensor(np.stack([val, msk]), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu().numpy()
    scores = np.where(t[1] > 0, t[0], np.nan)
    return ncl_all.astype(np.int32), scores


def select(scores):
    """which.max(rowMeans(scores, na.rm=TRUE)), which.max(scores[opt, ]) through the C ABI
    (tp_select is host-only); returns 1-based (n_pcs, n_clusters)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    sc = np.ascontiguousarray(scores, dtype=np.float64)
    oc, ol = ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.tp_select(sc.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), sc.shape[0], sc.shape[1], sc.shape[1],
                             ctypes.byref(oc), ctypes.byref(ol)))
    return oc.value + 1, ol.value + 1
