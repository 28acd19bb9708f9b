"""Multi-GPU plumbing: one process per GPU (torch.distributed), candidates of one sweep sharded
rank-interleaved (cost grows with the number of PCs, so interleaving balances the ranks), as
the reference's foreach/%dopar% shards them over fork workers (R/TADpole.R:103-104).  The only
exchange is the tiny per-candidate score rows, combined with one all-reduce; the CONISS data path
itself has no collective.  Independent calls (chromosome batches, arms) need no exchange at all.
"""
from __future__ import annotations

import numpy as np

__all__ = ["candidate_range", "sharded_sweep", "select", "DistEnv", "arm_plan"]


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


def arm_plan(world):
    """Which ranks work on which chromosome arm (R/TADpole.R:357: the arms are independent from load_mat on):
    the first half of the ranks takes p, the second half q.  Returns {arm: [ranks]}; world == 1 -> both on rank 0."""
    if world == 1:
        return {"p": [0], "q": [0]}
    half = world // 2
    return {"p": list(range(half)), "q": list(range(half, world))}


class DistEnv:
    """torch.distributed bootstrap for the library's own NCCL communicators (one process per GPU).

    slot 0: all ranks (one call spread over the whole job); slot 1: the ranks sharing this rank's chromosome arm.
    The 128-byte NCCL ids travel through torch.distributed object broadcasts; the data path never goes through
    torch.  With a CPU (gloo) process group only the host-side plan and the object exchange are available, which is
    what the CPU tests cover."""

    def __init__(self, ctx=None, init_nccl=True):
        import torch.distributed as dist
        self.dist = dist
        self.ctx = ctx
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.arms = arm_plan(self.world)
        self.my_arm = "p" if self.rank in self.arms["p"] else "q"
        self.arm_ranks = self.arms[self.my_arm]
        self.arm_slot = -1
        self.world_slot = -1
        # sub-groups must be created by every rank, in the same order
        self._groups = {}
        if self.world > 1:
            for arm in ("p", "q"):
                self._groups[arm] = dist.new_group(self.arms[arm]) if len(self.arms[arm]) > 1 else None
        if ctx is not None and init_nccl and self.world > 1:
            self._init_slot(0, list(range(self.world)), None)
            self.world_slot = 0
            if len(self.arm_ranks) > 1:
                self._init_slot(1, self.arm_ranks, self._groups[self.my_arm])
                self.arm_slot = 1
            ctx.comm_select(self.world_slot)

    def _init_slot(self, slot, ranks, group):
        ids = [self.ctx.comm_unique_id() if self.rank == ranks[0] else None]
        self.dist.broadcast_object_list(ids, src=ranks[0], group=group)
        self.ctx.comm_init(ids[0], ranks.index(self.rank), len(ranks), slot)

    def select_world(self):
        if self.ctx is not None:
            self.ctx.comm_select(self.world_slot)

    def select_arm(self):
        if self.ctx is not None:
            self.ctx.comm_select(self.arm_slot)

    def exchange(self, obj):
        """all-gather of small host objects (result summaries); returns the list indexed by rank."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out
