"""The 'dendro' member of the tadpole object: rioja's chclust / hclust list, rebuilt from the
seqdist vector the GPU returns (host integer logic, as the R wrapper would do).

rioja::chclust returns list(merge, height, order, labels, method, call, dist.method) with class
c("chclust", "hclust"); merge comes from seqdist through .find.groups (repeated which.min, first
index on ties), height = sort(seqdist), order = 1:n.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Dendro", "find_groups", "cutree"]


def find_groups(seqdist):
    """hclust merge matrix, R convention: negative = singleton object, positive = earlier step (tp_find_groups in the
    library: union-find over the boundaries in rioja's .find.groups order)."""
    from . import _lib
    return _lib.find_groups(seqdist)


def cutree(seqdist, k):
    """stats::cutree(tree, k) for a constrained (contiguous) dendrogram: labels 1..k, left to right."""
    x = np.asarray(seqdist, dtype=np.float64)
    lab = np.ones(x.size + 1, dtype=np.int64)
    if k > 1:
        idx = np.lexsort((np.arange(x.size), x))[x.size - (k - 1):]
        cuts = np.zeros(x.size + 1, dtype=np.int64)
        cuts[idx + 1] = 1
        lab += np.cumsum(cuts)
    return lab


class Dendro:
    """chclust/hclust-like object; merge is built on first use."""

    method = "coniss"
    dist_method = "euclidean"
    call = "rioja::chclust(d = dist(pcs))"

    def __init__(self, seqdist, labels=None):
        self.seqdist = np.asarray(seqdist, dtype=np.float64)
        n = self.seqdist.size + 1
        self.height = np.sort(self.seqdist)
        self.order = np.arange(1, n + 1)
        self.labels = np.arange(1, n + 1) if labels is None else np.asarray(labels)
        self._merge = None

    @property
    def merge(self):
        if self._merge is None:
            self._merge = find_groups(self.seqdist)
        return self._merge

    def cutree(self, k):
        return cutree(self.seqdist, k)

    def __repr__(self):
        return (f"\nCall:\n{self.call}\n\nCluster method   : {self.method}\nDistance         : {self.dist_method}\n"
                f"Number of objects: {self.seqdist.size + 1}\n")
