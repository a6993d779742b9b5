"""Oracle: nearest-neighbour retrieval + inverse-distance weighted k-NN vote.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference's ``ANNClassifier`` (reference src/ann.py:6-34) delegates the search to
the third-party ``pynndescent.NNDescent`` (reference src/ann.py:3,11,12,16; version
unpinned, absent from this image).  The reference's author configures it "to
mimic deterministic NN-search" (experiments.ipynb cell 9), so the oracle restates
the search as an EXACT euclidean search with pynndescent's call shape
(``ExactIndex``): parity for the approximate graph search itself is unpinned.
Everything else (h-stacking neighbour lists over query modalities, 1/dist
weights with the zero-distance rule, fp64 weighted vote with ties going to the
lowest class id) follows the reference line by line.
"""
from __future__ import annotations

import numpy as np


class ExactIndex:
    """Exact euclidean index with the call shape of ``pynndescent.NNDescent``.

    ``ExactIndex(X, **ignored)``, ``.prepare()``, ``.query(x, k=, epsilon=)`` ->
    ``(int32 [Nq,k] indices, float32 [Nq,k] distances)`` sorted ascending by
    distance; equal distances are ordered by gallery index (stable).
    Distances are the direct form sqrt(sum((x-y)^2)) evaluated in float64 and
    rounded to float32 (a correctly-rounded stand-in for pynndescent's float32
    euclidean kernel).
    """

    def __init__(self, X, **_ignored):
        self.X = np.ascontiguousarray(X, dtype=np.float32)

    def prepare(self):
        return None

    def query(self, x, k=10, epsilon=0.1, block=256):
        x = np.ascontiguousarray(x, dtype=np.float32)
        Xg = self.X.astype(np.float64)
        nq, ng = x.shape[0], Xg.shape[0]
        k = min(k, ng)
        idx = np.empty((nq, k), dtype=np.int32)
        dist = np.empty((nq, k), dtype=np.float32)
        for s in range(0, nq, block):
            q = x[s:s + block].astype(np.float64)
            d2 = ((q[:, None, :] - Xg[None, :, :]) ** 2).sum(-1) if ng * q.shape[0] * q.shape[1] <= 1 << 24 \
                else _direct_sqdist(q, Xg)
            dd = np.sqrt(d2).astype(np.float32)
            order = np.argsort(dd, axis=1, kind="stable")[:, :k]
            idx[s:s + block] = order.astype(np.int32)
            dist[s:s + block] = np.take_along_axis(dd, order, axis=1)
        return idx, dist


def _direct_sqdist(q, X):
    """sum((q-x)^2) without the [Nq,Ng,d] temporary: loop over queries."""
    out = np.empty((q.shape[0], X.shape[0]), dtype=np.float64)
    for i in range(q.shape[0]):
        diff = X - q[i]
        out[i] = np.einsum("nd,nd->n", diff, diff)
    return out


def inverse_distance_weights(dist):
    """reference src/ann.py:28-34.  w = 1/dist; a row that contains any zero distance
    gets weight 1 on the zero-distance neighbours and 0 elsewhere."""
    dist = np.array(dist, copy=True)
    with np.errstate(divide="ignore"):
        w = 1.0 / dist
    inf_mask = np.isinf(w)
    inf_row = inf_mask.any(axis=1)
    w[inf_row] = inf_mask[inf_row]
    return w


def weighted_vote(classes, weights):
    """Row-wise weighted mode, semantics of sklearn.utils.extmath.weighted_mode
    as used at reference src/ann.py:24: per-class weight sums accumulated in float64,
    classes visited in ascending order with a strict '>' so ties resolve to the
    lowest class id."""
    classes = np.asarray(classes)
    weights = np.asarray(weights)
    best = np.zeros(classes.shape[0], dtype=np.float64)
    best_w = np.zeros(classes.shape[0], dtype=np.float64)
    for c in np.unique(classes):
        tot = np.where(classes == c, weights, 0).astype(np.float64).sum(axis=1)
        take = tot > best_w
        best = np.where(take, c, best)
        best_w = np.maximum(tot, best_w)
    return best.astype(int)


class OracleANNClassifier:
    """Restatement of reference src/ann.py:6-34 over ``ExactIndex``."""

    def __init__(self, X, y, **index_args):          # reference src/ann.py:9-12
        self.y_ = np.array(y, copy=True)
        self.index = ExactIndex(X, **index_args)
        self.index.prepare()

    def kneighbors(self, *X, **query_args):          # reference src/ann.py:15-16
        return tuple(self.index.query(x, **query_args) for x in X)

    def predict(self, *X, **query_args):             # reference src/ann.py:19-25
        pairs = self.kneighbors(*X, **query_args)
        idx = np.hstack([p[0] for p in pairs])
        dist = np.hstack([p[1] for p in pairs])
        w = inverse_distance_weights(dist)
        return weighted_vote(self.y_[idx], w).ravel()
