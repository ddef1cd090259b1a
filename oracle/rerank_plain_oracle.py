"""CPU restatement of the alternate re-ranker.  TEST INFRASTRUCTURE ONLY.

``reid/rerank_plain.py:127-178 re_ranking(input_feature_source, input_feature, k=20, lambda_value=0.1)``: the plain
kNN-set Jaccard variant both drivers carry as a commented-out import (``selftraining.py:29``).  No CUDA path exists for
it yet; this file pins the arithmetic so that one can be written against it:

  S_i    = { j != i : d2[i,j] <= (k-th smallest entry of row i of d2, the diagonal included) }      (:165-170)
  J[i,j] = 1 - |S_i & S_j| / |S_i | S_j|      (scipy ``cdist(.., 'jaccard')`` on the boolean rows; 0 when both are empty)
  final  = J * (1 - lambda) + (v_i + v_j) * lambda,   v as in reid/rerank.py:36-40                  (:133-146,174)

restated through sorted neighbour lists and set intersections instead of an N x N boolean matrix and a dense boolean
``cdist`` -- the form a GPU kernel takes (rows hold ~k entries; ties at the threshold can make them longer).
``mode`` as in oracle/ssg_oracle.py: 'ref' = float16 storage as shipped, 'f32' = float16 -> float32.
Pinned bit for bit against the unmodified reference in tests/test_oracle_vs_reference.py.
"""
import numpy as np
from scipy.spatial.distance import cdist

from . import ssg_oracle as O


def knn_sets(od, k):
    """rerank_plain.py:165-170: per row the indices within the k-th smallest distance (ties kept), self removed."""
    sets = []
    for i in range(od.shape[0]):
        row = od[i]
        thr = np.partition(row, k - 1)[k - 1]
        idx = np.nonzero(row <= thr)[0]
        sets.append(idx[idx != i])
    return sets


def jaccard_sets(sets, n, dtype):
    """scipy's boolean Jaccard distance (rerank_plain.py:173) from the neighbour lists."""
    inv = [[] for _ in range(n)]
    for i, s in enumerate(sets):
        for j in s:
            inv[j].append(i)
    size = np.array([len(s) for s in sets], dtype=np.float64)
    out = np.empty((n, n), dtype=np.float64)
    for i, s in enumerate(sets):
        inter = np.zeros(n, dtype=np.float64)
        for j in s:
            inter[inv[j]] += 1.0
        union = size[i] + size - inter
        with np.errstate(invalid="ignore", divide="ignore"):
            out[i] = np.where(union > 0, (union - inter) / union, 0.0)
    return out.astype(dtype)


def re_ranking_plain(src, tgt, k=20, lambda_value=0.1, mode="f32"):
    """-> final_dist float64 [N,N] (the reference returns it twice, rerank_plain.py:178)."""
    dt = np.float32 if mode == "f32" else np.float16
    vec = O.source_vector(tgt, src, mode)                        # rerank_plain.py:133-138 == rerank.py:36-40
    od = O.original_distance(tgt, mode)                          # :160-161
    n = od.shape[0]
    jac = jaccard_sets(knn_sets(od, k), n, dt)
    return O.final_distance(jac, vec, lambda_value)              # :139-146,174 (v_i + v_j summed in the storage dtype)


def re_ranking_lh(src, tgt, k1=20, k2=6, lambda_value=0.2, mode="f32"):
    """reid/rerank_plain.py:27-123 re_ranking_lh: reid/rerank.py:27 re_ranking with another source term --
    v = min_j cdist(t_i, s_j) taken on the UN-squared float64 distances, no 1 - exp(), v /= max(v) in float64
    (:36-40), source_dist = v_i + v_j in float64 (:41-43); the Jaccard part (:46-113) is the same code.
    final = J * (1 - lambda) [storage dtype] + source_dist * lambda [float64] (:120)."""
    st = {}
    O.re_ranking(src, tgt, k1=k1, k2=k2, lambda_value=lambda_value, mode=mode, stages=st)
    v = cdist(tgt, src).min(axis=1)
    v = v / v.max()
    n = v.shape[0]
    source_dist = np.zeros([n, n])
    for i in range(n):
        source_dist[i, :] = v + v[i]
    return st["J"] * (1 - lambda_value) + source_dist * lambda_value
