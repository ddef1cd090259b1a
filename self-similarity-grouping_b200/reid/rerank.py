"""Drop-in for reid/rerank.py of the reference (same public names, positional order and defaults).

The reference module has no ``__all__``, and the driver does ``from sklearn.cluster import DBSCAN`` immediately
followed by ``from reid.rerank import *`` (selftraining.py:27-28): exporting ``DBSCAN`` from here makes the
unmodified driver cluster on the GPU as well.
"""
import numpy as np

from ssg_b200.rerank import re_ranking  # noqa: F401  (reid/rerank.py:27)
from ssg_b200.cluster import DBSCAN  # noqa: F401


def k_reciprocal_neigh(initial_rank, i, k1):
    """reid/rerank.py:165-169 (host helper kept for API compatibility; the CUDA path does this per row on the GPU)."""
    forward_k_neigh_index = initial_rank[i, :k1 + 1]
    backward_k_neigh_index = initial_rank[forward_k_neigh_index, :k1 + 1]
    fi = np.where(backward_k_neigh_index == i)[0]
    return forward_k_neigh_index[fi]


def re_ranking_init(query_feature, gallery_feature, k1=20, k2=6, lambda_value=0.3):
    """reid/rerank.py:171-234: cosine k-reciprocal re-ranking from features."""
    from .rerank_initial import re_ranking_init as _init
    import torch
    from ssg_b200 import _lib
    dev = _lib.require_cuda()
    q = torch.as_tensor(np.ascontiguousarray(query_feature, dtype=np.float32)).to(dev)
    g = torch.as_tensor(np.ascontiguousarray(gallery_feature, dtype=np.float32)).to(dev)
    # np.dot of the reference (rerank.py:174-176) on the library's own kernel (exact products, float64 sum)
    from ssg_b200.rerank import dot
    q_g, q_q, g_g = dot(q, g), dot(q, q), dot(g, g)
    return _init(q_g, q_q, g_g, k1=k1, k2=k2, lambda_value=lambda_value).cpu().numpy()
