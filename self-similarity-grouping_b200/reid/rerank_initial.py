"""Drop-in for reid/rerank_initial.py:40-99 re_ranking_init(q_g_dist, q_q_dist, g_g_dist, k1, k2, lambda_value)
(float32 k-reciprocal re-ranking on precomputed similarity blocks; imported by reid/eug.py:17)."""
import numpy as np


def k_reciprocal_neigh(initial_rank, i, k1):
    """reid/rerank_initial.py:34-38."""
    forward_k_neigh_index = initial_rank[i, :k1 + 1]
    backward_k_neigh_index = initial_rank[forward_k_neigh_index, :k1 + 1]
    fi = np.where(backward_k_neigh_index == i)[0]
    return forward_k_neigh_index[fi]


def re_ranking_init(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    from ssg_b200.rerank import re_ranking_init_blocks
    return re_ranking_init_blocks(q_g_dist, q_q_dist, g_g_dist, k1, k2, lambda_value)
