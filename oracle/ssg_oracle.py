"""CPU restatement of the SSG pseudo-label hot path (numpy).  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it restates (paths relative to /root/reference).  The
restatement exists so that parity tests can run where /root/reference does not (the GPU box) and so
that single stages can be fed identical inputs on both sides.  It is pinned bit-for-bit against the
unmodified reference executed in the build container (tests/test_oracle_vs_reference.py, and the
committed vectors in tests/golden/ made by oracle/make_goldens.py).

Two arithmetic modes:
  * ``mode='f32'`` — O-f32: every float16 of reid/rerank.py becomes float32 and argsort is stable.
                     This is the parity target of the CUDA path (1e-4 / bit-exact labels).
  * ``mode='ref'`` — O-ref: float16 storage and numpy's default (unstable) argsort, as shipped.
"""
import numpy as np
from scipy.spatial.distance import cdist


def _dt(mode):
    return np.float32 if mode == "f32" else np.float16


# ----------------------------------------------------------------------------- re_ranking stages
def source_vector(tgt, src, mode="f32"):
    """reid/rerank.py:36-40 — v_i = min_j(1-exp(-||t_i-s_j||^2)), then v / max(v)."""
    dt = _dt(mode)
    sour_tar = np.power(cdist(tgt, src), 2).astype(dt)
    sour_tar = 1 - np.exp(-sour_tar)
    vec = np.min(sour_tar, axis=1)
    return vec / np.max(vec)


def original_distance(tgt, mode="f32"):
    """reid/rerank.py:33,61-62 — squared Euclidean distance of the (quantised) target features."""
    dt = _dt(mode)
    feat = tgt.astype(dt)
    od = cdist(feat, feat).astype(dt)
    return np.power(od, 2).astype(dt)


def normalise(od):
    """reid/rerank.py:68 — divide by the column max and transpose (== row-normalise; od symmetric)."""
    return np.transpose(od / np.max(od, axis=0))


def initial_rank(odn, mode="f32"):
    """reid/rerank.py:70 — full argsort of every row (stable in O-f32)."""
    if mode == "f32":
        return np.argsort(odn, kind="stable").astype(np.int32)
    return np.argsort(odn).astype(np.int32)


def k_reciprocal_neigh(rank, i, k):
    """reid/rerank.py:165-169."""
    fwd = rank[i, :k + 1]
    bwd = rank[fwd, :k + 1]
    fi = np.where(bwd == i)[0]
    return fwd[fi]


def k_reciprocal_rows(odn, rank, k1=20):
    """reid/rerank.py:74-92 — returns the sparse rows of V as lists (idx ascending, weights)."""
    n = odn.shape[0]
    k_half = int(np.around(k1 / 2))
    idx_rows, val_rows = [], []
    for i in range(n):
        r = k_reciprocal_neigh(rank, i, k1)
        exp_idx = r
        for c in r:
            rc = k_reciprocal_neigh(rank, c, k_half)
            if len(np.intersect1d(rc, r)) > 2 / 3 * len(rc):
                exp_idx = np.append(exp_idx, rc)
        exp_idx = np.unique(exp_idx)
        w = np.exp(-odn[i, exp_idx])
        idx_rows.append(exp_idx.astype(np.int64))
        val_rows.append(w / np.sum(w))
    return idx_rows, val_rows


def dense_V(idx_rows, val_rows, n, dtype):
    V = np.zeros((n, n), dtype=dtype)
    for i, (ix, vv) in enumerate(zip(idx_rows, val_rows)):
        V[i, ix] = vv
    return V


def query_expand(V, rank, k2=6):
    """reid/rerank.py:94-98."""
    if k2 == 1:
        return V
    V_qe = np.zeros_like(V)
    for i in range(V.shape[0]):
        V_qe[i, :] = np.mean(V[rank[i, :k2], :], axis=0)
    return V_qe


def jaccard(V, rows=None):
    """reid/rerank.py:101-118 — Jaccard distance through the inverted index (rows: first `rows`)."""
    n = V.shape[0]
    rows = n if rows is None else rows
    inv = [np.where(V[:, k] != 0)[0] for k in range(n)]
    J = np.zeros((rows, n), dtype=V.dtype)
    two = V.dtype.type(2)
    for i in range(rows):
        temp_min = np.zeros((1, n), dtype=V.dtype)
        nz = np.where(V[i, :] != 0)[0]
        for k in nz:
            temp_min[0, inv[k]] = temp_min[0, inv[k]] + np.minimum(V[i, k], V[inv[k], k])
        J[i] = 1 - temp_min / (two - temp_min)
    J[J < 0] = 0.0
    return J


def final_distance(J, vec, lambda_value):
    """reid/rerank.py:41-43,122 — final = J*(1-lambda) + (v_i + v_j)*lambda (float64 result)."""
    n = J.shape[0]
    source_dist = np.zeros([n, n])
    for i in range(n):
        source_dist[i, :] = vec + vec[i]
    return J * (1 - lambda_value) + source_dist * lambda_value


def re_ranking(src, tgt, k1=20, k2=6, lambda_value=0.2, no_rerank=False, mode="f32", stages=None):
    """reid/rerank.py:27-127.  Returns (euclidean_dist, final_dist); fills `stages` if a dict."""
    dt = _dt(mode)
    vec = source_vector(tgt, src, mode)
    od = original_distance(tgt, mode)
    euclidean = od
    if no_rerank:
        return euclidean, None
    odn = normalise(od)
    rank = initial_rank(odn, mode)
    idx_rows, val_rows = k_reciprocal_rows(odn, rank, k1)
    V = dense_V(idx_rows, val_rows, od.shape[0], dt)
    Vq = query_expand(V, rank, k2)
    J = jaccard(Vq)
    final = final_distance(J, vec, lambda_value)
    if stages is not None:
        stages.update(vec=vec, od=od, odn=odn, rank=rank, V=V, Vq=Vq, J=J)
    return euclidean, final


# ----------------------------------------------------------------------------- re_ranking_init
def re_ranking_init(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """reid/rerank_initial.py:40-99 (float32; inputs are similarities; stable full argsort stands in
    for argpartition, whose tie order is unspecified)."""
    od = np.concatenate([np.concatenate([q_q_dist, q_g_dist], axis=1),
                         np.concatenate([q_g_dist.T, g_g_dist], axis=1)], axis=0)
    od = 2. - 2 * od
    od = np.transpose(1. * od / np.max(od, axis=0))
    rank = np.argsort(od, kind="stable")
    qn = q_g_dist.shape[0]
    n = od.shape[0]
    idx_rows, val_rows = k_reciprocal_rows(od, rank, k1)
    V = dense_V(idx_rows, val_rows, n, np.float32)
    Vq = query_expand(V, rank, k2)
    inv = [np.where(Vq[:, k] != 0)[0] for k in range(n)]
    J = np.zeros((qn, n), dtype=np.float32)
    for i in range(qn):
        temp_min = np.zeros((1, n), dtype=np.float32)
        nz = np.where(Vq[i, :] != 0)[0]
        for k in nz:
            temp_min[0, inv[k]] = temp_min[0, inv[k]] + np.minimum(Vq[i, k], Vq[inv[k], k])
        J[i] = 1 - temp_min / (2. - temp_min)
    final = J * (1 - lambda_value) + od[:qn] * lambda_value
    return final[:qn, qn:]


def re_ranking_init_features(query_feature, gallery_feature, **kw):
    """reid/rerank.py:171-234 — cosine variant taking features."""
    q_g = np.dot(query_feature, gallery_feature.T)
    q_q = np.dot(query_feature, query_feature.T)
    g_g = np.dot(gallery_feature, gallery_feature.T)
    return re_ranking_init(q_g, q_q, g_g, **kw)


# ----------------------------------------------------------------------------- eps + DBSCAN
def eps_estimate(dist, rho):
    """selftraining.py:289-293 — mean of the round(rho*M) smallest non-zero upper-triangle entries."""
    tri = np.triu(dist, 1)
    tri = tri[np.nonzero(tri)]
    tri = np.sort(tri, axis=None)
    top_num = np.round(rho * tri.size).astype(int)
    return tri[:top_num].mean()


def dbscan_dfs(dist, eps, min_samples=4):
    """sklearn 1.9 DBSCAN(metric='precomputed') on a dense matrix, restated:
    neighbours = np.where(d <= eps) per row (sklearn/neighbors/_base.py:1074), core = count >=
    min_samples (cluster/_dbscan.py:452-462), labelling = dbscan_inner stack DFS in index order
    (cluster/_dbscan_inner.pyx:19-41).  Call sites: selftraining.py:295,306."""
    n = dist.shape[0]
    neigh = [np.where(row <= eps)[0] for row in dist]
    core = np.array([len(x) >= min_samples for x in neigh])
    labels = np.full(n, -1, dtype=np.int64)
    label_num = 0
    for i in range(n):
        if labels[i] != -1 or not core[i]:
            continue
        stack = []
        while True:
            if labels[i] == -1:
                labels[i] = label_num
                if core[i]:
                    for v in neigh[i]:
                        if labels[v] == -1:
                            stack.append(v)
            if not stack:
                break
            i = stack.pop()
        label_num += 1
    return labels


def dbscan_components(dist, eps, min_samples=4):
    """Order-free restatement used by the CUDA path (requires symmetric `dist`):
    components of the core graph, ids ranked by minimum core index, border = min id over core
    neighbours, noise = -1."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    A = dist <= eps
    core = A.sum(1) >= min_samples
    n = dist.shape[0]
    labels = np.full(n, -1, dtype=np.int64)
    ci = np.where(core)[0]
    if len(ci) == 0:
        return labels
    _, comp = connected_components(csr_matrix(A[np.ix_(ci, ci)]), directed=False)
    first = {}
    for pos, c in enumerate(comp):
        first.setdefault(c, pos)
    order = sorted(first, key=lambda c: first[c])
    remap = {c: r for r, c in enumerate(order)}
    labels[ci] = [remap[c] for c in comp]
    for i in np.where(~core)[0]:
        nb = np.where(A[i] & core)[0]
        if len(nb):
            labels[i] = labels[nb].min()
    return labels


def generate_selflabel(r_dist, rho, eps_list=None, min_samples=4):
    """selftraining.py:280-313 (rerank branch): eps at iteration 0 per bank, then DBSCAN labels."""
    labels_list, eps_out = [], []
    for s, D in enumerate(r_dist):
        eps = eps_estimate(D, rho) if eps_list is None else eps_list[s]
        eps_out.append(eps)
        labels_list.append(dbscan_dfs(D, eps, min_samples))
    return labels_list, eps_out


def keep_mask(labels_list):
    """selftraining.py:316-323 — an image is kept iff no bank labelled it -1."""
    L = np.stack(labels_list, 0)
    return ~(L == -1).any(0)


# ----------------------------------------------------------------------------- evaluators
def pairwise_distance(x, y=None):
    """reid/evaluators.py:63-85 on stacked feature matrices (float32 torch semantics in numpy)."""
    x = np.asarray(x, np.float32)
    if y is None:
        d = (x * x).sum(1, keepdims=True) * 2
        return np.broadcast_to(d, (x.shape[0], x.shape[0])) - 2 * (x @ x.T)
    y = np.asarray(y, np.float32)
    return (x * x).sum(1, keepdims=True) + (y * y).sum(1, keepdims=True).T - 2 * (x @ y.T)


# ----------------------------------------------------------------------------- synthetic inputs
def synth_features(n, d=2048, seed=0, per_cluster=20, noise=0.5):
    """SURVEY.md §8d feature-level generator: n/per_cluster Gaussian centres + noise, L2-normalised."""
    rng = np.random.RandomState(seed)
    c = max(n // per_cluster, 1)
    centres = rng.randn(c, d)
    lab = rng.randint(0, c, n)
    f = centres[lab] + noise * rng.randn(n, d)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    return f.astype(np.float32), lab


# ----------------------------------------------------------------------------- evaluation metrics (row f2)
def cmc(distmat, query_ids, gallery_ids, query_cams, gallery_cams, topk=100, first_match_break=False,
        separate_camera_set=False):
    """reid/evaluation_metrics/ranking.py:18-79 (single_gallery_shot=False); ties ranked by gallery index."""
    distmat = np.asarray(distmat)
    m, n = distmat.shape
    query_ids, gallery_ids = np.asarray(query_ids), np.asarray(gallery_ids)
    query_cams, gallery_cams = np.asarray(query_cams), np.asarray(gallery_cams)
    indices = np.argsort(distmat, axis=1, kind="stable")
    matches = (gallery_ids[indices] == query_ids[:, np.newaxis])
    ret = np.zeros(topk)
    num_valid_queries = 0
    for i in range(m):
        valid = ((gallery_ids[indices[i]] != query_ids[i]) | (gallery_cams[indices[i]] != query_cams[i]))
        if separate_camera_set:
            valid &= (gallery_cams[indices[i]] != query_cams[i])
        if not np.any(matches[i, valid]):
            continue
        index = np.nonzero(matches[i, valid])[0]
        delta = 1. / len(index)
        for j, k in enumerate(index):
            if k - j >= topk:
                break
            if first_match_break:
                ret[k - j] += 1
                break
            ret[k - j] += delta
        num_valid_queries += 1
    if num_valid_queries == 0:
        raise RuntimeError("No valid query")
    return ret.cumsum() / num_valid_queries


def mean_ap(distmat, query_ids, gallery_ids, query_cams, gallery_cams):
    """reid/evaluation_metrics/ranking.py:82-115 (sklearn average_precision_score per query)."""
    from sklearn.metrics import average_precision_score
    distmat = np.asarray(distmat)
    m, n = distmat.shape
    query_ids, gallery_ids = np.asarray(query_ids), np.asarray(gallery_ids)
    query_cams, gallery_cams = np.asarray(query_cams), np.asarray(gallery_cams)
    indices = np.argsort(distmat, axis=1, kind="stable")
    matches = (gallery_ids[indices] == query_ids[:, np.newaxis])
    aps = []
    for i in range(m):
        valid = ((gallery_ids[indices[i]] != query_ids[i]) | (gallery_cams[indices[i]] != query_cams[i]))
        y_true = matches[i, valid]
        y_score = -distmat[i][indices[i]][valid]
        if not np.any(y_true):
            continue
        aps.append(average_precision_score(y_true, y_score))
    if len(aps) == 0:
        raise RuntimeError("No valid query")
    return np.mean(aps)
