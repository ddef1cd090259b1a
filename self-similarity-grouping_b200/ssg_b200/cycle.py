"""The pseudo-label cycle of selftraining.py:189-222 as one device-resident pass.

``compute_dist`` / ``generate_selflabel`` / ``generate_keep_mask`` mirror the driver's own helper
functions (selftraining.py:255-277, 280-313, 316-323: same names, argument meaning and prints) but
keep the N x N matrices on the GPU between the two calls; ``pseudo_label_cycle`` chains them for
callers that only need the labels (which is all the driver uses the matrices for).
"""
import os

import numpy as np

from . import _lib
from .rerank import get_plan as _rerank_plan
from .cluster import get_plan as _cluster_plan, _with_capacity_retry


def _dist_mode(dist_mode):
    if dist_mode is None:
        return int(os.environ.get("SSG_DIST_MODE", _lib.DIST_TENSOR))
    return int(dist_mode)


def _to_device(x, dev):
    """numpy / CPU tensor / CUDA tensor -> contiguous float32 CUDA tensor (async when pinned)."""
    import torch
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not x.is_cuda:
        x = x.to(dev, dtype=torch.float32, non_blocking=True)
    return x.to(torch.float32).contiguous()


def compute_dist(source_features, target_features, lambda_value, no_rerank, num_split=2, dist_mode=None,
                 device=None):
    """selftraining.py:255-277.  Returns (euclidean_dist_list, rerank_dist_list); the re-ranked
    matrices are float64 CUDA tensors [N,N] (the reference returns host ndarrays), the Euclidean slot
    holds empty lists exactly as the reference does (selftraining.py:266,276)."""
    import torch
    dev = _lib.require_cuda(device)
    mode = _dist_mode(dist_mode)
    banks_s = source_features if isinstance(source_features, (list, tuple)) else [source_features]
    banks_t = target_features if isinstance(target_features, (list, tuple)) else [target_features]
    euclidean_dist_list, rerank_dist_list = [], []
    for s, t in zip(banks_s, banks_t):
        s, t = _to_device(s, dev), _to_device(t, dev)
        plan = _rerank_plan(t.shape[0], s.shape[0], t.shape[1], dev.index)
        if no_rerank:
            raise NotImplementedError("--no-rerank is broken in the reference itself (selftraining.py:266,285: "
                                      "the Euclidean list it clusters on holds empty lists)")
        _, final = plan.run(s, t, 20, 6, lambda_value, mode, want_euclid=False)
        rerank_dist_list.append(final)
        euclidean_dist_list.append([])
    return euclidean_dist_list, rerank_dist_list


class _Cluster(object):
    """What the driver caches in cluster_list (selftraining.py:296-298): eps frozen at iteration 0."""

    def __init__(self, eps, min_samples=4):
        self.eps, self.min_samples = eps, min_samples

    def fit_predict(self, dist):
        from .cluster import dbscan_labels
        return dbscan_labels(dist, self.eps, self.min_samples)


def generate_selflabel(e_dist, r_dist, n_iter, args, cluster_list=[]):
    """selftraining.py:280-313 on device-resident (or host) matrices.  args needs .rho and .no_rerank."""
    import torch
    labels_list = []
    for s in range(len(r_dist)):
        tmp_dist = e_dist[s] if getattr(args, "no_rerank", False) else r_dist[s]
        if n_iter == 0:
            if isinstance(tmp_dist, np.ndarray):
                eps = _cluster_plan(tmp_dist.shape[0]).eps_host(tmp_dist, args.rho)[0]
            else:
                eps = _cluster_plan(tmp_dist.shape[0], tmp_dist.device.index).eps(tmp_dist, args.rho)[0]
            print('eps in cluster: {:.3f}'.format(eps))
            cluster = _Cluster(eps, 4)
            cluster_list.append(cluster)
        else:
            cluster = cluster_list[s]
        print('Clustering and labeling...')
        labels = cluster.fit_predict(tmp_dist)
        if not isinstance(labels, np.ndarray):
            labels = labels.cpu().numpy()
        num_ids = len(set(labels)) - 1
        print('Iteration {} have {} training ids'.format(n_iter + 1, num_ids))
        labels_list.append(labels)
    return labels_list, cluster_list


def generate_keep_mask(labels_list):
    """selftraining.py:316-323: an image is kept iff no bank labelled it -1."""
    L = np.stack([np.asarray(l) for l in labels_list], 0)
    return ~(L == -1).any(0)


def _sparse_default():
    return os.environ.get("SSG_SPARSE_FINISH", "0") not in ("", "0")


def pseudo_label_cycle(source_features, target_features, lambda_value=0.1, rho=1.6e-3, eps_list=None,
                       min_samples=4, k1=20, k2=6, dist_mode=None, device=None, quiet=True, sparse=None):
    """Features (per bank: [Ns,d], [N,d]; host or device) -> (labels_list, eps_list, keep_mask).

    One bank at a time: upload, re-rank into a re-used [N,N] float64 device buffer, eps (unless frozen
    values are passed, as in iterations > 0), DBSCAN, download the labels.  Nothing N x N leaves the GPU.

    sparse=True (opt-in, SSG_SPARSE_FINISH=1): final_dist is never materialised.  Every entry outside the ~1 % of
    "touched" pairs is >= fl32(1 - lambda) (Jaccard distance 1), so eps (a rho-quantile far below that) and DBSCAN's
    region queries only need the touched entries, which RerankPlan.finish_sparse returns as a CSR.  The shortcut is
    taken only when it is certified (the rho-slice lies below the bound, eps < bound); otherwise the bank falls back
    to the dense matrix.  Labels are identical either way; eps agrees up to the order of the float64 additions.
    """
    import torch
    dev = _lib.require_cuda(device)
    mode = _dist_mode(dist_mode)
    banks_s = source_features if isinstance(source_features, (list, tuple)) else [source_features]
    banks_t = target_features if isinstance(target_features, (list, tuple)) else [target_features]
    labels_list, eps_out = [], []
    final = None
    for b, (s, t) in enumerate(zip(banks_s, banks_t)):
        s, t = _to_device(s, dev), _to_device(t, dev)
        n = t.shape[0]
        plan = _rerank_plan(n, s.shape[0], t.shape[1], dev.index)
        if (_sparse_default() if sparse is None else sparse) and 0.0 <= lambda_value < 1.0:
            plan.distance_rows(s, t, max(k1, k2 - 1), mode)      # rerank.py:97 reads k2 rank columns (as ssg_rerank_run)
            rowptr, col, val, bound = plan.finish_sparse(t, k1, k2, lambda_value)
            cplan = _cluster_plan(n, dev.index)
            if eps_list is None:
                eps, _, ok = cplan.eps_sparse(n, rowptr, col, val, bound, rho)
            else:
                eps, ok = float(eps_list[b]), True
            if ok and eps < bound:
                labels = _with_capacity_retry(n, dev.index,
                                              lambda p: p.dbscan_sparse(n, rowptr, col, val, eps, min_samples)[0])
                eps_out.append(eps)
                labels_list.append(labels)
                continue
            # not certified (the rho-slice or eps reaches the untouched entries): dense matrix for this bank; the tables
            # are already in the plan, ssg_rerank_finish repeats the cheap sparse stages and adds the N x N fill
            if final is None or final.shape[0] != n:
                final = torch.empty((n, n), dtype=torch.float64, device=dev)
            _lib.check(_lib.load().ssg_rerank_finish(plan._h, t.data_ptr(), n, t.shape[1], int(k1), int(k2),
                                                     float(lambda_value), final.data_ptr(), _lib.stream_ptr(dev)))
        else:
            if final is None or final.shape[0] != n:
                final = torch.empty((n, n), dtype=torch.float64, device=dev)
            plan.run(s, t, k1, k2, lambda_value, mode, want_euclid=False, out=final)
        cplan = _cluster_plan(n, dev.index)
        eps = cplan.eps(final, rho)[0] if eps_list is None else float(eps_list[b])
        labels = _with_capacity_retry(n, dev.index, lambda p: p.dbscan(final, eps, min_samples)[0])
        eps_out.append(eps)
        labels_list.append(labels)
    labels_host = [l.cpu().numpy() for l in labels_list]
    if not quiet:
        for b, l in enumerate(labels_host):
            print('bank {}: eps {:.3f}, {} training ids'.format(b, eps_out[b], len(set(l.tolist())) - 1))
    return labels_host, eps_out, generate_keep_mask(labels_host)
