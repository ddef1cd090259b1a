"""eps estimate + DBSCAN host API (selftraining.py:289-306 / sklearn.cluster.DBSCAN, precomputed)."""
import ctypes

import numpy as np

from . import _lib

_plans = {}


class ClusterPlan(object):
    def __init__(self, n_max, max_neighbors=0, device=None):
        dev = _lib.require_cuda(device)
        self.device = dev
        self.n_max = int(n_max)
        self.max_neighbors = int(max_neighbors)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.load().ssg_cluster_plan_create(ctypes.byref(self._h), dev.index, self.n_max,
                                                       self.max_neighbors))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().ssg_cluster_plan_destroy(h)
            except Exception:
                pass

    # ---- device matrices (torch CUDA tensors, float64 or float32, [n,n] contiguous)
    def _check_device(self, t):
        if not _lib.same_device(t, self.device):
            raise ValueError("ssg_b200: the matrix must live on the plan's device (%s), got %s" % (self.device, t.device))

    def eps(self, dist, rho):
        self._check_device(dist)
        dt = _dtype_code(dist)
        eps, top = ctypes.c_double(), ctypes.c_longlong()
        _lib.check(_lib.load().ssg_eps_estimate(self._h, dist.data_ptr(), dt, dist.shape[0], float(rho),
                                                ctypes.byref(eps), ctypes.byref(top), _lib.stream_ptr(self.device)))
        return eps.value, top.value

    def dbscan(self, dist, eps, min_samples=4):
        """-> (labels int64 CUDA tensor, number of clusters).  Synchronises the stream once (the library always reads
        the neighbour-capacity flag back; OverflowError -> _with_capacity_retry grows the plan)."""
        import torch
        self._check_device(dist)
        dt = _dtype_code(dist)
        n = dist.shape[0]
        labels = torch.empty((n,), dtype=torch.int64, device=dist.device)
        ncl = ctypes.c_int()
        _lib.check(_lib.load().ssg_dbscan(self._h, dist.data_ptr(), dt, n, float(eps), int(min_samples),
                                          labels.data_ptr(), ctypes.byref(ncl), _lib.stream_ptr(self.device)))
        return labels, ncl.value

    def core_mask(self, n):
        out = np.empty((n,), dtype=np.uint8)
        _lib.check(_lib.load().ssg_dbscan_core_mask(self._h, out.ctypes.data, n))
        return out.astype(bool)

    # ---- sparse form of final_dist (RerankPlan.finish_sparse)
    def eps_sparse(self, n, rowptr, col, val, threshold, rho):
        """-> (eps, top_num, certified).  certified False: the caller must use the dense matrix."""
        eps, top, ok = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_int()
        _lib.check(_lib.load().ssg_eps_sparse(self._h, int(n), rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                              float(threshold), float(rho), ctypes.byref(eps), ctypes.byref(top),
                                              ctypes.byref(ok), _lib.stream_ptr(self.device)))
        return eps.value, top.value, bool(ok.value)

    def dbscan_sparse(self, n, rowptr, col, val, eps, min_samples=4):
        import torch
        labels = torch.empty((n,), dtype=torch.int64, device=self.device)
        ncl = ctypes.c_int()
        _lib.check(_lib.load().ssg_dbscan_sparse(self._h, int(n), rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                                 float(eps), int(min_samples), labels.data_ptr(), ctypes.byref(ncl),
                                                 _lib.stream_ptr(self.device)))
        return labels, ncl.value

    # ---- row-sharded primitives (one process per GPU; the collectives between them are run by ssg_b200.dist)
    def buffers(self, n, nbr_len=0):
        """Zero-copy torch views of the plan's exchange buffers: hist int64[4096], state int64[8], partial
        float64[n], list float64[2^20], cnt int32[n], nbr int32[nbr_len]."""
        import torch
        ptrs = [ctypes.c_void_p() for _ in range(6)]
        _lib.check(_lib.load().ssg_cluster_buffers(self._h, *[ctypes.byref(p) for p in ptrs]))
        specs = [((_lib.EPS_BINS,), "<i8"), ((8,), "<i8"), ((n,), "<f8"), ((_lib.EPS_LIST_CAP,), "<f8"),
                 ((n,), "<i4"), ((max(int(nbr_len), 1),), "<i4")]
        names = ("hist", "state", "partial", "list", "cnt", "nbr")
        return {k: torch.as_tensor(_DevArray(p.value, shape, ts), device=self.device)
                for k, p, (shape, ts) in zip(names, ptrs, specs)}

    def eps_shard_begin(self):
        _lib.check(_lib.load().ssg_eps_shard_begin(self._h, _lib.stream_ptr(self.device)))

    def eps_shard_hist(self, rows, n, world, rank, npass):
        _lib.check(_lib.load().ssg_eps_shard_hist(self._h, _ptr(rows), _rows_dtype(rows, n), n, world, rank,
                                                  int(npass), _lib.stream_ptr(self.device)))

    def eps_shard_pick(self, npass, rho):
        _lib.check(_lib.load().ssg_eps_shard_pick(self._h, int(npass), float(rho), _lib.stream_ptr(self.device)))

    def eps_shard_gather(self, rows, n, world, rank, exact):
        cnt = ctypes.c_longlong()
        _lib.check(_lib.load().ssg_eps_shard_gather(self._h, _ptr(rows), _rows_dtype(rows, n), n, world, rank,
                                                    int(bool(exact)), None if exact else ctypes.byref(cnt),
                                                    _lib.stream_ptr(self.device)))
        return None if exact else cnt.value

    def eps_shard_finish(self, n, exact):
        eps, top = ctypes.c_double(), ctypes.c_longlong()
        _lib.check(_lib.load().ssg_eps_shard_finish(self._h, n, int(bool(exact)), ctypes.byref(eps), ctypes.byref(top),
                                                    _lib.stream_ptr(self.device)))
        return eps.value, top.value

    def dbscan_shard_count(self, rows, n, row0, eps):
        _lib.check(_lib.load().ssg_dbscan_shard_count(self._h, _ptr(rows), _rows_dtype(rows, n), n, int(row0),
                                                      rows.shape[0], float(eps), _lib.stream_ptr(self.device)))

    def dbscan_shard_fill(self, rows, n, row0, eps):
        total = ctypes.c_longlong()
        _lib.check(_lib.load().ssg_dbscan_shard_fill(self._h, _ptr(rows), _rows_dtype(rows, n), n, int(row0),
                                                     rows.shape[0], float(eps), ctypes.byref(total),
                                                     _lib.stream_ptr(self.device)))
        return total.value

    def dbscan_shard_label(self, n, min_samples=4):
        import torch
        labels = torch.empty((n,), dtype=torch.int64, device=self.device)
        ncl = ctypes.c_int()
        _lib.check(_lib.load().ssg_dbscan_shard_label(self._h, n, int(min_samples), labels.data_ptr(),
                                                      ctypes.byref(ncl), _lib.stream_ptr(self.device)))
        return labels, ncl.value

    # ---- host matrices (numpy)
    def eps_host(self, dist, rho):
        dist, dt = _host_matrix(dist)
        eps, top = ctypes.c_double(), ctypes.c_longlong()
        _lib.check(_lib.load().ssg_eps_estimate_host(self._h, dist.ctypes.data, dt, dist.shape[0], float(rho),
                                                     ctypes.byref(eps), ctypes.byref(top)))
        return eps.value, top.value

    def dbscan_host(self, dist, eps, min_samples=4):
        dist, dt = _host_matrix(dist)
        n = dist.shape[0]
        labels = np.empty((n,), dtype=np.int64)
        ncl = ctypes.c_int()
        _lib.check(_lib.load().ssg_dbscan_host(self._h, dist.ctypes.data, dt, n, float(eps), int(min_samples),
                                               labels.ctypes.data, ctypes.byref(ncl)))
        return labels, ncl.value


class _DevArray(object):
    """Minimal __cuda_array_interface__ holder for a raw device pointer owned by a plan."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _ptr(rows):
    return rows.data_ptr() if rows.shape[0] else None


def _rows_dtype(rows, n):
    """A [rows, n] contiguous CUDA block of a float64 / float32 matrix -> SSG_F64 / SSG_F32."""
    import torch
    assert rows.is_cuda and rows.dim() == 2 and rows.shape[1] == n and rows.is_contiguous()
    if rows.dtype == torch.float64:
        return _lib.F64
    if rows.dtype == torch.float32:
        return _lib.F32
    raise ValueError("distance rows must be float64 or float32")


def _dtype_code(t):
    import torch
    assert t.is_cuda and t.dim() == 2 and t.shape[0] == t.shape[1] and t.is_contiguous()
    if t.dtype == torch.float64:
        return _lib.F64
    if t.dtype == torch.float32:
        return _lib.F32
    raise ValueError("distance matrix must be float64 or float32")


def _host_matrix(dist):
    dist = np.asarray(dist)
    if dist.ndim != 2 or dist.shape[0] != dist.shape[1]:
        raise ValueError("precomputed distance matrix must be square")
    if dist.dtype == np.float32:
        return np.ascontiguousarray(dist), _lib.F32
    # float16 / ints / float64: float64 holds all of them exactly
    return np.ascontiguousarray(dist, dtype=np.float64), _lib.F64


def get_plan(n, device=None, max_neighbors=0):
    dev = _lib.require_cuda(device)
    plan = _plans.get(dev.index)
    if plan is None or plan.n_max < n or (max_neighbors and plan.max_neighbors < max_neighbors):
        _plans.pop(dev.index, None)
        plan = ClusterPlan(max(n, plan.n_max if plan else 0), max_neighbors, dev.index)
        _plans[dev.index] = plan
    return plan


def _with_capacity_retry(n, device, fn):
    """Run fn(plan); on neighbour-list overflow re-create the plan with more room (up to n*n)."""
    cap = 0
    while True:
        plan = get_plan(n, device, cap)
        try:
            return fn(plan)
        except OverflowError:
            cur = plan.max_neighbors or (64 * plan.n_max + (1 << 20))
            if cur >= n * n:
                raise
            cap = min(n * n, cur * 8)


def eps_estimate(dist, rho):
    """selftraining.py:289-293 on a numpy matrix or a CUDA tensor."""
    if isinstance(dist, np.ndarray):
        return get_plan(dist.shape[0]).eps_host(dist, rho)[0]
    return get_plan(dist.shape[0], dist.device.index).eps(dist, rho)[0]


def dbscan_labels(dist, eps, min_samples=4):
    """Labels as sklearn's DBSCAN(metric='precomputed').fit_predict would give them."""
    if isinstance(dist, np.ndarray):
        return _with_capacity_retry(dist.shape[0], None, lambda p: p.dbscan_host(dist, eps, min_samples)[0])
    return _with_capacity_retry(dist.shape[0], dist.device.index,
                                lambda p: p.dbscan(dist, eps, min_samples)[0])


class DBSCAN(object):
    """GPU stand-in for sklearn.cluster.DBSCAN as the reference uses it (selftraining.py:295-306):
    ``DBSCAN(eps=eps, min_samples=4, metric='precomputed', n_jobs=8).fit_predict(dist)``.
    The estimator is re-usable across calls (the reference caches it in ``cluster_list``).
    Only ``metric='precomputed'`` on a dense matrix is supported."""

    def __init__(self, eps=0.5, min_samples=5, metric='euclidean', metric_params=None, algorithm='auto',
                 leaf_size=30, p=None, n_jobs=None):
        self.eps = eps
        self.min_samples = min_samples
        self.metric = metric
        self.metric_params = metric_params
        self.algorithm = algorithm
        self.leaf_size = leaf_size
        self.p = p
        self.n_jobs = n_jobs

    def fit(self, X, y=None, sample_weight=None):
        if self.metric != 'precomputed':
            raise ValueError("ssg_b200.DBSCAN supports metric='precomputed' only")
        if sample_weight is not None:
            raise ValueError("ssg_b200.DBSCAN does not support sample_weight")
        eps = float(self.eps)
        # sklearn validates at fit(): eps is a real in (0, inf) -- NaN (an empty rho-slice, selftraining.py:293) and
        # 0 are rejected with InvalidParameterError, a ValueError -- and min_samples an integer >= 1
        if not (eps > 0.0) or eps == float("inf"):
            raise ValueError("The 'eps' parameter of DBSCAN must be a float in the range (0.0, inf). Got %r instead."
                             % (self.eps,))
        if int(self.min_samples) != self.min_samples or self.min_samples < 1:
            raise ValueError("The 'min_samples' parameter of DBSCAN must be an int in the range [1, inf). Got %r "
                             "instead." % (self.min_samples,))
        if isinstance(X, np.ndarray) or not hasattr(X, "is_cuda"):
            X = np.asarray(X)
            if X.dtype == np.float32 and not isinstance(self.eps, np.floating):
                eps = float(np.float32(eps))    # numpy compares a float32 matrix with a Python float in float32
            n = X.shape[0]
            labels = _with_capacity_retry(n, None, lambda p: p.dbscan_host(X, eps, self.min_samples)[0])
            core = get_plan(n).core_mask(n)
        else:
            n = X.shape[0]
            labels = _with_capacity_retry(n, X.device.index,
                                          lambda p: p.dbscan(X, eps, self.min_samples)[0]).cpu().numpy()
            core = get_plan(n, X.device.index).core_mask(n)
        self.labels_ = labels
        self.core_sample_indices_ = np.where(core)[0]
        self.n_features_in_ = n
        return self

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_

    def get_params(self, deep=True):
        return dict(eps=self.eps, min_samples=self.min_samples, metric=self.metric,
                    metric_params=self.metric_params, algorithm=self.algorithm, leaf_size=self.leaf_size,
                    p=self.p, n_jobs=self.n_jobs)

    def set_params(self, **params):
        for k, v in params.items():
            setattr(self, k, v)
        return self
