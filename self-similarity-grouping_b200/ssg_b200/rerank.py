"""Re-ranking host API (reid/rerank.py:27-127 of the reference) on top of the C ABI."""
import ctypes

import numpy as np

from . import _lib

_plans = {}


class RerankPlan(object):
    """Owns the device workspace for one (n_max, ns_max, d) problem size on one GPU."""

    def __init__(self, n_max, ns_max, d, device=None):
        dev = _lib.require_cuda(device)
        self.device = dev
        self.n_max, self.ns_max, self.d = int(n_max), int(ns_max), int(d)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.load().ssg_rerank_plan_create(ctypes.byref(self._h), dev.index, self.n_max,
                                                      self.ns_max, self.d))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().ssg_rerank_plan_destroy(h)
            except Exception:
                pass

    @property
    def nbytes(self):
        return int(_lib.load().ssg_rerank_plan_bytes(self._h))

    def run(self, src, tgt, k1=20, k2=6, lambda_value=0.2, dist_mode=_lib.DIST_EXACT, want_euclid=False,
            out=None):
        """src [ns,d], tgt [n,d]: float32 CUDA tensors.  Returns (euclid or None, final) CUDA tensors
        (float32 [n,n], float64 [n,n]); asynchronous on the current stream."""
        import torch
        if not (_lib.same_device(src, self.device) and _lib.same_device(tgt, self.device)):
            raise ValueError("ssg_b200: features must live on the plan's device (%s), got %s / %s"
                             % (self.device, src.device, tgt.device))
        assert src.dtype == torch.float32 and tgt.dtype == torch.float32
        src, tgt = src.contiguous(), tgt.contiguous()
        n, d = tgt.shape
        ns = src.shape[0]
        final = out if out is not None else torch.empty((n, n), dtype=torch.float64, device=tgt.device)
        euclid = torch.empty((n, n), dtype=torch.float32, device=tgt.device) if want_euclid else None
        _lib.check(_lib.load().ssg_rerank_run(
            self._h, src.data_ptr(), ns, tgt.data_ptr(), n, d, int(k1), int(k2), float(lambda_value),
            int(dist_mode), final.data_ptr(), euclid.data_ptr() if want_euclid else None, _lib.stream_ptr(self.device)))
        return euclid, final

    def run_host(self, src, tgt, k1=20, k2=6, lambda_value=0.2, dist_mode=_lib.DIST_EXACT, no_rerank=False,
                 want_euclid=True):
        """numpy in / numpy out through ssg_rerank_host (pinned result buffers)."""
        src = np.ascontiguousarray(src, dtype=np.float32)
        tgt = np.ascontiguousarray(tgt, dtype=np.float32)
        n, d = tgt.shape
        ns = src.shape[0]
        final = None if no_rerank else _pinned((n, n), np.float64)
        euclid = _pinned((n, n), np.float32) if want_euclid else None
        _lib.check(_lib.load().ssg_rerank_host(
            self._h, src.ctypes.data, ns, tgt.ctypes.data, n, d, int(k1), int(k2), float(lambda_value),
            int(dist_mode), int(bool(no_rerank)), final.ctypes.data if final is not None else None,
            euclid.ctypes.data if euclid is not None else None))
        return euclid, final

    def distance_rows(self, src, tgt, k1=20, dist_mode=_lib.DIST_EXACT, row0=0, rows=None):
        """Stages (i)-(iv) for target rows [row0, row0+rows) (all rows by default): fills the plan's tables."""
        src, tgt = src.contiguous(), tgt.contiguous()
        rows = tgt.shape[0] - row0 if rows is None else rows
        _lib.check(_lib.load().ssg_rerank_distance_rows(self._h, src.data_ptr(), src.shape[0], tgt.data_ptr(),
                                                        tgt.shape[0], tgt.shape[1], int(k1), int(dist_mode), int(row0),
                                                        int(rows), None, _lib.stream_ptr(self.device)))

    def finish_sparse(self, tgt, k1=20, k2=6, lambda_value=0.2):
        """final_dist as a CSR over the touched columns (see include/ssg_b200.h).  Returns (rowptr int32 [n+1],
        col int32 [nnz], val float64 [nnz], threshold): zero-copy views of plan-owned buffers, valid until the next
        call on this plan."""
        import torch
        tgt = tgt.contiguous()
        n = tgt.shape[0]
        nnz = ctypes.c_longlong()
        _lib.check(_lib.load().ssg_rerank_finish_sparse(self._h, tgt.data_ptr(), n, tgt.shape[1], int(k1), int(k2),
                                                        float(lambda_value), ctypes.byref(nnz), _lib.stream_ptr(self.device)))
        ptrs = [ctypes.c_void_p() for _ in range(3)]
        cnt, thr = ctypes.c_longlong(), ctypes.c_double()
        _lib.check(_lib.load().ssg_rerank_sparse_view(self._h, ctypes.byref(ptrs[0]), ctypes.byref(ptrs[1]),
                                                      ctypes.byref(ptrs[2]), ctypes.byref(cnt), ctypes.byref(thr)))
        from .cluster import _DevArray
        m = max(int(cnt.value), 1)
        views = [torch.as_tensor(_DevArray(p.value, shape, ts), device=self.device)
                 for p, (shape, ts) in zip(ptrs, [((n + 1,), "<i4"), ((m,), "<i4"), ((m,), "<f8")])]
        return views[0], views[1][: cnt.value], views[2][: cnt.value], thr.value

    def stage(self, which, n):
        """Copy an intermediate of the last run to the host (stage-isolated parity tests)."""
        shapes = {
            _lib.STAGE_VEC: ((n,), np.float32), _lib.STAGE_ROWMAX: ((n,), np.float32),
            _lib.STAGE_RANK: ((n, _lib.RANK_STRIDE), np.int32),
            _lib.STAGE_RANK_VAL: ((n, _lib.RANK_STRIDE), np.float32),
            _lib.STAGE_V_CNT: ((n,), np.int32), _lib.STAGE_V_IDX: ((n, _lib.V_STRIDE), np.int32),
            _lib.STAGE_V_VAL: ((n, _lib.V_STRIDE), np.float32),
            _lib.STAGE_VQ_CNT: ((n,), np.int32), _lib.STAGE_VQ_IDX: ((n, _lib.VQ_STRIDE), np.int32),
            _lib.STAGE_VQ_VAL: ((n, _lib.VQ_STRIDE), np.float32),
            _lib.STAGE_FLAGGED: ((1,), np.int32),
        }
        shape, dt = shapes[which]
        out = np.empty(shape, dtype=dt)
        _lib.check(_lib.load().ssg_rerank_get_stage(self._h, which, out.ctypes.data, out.nbytes))
        return out


class _PinnedPool(object):
    """Pinned host result buffers, recycled.  Page-locking an N x N float64 matrix costs more than copying it (2.2 GB at
    N = 16 702: ~0.3 s of cudaHostAlloc against 0.04 s of PCIe), and the drop-in re_ranking hands a fresh matrix to the
    caller for every bank and iteration (selftraining.py:259-276).  A buffer goes back to the pool when the ndarray that
    was returned to the caller (and every view of it) has been garbage collected -- never earlier, so a caller that keeps
    the matrices of all banks alive simply gets distinct buffers."""

    def __init__(self, keep_bytes=8 << 30):
        self.free = {}
        self.keep_bytes = keep_bytes
        self.held = 0

    def _release(self, key, tensor):
        nbytes = tensor.numel() * tensor.element_size()
        if self.held + nbytes <= self.keep_bytes:
            self.free.setdefault(key, []).append(tensor)
            self.held += nbytes

    def get(self, shape, dtype):
        import weakref
        import torch
        tdt = {np.float64: torch.float64, np.float32: torch.float32, np.int64: torch.int64}[dtype]
        key = (tuple(shape), tdt)
        lst = self.free.get(key)
        if lst:
            t = lst.pop()
            self.held -= t.numel() * t.element_size()
        else:
            t = torch.empty(shape, dtype=tdt, pin_memory=True)
        arr = t.numpy()
        weakref.finalize(arr, self._release, key, t)
        return arr


_pool = _PinnedPool()


def _pinned(shape, dtype):
    """A numpy array backed by pinned host memory (fast D2H); recycled through _PinnedPool once the caller drops it."""
    return _pool.get(shape, dtype)


def get_plan(n, ns, d, device=None):
    dev = _lib.require_cuda(device)
    key = (dev.index, d)
    plan = _plans.get(key)
    if plan is None or plan.n_max < n or plan.ns_max < ns:
        _plans.pop(key, None)
        plan = RerankPlan(max(n, plan.n_max if plan else 0), max(ns, plan.ns_max if plan else 0), d, dev.index)
        _plans[key] = plan
    return plan


def re_ranking_device(src, tgt, k1=20, k2=6, lambda_value=0.2, dist_mode=_lib.DIST_EXACT, want_euclid=False):
    """Device-resident variant: CUDA float32 tensors in, (euclid|None, final float64) CUDA tensors out."""
    plan = get_plan(tgt.shape[0], src.shape[0], tgt.shape[1], tgt.device.index)
    return plan.run(src, tgt, k1, k2, lambda_value, dist_mode, want_euclid)


def re_ranking(input_feature_source, input_feature, k1=20, k2=6, lambda_value=0.2, MemorySave=False,
               Minibatch=2000, no_rerank=False, dist_mode=None):
    """Drop-in for reid/rerank.py:27 re_ranking (same positional order and defaults).

    Returns (euclidean_dist, final_dist) as numpy arrays — float32 [N,N] and float64 [N,N]
    (``final_dist`` is None when ``no_rerank``).  MemorySave/Minibatch only chunk the reference's cdist and have no
    effect on results, so they are accepted and ignored.

    Deviation from the reference, by design: the reference stores its intermediates
    in float16 (rerank.py:33-70) and ranks with numpy's unstable argsort, so its exact bits depend on the host CPU's
    float16 ``exp`` and sort kernels and cannot be reproduced anywhere else.  This function computes the same
    pipeline with float16 replaced by float32 and ties broken by (value, index) -- the "O-f32" oracle -- and
    ``euclidean_dist`` comes back as float32 instead of float16.  On tie-heavy float16 data eps and labels can
    therefore differ from a run of the reference itself (as two runs of the reference on different CPUs can).

    dist_mode: one rule for the whole package -- ``SSG_DIST_MODE`` when set, else the exact float64 distances
    (``DIST_EXACT``) for calls that RETURN the Euclidean matrix (this function: the matrix is then bit-identical to
    scipy's cdist squared) and the tensor-core distances with exact re-scoring (``DIST_TENSOR``) for calls that
    only return ``final_dist`` / labels (``compute_dist``, ``pseudo_label_cycle``).  ``final_dist`` is bit-identical
    in both modes (tests/test_gpu_tensor.py); only the returned Euclidean matrix differs (|err| <= 4.1e-5).
    """
    import os
    if dist_mode is None:
        dist_mode = int(os.environ.get("SSG_DIST_MODE", _lib.DIST_EXACT))
    src = np.ascontiguousarray(input_feature_source, dtype=np.float32)
    tgt = np.ascontiguousarray(input_feature, dtype=np.float32)
    plan = get_plan(tgt.shape[0], src.shape[0], tgt.shape[1])
    print('computing source distance...')
    print('computing original distance...')
    if not no_rerank:
        print('starting re_ranking...')
    return plan.run_host(src, tgt, k1, k2, lambda_value, dist_mode, no_rerank, want_euclid=True)


def re_ranking_plain(input_feature_source, input_feature, k=20, lambda_value=0.1, MemorySave=False, Minibatch=2000,
                     dist_mode=None):
    """Drop-in for reid/rerank_plain.py:125 re_ranking (the kNN-set Jaccard variant): same
    positional order and defaults, returns ``(final_dist, final_dist)`` as the reference does (float64 [N,N]).
    MemorySave / Minibatch only chunk the reference's cdist and are ignored."""
    import os
    import torch
    dev = _lib.require_cuda()
    if dist_mode is None:
        dist_mode = int(os.environ.get("SSG_DIST_MODE", _lib.DIST_EXACT))
    src = torch.from_numpy(np.ascontiguousarray(input_feature_source, dtype=np.float32)).to(dev)
    tgt = torch.from_numpy(np.ascontiguousarray(input_feature, dtype=np.float32)).to(dev)
    n, d = tgt.shape
    plan = get_plan(n, src.shape[0], d, dev.index)
    print('computing source distance...')
    print('computing original distance...')
    final = torch.empty((n, n), dtype=torch.float64, device=dev)
    _lib.check(_lib.load().ssg_rerank_plain(plan._h, src.data_ptr(), src.shape[0], tgt.data_ptr(), n, d, int(k),
                                            float(lambda_value), int(dist_mode), final.data_ptr(), _lib.stream_ptr(dev)))
    out = final.cpu().numpy()
    return out, out


def re_ranking_lh(input_feature_source, input_feature, k1=20, k2=6, lambda_value=0.2, MemorySave=False, Minibatch=2000,
                  dist_mode=None):
    """Drop-in for reid/rerank_plain.py:27 re_ranking_lh: returns (euclidean_dist float32, final_dist float64)."""
    import os
    import torch
    dev = _lib.require_cuda()
    if dist_mode is None:
        dist_mode = int(os.environ.get("SSG_DIST_MODE", _lib.DIST_EXACT))
    src = torch.from_numpy(np.ascontiguousarray(input_feature_source, dtype=np.float32)).to(dev)
    tgt = torch.from_numpy(np.ascontiguousarray(input_feature, dtype=np.float32)).to(dev)
    n, d = tgt.shape
    plan = get_plan(n, src.shape[0], d, dev.index)
    print('computing source distance...')
    print('computing original distance...')
    print('starting re_ranking...')
    final = torch.empty((n, n), dtype=torch.float64, device=dev)
    _lib.check(_lib.load().ssg_rerank_lh(plan._h, src.data_ptr(), src.shape[0], tgt.data_ptr(), n, d, int(k1), int(k2),
                                         float(lambda_value), int(dist_mode), final.data_ptr(), _lib.stream_ptr(dev)))
    euclid = sqdist(tgt, tgt, _lib.DIST_EXACT)
    return euclid.cpu().numpy(), final.cpu().numpy()


def re_ranking_init_blocks(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """reid/rerank_initial.py:40 re_ranking_init on similarity blocks (numpy or CUDA tensors).
    Returns a numpy float32 [q,g] array for numpy inputs, a CUDA tensor for CUDA inputs."""
    import torch
    dev = _lib.require_cuda()
    host = not (hasattr(q_g_dist, "is_cuda") and q_g_dist.is_cuda)

    def dv(a):
        if hasattr(a, "is_cuda"):
            return a.to(dev, dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    qg, qq, gg = dv(q_g_dist), dv(q_q_dist), dv(g_g_dist)
    q, g = qg.shape
    plan = get_plan(q + g, 1, 64, dev.index)
    out = torch.empty((q, g), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().ssg_rerank_init(plan._h, qg.data_ptr(), qq.data_ptr(), gg.data_ptr(), q, g, int(k1),
                                           int(k2), float(lambda_value), out.data_ptr(), _lib.stream_ptr(dev)))
    return out.cpu().numpy() if host else out


def sqdist(x, y, mode=_lib.DIST_EXACT):
    """Squared Euclidean distance matrix of two float32 CUDA tensors (ssg_sqdist)."""
    import torch
    x, y = x.contiguous(), y.contiguous()
    out = torch.empty((x.shape[0], y.shape[0]), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ssg_sqdist(x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0], x.shape[1], int(mode),
                                      out.data_ptr(), y.shape[0], _lib.stream_ptr(x.device)))
    return out


def dot(x, y):
    """Dot-product block x @ y.T of two float32 CUDA tensors on the library's own kernel (ssg_dot: products exact,
    float64 sequential sum) -- the np.dot blocks of reid/rerank.py:174-176 and reid/eug.py:223-225."""
    import torch
    x, y = x.contiguous(), y.contiguous()
    out = torch.empty((x.shape[0], y.shape[0]), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ssg_dot(x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0], x.shape[1], out.data_ptr(),
                                   y.shape[0], _lib.stream_ptr(x.device)))
    return out
