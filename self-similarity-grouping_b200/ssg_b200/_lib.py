"""ctypes binding of libssg_b200.so (the C ABI declared in include/ssg_b200.h).

The CUDA library is the product: there is no CPU fallback.  Loading works without a GPU (symbols can
be inspected); every compute call needs an sm_100 device and raises otherwise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libssg_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
DIST_EXACT, DIST_TENSOR = 0, 1
F32, F64 = 0, 1
RANK_STRIDE, V_STRIDE, VQ_STRIDE = 32, 256, 1536
EPS_BINS, EPS_LIST_CAP = 4096, 1 << 20
(STAGE_VEC, STAGE_ROWMAX, STAGE_RANK, STAGE_RANK_VAL, STAGE_V_CNT, STAGE_V_IDX, STAGE_V_VAL,
 STAGE_VQ_CNT, STAGE_VQ_IDX, STAGE_VQ_VAL, STAGE_FLAGGED) = range(11)

c_int, c_void_p, c_size_t, c_double, c_ll = (ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_double, ctypes.c_longlong)
P = ctypes.POINTER

# name -> (restype, argtypes); mirrors include/ssg_b200.h one to one
PROTOTYPES = {
    "ssg_version": (c_int, []),
    "ssg_last_error": (ctypes.c_char_p, []),
    "ssg_device_info": (c_int, [c_int, P(c_int), P(c_int)]),
    "ssg_sqdist": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ssg_dot": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ssg_rerank_plan_create": (c_int, [P(c_void_p), c_int, c_int, c_int, c_int]),
    "ssg_rerank_plan_destroy": (c_int, [c_void_p]),
    "ssg_rerank_plan_bytes": (c_size_t, [c_void_p]),
    "ssg_rerank_run": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_double,
                               c_int, c_void_p, c_void_p, c_void_p]),
    "ssg_rerank_distance_rows": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p]),
    "ssg_rerank_tables": (c_int, [c_void_p, P(c_void_p), P(c_void_p), P(c_void_p), P(c_void_p)]),
    "ssg_rerank_finish": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p]),
    "ssg_rerank_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_double,
                                c_int, c_int, c_void_p, c_void_p]),
    "ssg_rerank_init": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p,
                                c_void_p]),
    "ssg_rerank_get_stage": (c_int, [c_void_p, c_int, c_void_p, c_size_t]),
    "ssg_cluster_plan_create": (c_int, [P(c_void_p), c_int, c_int, c_ll]),
    "ssg_cluster_plan_destroy": (c_int, [c_void_p]),
    "ssg_cluster_plan_bytes": (c_size_t, [c_void_p]),
    "ssg_eps_estimate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, P(c_double), P(c_ll), c_void_p]),
    "ssg_dbscan": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_void_p, P(c_int), c_void_p]),
    "ssg_dbscan_core_mask": (c_int, [c_void_p, c_void_p, c_int]),
    "ssg_eps_estimate_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, P(c_double), P(c_ll)]),
    "ssg_dbscan_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_void_p, P(c_int)]),
    "ssg_rerank_finish_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int, c_int, c_void_p,
                                       c_void_p]),
    "ssg_cluster_buffers": (c_int, [c_void_p, P(c_void_p), P(c_void_p), P(c_void_p), P(c_void_p), P(c_void_p),
                                    P(c_void_p)]),
    "ssg_eps_shard_begin": (c_int, [c_void_p, c_void_p]),
    "ssg_eps_shard_hist": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ssg_eps_shard_pick": (c_int, [c_void_p, c_int, c_double, c_void_p]),
    "ssg_eps_shard_gather": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, P(c_ll), c_void_p]),
    "ssg_eps_shard_finish": (c_int, [c_void_p, c_int, c_int, P(c_double), P(c_ll), c_void_p]),
    "ssg_dbscan_shard_count": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p]),
    "ssg_dbscan_shard_fill": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, P(c_ll), c_void_p]),
    "ssg_dbscan_shard_label": (c_int, [c_void_p, c_int, c_int, c_void_p, P(c_int), c_void_p]),
    "ssg_rerank_plain": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p,
                                 c_void_p]),
    "ssg_rerank_lh": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int, c_void_p,
                              c_void_p]),
    "ssg_rerank_finish_sparse": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, P(c_ll), c_void_p]),
    "ssg_rerank_sparse_view": (c_int, [c_void_p, P(c_void_p), P(c_void_p), P(c_void_p), P(c_ll), P(c_double)]),
    "ssg_eps_sparse": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double, P(c_double), P(c_ll),
                               P(c_int), c_void_p]),
    "ssg_dbscan_sparse": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double, c_int, c_void_p, P(c_int),
                                  c_void_p]),
    "ssg_embed_num_layers": (c_int, []),
    "ssg_embed_layer_info": (c_int, [c_int, P(c_int), P(c_int), P(c_int), P(c_int), ctypes.c_char_p, ctypes.c_char_p,
                                     c_size_t]),
    "ssg_embed_plan_create": (c_int, [P(c_void_p), c_int, c_int, c_int, c_int]),
    "ssg_embed_plan_destroy": (c_int, [c_void_p]),
    "ssg_embed_plan_bytes": (c_size_t, [c_void_p]),
    "ssg_embed_load_layer": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     ctypes.c_float, c_void_p]),
    "ssg_embed_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "ssg_embed_forward_u8": (c_int, [c_void_p, c_void_p, P(ctypes.c_float), P(ctypes.c_float), c_int, c_int, c_int,
                                     c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "ssg_op_conv": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                            c_int, c_void_p, c_void_p, c_void_p]),
    "ssg_op_fold_bn": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_float,
                               c_int, c_void_p, c_void_p, c_void_p]),
    "ssg_op_stem": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ssg_op_pooled_tail": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    "ssg_triplet_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_float, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "ssg_triplet_backward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ssg_op_conv_pack_weight": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ssg_op_conv_dgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ssg_op_conv_wgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ssg_op_stem_im2col": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ssg_rank_metrics": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "ssg_profile_enable": (c_int, [c_int]),
    "ssg_profile_reset": (c_int, []),
    "ssg_profile_collect": (c_int, []),
    "ssg_profile_entry": (c_int, [c_int, ctypes.c_char_p, c_size_t, P(c_double), P(c_ll)]),
}

_lib = None


class SsgError(RuntimeError):
    pass


def load():
    """dlopen the library (built in-tree by __graft_entry__.build() / make) and set prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SsgError("libssg_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(expected at %s)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)     # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == OK:
        return
    msg = load().ssg_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError("ssg_b200: " + msg)
    if rc == ERR_CAPACITY:
        raise OverflowError("ssg_b200: " + msg)
    raise SsgError("ssg_b200 (code %d): %s" % (rc, msg))


def require_cuda(device=None):
    """The hot path has no CPU implementation: fail loudly when there is no sm_100 GPU."""
    import torch
    if not torch.cuda.is_available():
        raise SsgError("ssg_b200 needs a CUDA device (sm_100); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    return dev


def profile(on=None, reset=False):
    """Enable/disable the library's per-kernel CUDA-event timers; returns {name: (ms, launches)}."""
    lib = load()
    if reset:
        lib.ssg_profile_reset()
    if on is not None:
        lib.ssg_profile_enable(int(bool(on)))
        return {}
    n = lib.ssg_profile_collect()
    out = {}
    buf = ctypes.create_string_buffer(64)
    for i in range(n):
        ms, cnt = c_double(), c_ll()
        check(lib.ssg_profile_entry(i, buf, 64, ctypes.byref(ms), ctypes.byref(cnt)))
        out[buf.value.decode()] = (ms.value, cnt.value)
    return out


def stream_ptr(device=None):
    """The current torch stream OF `device` (default: the current device) as a void*.  A stream belongs to one
    device: plans pass their own device so that a call made while another device is current still launches on a
    stream of the plan's device."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def same_device(t, device):
    """True when CUDA tensor `t` lives on `device` (type AND index)."""
    return t.is_cuda and t.device.index == device.index
