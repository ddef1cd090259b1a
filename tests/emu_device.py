"""Test helper: point the Python host layer (ssg_b200._lib / rerank / cluster / dist) at the CPU-emulated library of
tests/cpu_cuda, so that the REAL ctypes wrappers and the multi-GPU choreography run end to end without a GPU.

What is swapped, and only inside the test process: the shared library (libssg_emu.so, exact distance mode; symbols it
lacks -- the embedding -- stay unbound), ``require_cuda`` (returns torch.device("cpu", 0)), the stream pointer (NULL),
the zero-copy views of plan-owned buffers (numpy views instead of __cuda_array_interface__ objects; torch.as_tensor
is wrapped so that they stay views) and the
``is_cuda`` assertions of the wrappers.  The product path is untouched: without this helper every call still raises
on a machine without a GPU."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(tc=False):
    """Returns an ``undo`` callable.  tc=True: the library built against the functional tcgen05 / TMA emulation
    (build_emu.build_tc: every entry point, tensor distance mode and convolutions included)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_cuda"))
    import build_emu
    lib_path = build_emu.build_tc() if tc else build_emu.build()[0]
    from ssg_b200 import _lib, cluster, dist, rerank
    saved = dict(lib=_lib._lib, require=_lib.require_cuda, stream=_lib.stream_ptr, same=_lib.same_device, c_dev=cluster._DevArray,
                 d_dev=dist._DevArray, c_dt=cluster._dtype_code, c_rd=cluster._rows_dtype, plans=dict(rerank._plans),
                 cplans=dict(cluster._plans))
    lib = ctypes.CDLL(lib_path)
    for name, (res, args) in _lib.PROTOTYPES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype, fn.argtypes = res, args
    _lib._lib = lib
    dev = torch.device("cpu", 0)
    _lib.require_cuda = lambda device=None: dev
    _lib.stream_ptr = lambda device=None: ctypes.c_void_p(None)
    _lib.same_device = lambda t, device: True
    np_types = {"<i4": ctypes.c_int32, "<f4": ctypes.c_float, "<i8": ctypes.c_int64, "<f8": ctypes.c_double}

    def host_view(ptr, shape, typestr):
        count = int(np.prod(shape))
        arr = np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(np_types[typestr])), shape=(max(count, 1),))
        return arr[:count].reshape(shape)
    cluster._DevArray = host_view
    dist._DevArray = host_view
    # torch.as_tensor(ndarray, device=cpu:0) COPIES (the device differs from plain "cpu"); the wrappers rely on views
    orig_as_tensor = torch.as_tensor

    def as_tensor(data, *a, **kw):
        if isinstance(data, np.ndarray):
            return torch.from_numpy(data)
        return orig_as_tensor(data, *a, **kw)
    torch.as_tensor = as_tensor

    def dtype_code(t):
        assert t.dim() == 2 and t.shape[0] == t.shape[1] and t.is_contiguous()
        return _lib.F64 if t.dtype == torch.float64 else _lib.F32

    def rows_dtype(rows, n):
        assert rows.dim() == 2 and rows.shape[1] == n and rows.is_contiguous()
        return _lib.F64 if rows.dtype == torch.float64 else _lib.F32
    cluster._dtype_code, cluster._rows_dtype = dtype_code, rows_dtype
    saved["pinned"] = rerank._pinned
    rerank._pinned = lambda shape, dtype: np.empty(shape, dtype=dtype)      # pinned host memory needs a driver
    rerank._plans.clear()
    cluster._plans.clear()

    def undo():
        rerank._pinned = saved["pinned"]
        torch.as_tensor = orig_as_tensor
        rerank._plans.clear()
        cluster._plans.clear()
        _lib._lib, _lib.require_cuda, _lib.stream_ptr = saved["lib"], saved["require"], saved["stream"]
        _lib.same_device = saved["same"]
        cluster._DevArray, dist._DevArray = saved["c_dev"], saved["d_dev"]
        cluster._dtype_code, cluster._rows_dtype = saved["c_dt"], saved["c_rd"]
        rerank._plans.update(saved["plans"])
        cluster._plans.update(saved["cplans"])
    return undo
