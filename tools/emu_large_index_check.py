#!/usr/bin/env python
"""One-off check (CPU emulation, tests/cpu_cuda): eps and DBSCAN kernels on a matrix with more than 2^31 elements.

n = 47 000 (n^2 = 2.209e9 > 2^31) float32 distances, all 1.0 except planted groups of five points at pairwise distance
0.25 -- two of them at row indices whose row offset i * n exceeds 2^31.  Expected, analytically: every upper-triangle entry
is non-zero, so M = n (n - 1) / 2 and top = round(rho * M); with rho chosen so that top <= the number of planted pairs,
eps = 0.25 exactly; DBSCAN(eps, min_samples = 4) labels the groups 0, 1, 2, ... in order of their first point and
everything else -1.  Any 32-bit row offset in the kernels reads the wrong rows.

    python tools/emu_large_index_check.py [n]      (about ten minutes and 10 GB at the default n)
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-similarity-grouping_b200"), os.path.join(ROOT, "tests", "cpu_cuda")]
import build_emu  # noqa: E402
from ssg_b200 import _lib as L  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 47000
    lib = ctypes.CDLL(build_emu.build()[0])
    for nm in ("ssg_cluster_plan_create", "ssg_cluster_plan_destroy", "ssg_eps_estimate_host", "ssg_dbscan_host", "ssg_last_error"):
        getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]
    groups = [list(range(100, 105)), [n - 600, n - 500, n - 400, n - 300, n - 200], list(range(n - 10, n - 5))]
    dist = np.ones((n, n), np.float32)
    for g in groups:
        for a in g:
            for b in g:
                dist[a, b] = 0.25
    np.fill_diagonal(dist, 0.0)
    pairs = sum(len(g) * (len(g) - 1) // 2 for g in groups)
    m = n * (n - 1) // 2
    rho = (pairs - 2.0) / m                          # top = round(rho * M) = pairs - 2 entries, all 0.25
    want = np.full(n, -1, np.int64)
    for c, g in enumerate(groups):
        want[g] = c
    print("n = %d, n^2 = %d (2^31 = %d), first row offset beyond 2^31: row %d" % (n, n * n, 2 ** 31, 2 ** 31 // n + 1), flush=True)
    plan = ctypes.c_void_p()
    assert lib.ssg_cluster_plan_create(ctypes.byref(plan), 0, n, 0) == 0, lib.ssg_last_error().decode()
    t0 = time.time()
    e, top = ctypes.c_double(), ctypes.c_longlong()
    rc = lib.ssg_eps_estimate_host(plan, dist.ctypes.data, 0, n, rho, ctypes.byref(e), ctypes.byref(top))
    assert rc == 0, lib.ssg_last_error().decode()
    print("eps = %.17g (want 0.25), top = %d (want %d)   [%.0f s]" % (e.value, top.value, pairs - 2, time.time() - t0), flush=True)
    t0 = time.time()
    labels, ncl = np.empty(n, np.int64), ctypes.c_int()
    rc = lib.ssg_dbscan_host(plan, dist.ctypes.data, 0, n, e.value, 4, labels.ctypes.data, ctypes.byref(ncl))
    assert rc == 0, lib.ssg_last_error().decode()
    lib.ssg_cluster_plan_destroy(plan)
    ok = e.value == 0.25 and top.value == pairs - 2 and ncl.value == len(groups) and np.array_equal(labels, want)
    print("dbscan: %d clusters (want %d), labels %s   [%.0f s]" % (ncl.value, len(groups), "equal" if np.array_equal(labels, want)
                                                                    else "DIFFERENT", time.time() - t0), flush=True)
    print("LARGE_INDEX_CHECK %s" % ("PASSED" if ok else "FAILED"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
