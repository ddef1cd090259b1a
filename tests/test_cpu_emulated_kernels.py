"""CPU: the library's plain-CUDA kernels EXECUTED on the host (tests/cpu_cuda: the .cu sources of the exact-mode
re-ranking, eps and DBSCAN -- dense, row-sharded and sparse -- compiled by g++ against a stand-in cuda_runtime.h whose
execution model is cooperative fibers), checked against the oracle, the reference's golden vectors and sklearn.

This is not the product (which needs an sm_100 GPU and has no CPU fallback): it is a second line of evidence for the
kernels' LOGIC -- index arithmetic, barrier and warp-collective protocols (a missed participant is reported as a
deadlock), tie handling, the certified sparse path -- that runs where no GPU is.  The tensor-core kernels (tcgen05 /
TMA) are not emulated, and neither are data races."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from oracle import ssg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_cuda"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def emu():
    import build_emu
    lib_path, bins = build_emu.build()
    lib = ctypes.CDLL(lib_path)
    lib.ssg_last_error.restype = ctypes.c_char_p
    c_int, c_vp, c_d, c_ll = ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_longlong
    lib.ssg_rerank_plan_create.argtypes = [ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int]
    lib.ssg_rerank_plan_destroy.argtypes = [c_vp]
    lib.ssg_rerank_host.argtypes = [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_d, c_int, c_int, c_vp, c_vp]
    lib.ssg_cluster_plan_create.argtypes = [ctypes.POINTER(c_vp), c_int, c_int, c_ll]
    lib.ssg_cluster_plan_destroy.argtypes = [c_vp]
    lib.ssg_eps_estimate_host.argtypes = [c_vp, c_vp, c_int, c_int, c_d, ctypes.POINTER(c_d), ctypes.POINTER(c_ll)]
    lib.ssg_dbscan_host.argtypes = [c_vp, c_vp, c_int, c_int, c_d, c_int, c_vp, ctypes.POINTER(c_int)]

    bin_paths = {os.path.basename(b): b for b in bins}

    class Emu(object):
        bins = bin_paths

        @staticmethod
        def check(rc):
            assert rc == 0, lib.ssg_last_error().decode()

        def re_ranking(self, src, tgt, k1=20, k2=6, lam=0.2):
            src, tgt = np.ascontiguousarray(src, np.float32), np.ascontiguousarray(tgt, np.float32)
            n, d = tgt.shape
            plan = c_vp()
            self.check(lib.ssg_rerank_plan_create(ctypes.byref(plan), 0, n, src.shape[0], d))
            final, euclid = np.empty((n, n), np.float64), np.empty((n, n), np.float32)
            self.check(lib.ssg_rerank_host(plan, src.ctypes.data, src.shape[0], tgt.ctypes.data, n, d, k1, k2, lam, 0, 0,
                                           final.ctypes.data, euclid.ctypes.data))
            lib.ssg_rerank_plan_destroy(plan)
            return euclid, final

        def eps_and_labels(self, dist, rho=None, eps=None, min_samples=4):
            dist = np.ascontiguousarray(dist)
            dt = 1 if dist.dtype == np.float64 else 0
            n = dist.shape[0]
            plan = c_vp()
            self.check(lib.ssg_cluster_plan_create(ctypes.byref(plan), 0, n, 0))
            if eps is None:
                e, top = c_d(), c_ll()
                self.check(lib.ssg_eps_estimate_host(plan, dist.ctypes.data, dt, n, rho, ctypes.byref(e), ctypes.byref(top)))
                eps = e.value
            labels, ncl = np.empty(n, np.int64), c_int()
            if eps == eps:
                self.check(lib.ssg_dbscan_host(plan, dist.ctypes.data, dt, n, eps, min_samples, labels.ctypes.data,
                                               ctypes.byref(ncl)))
            lib.ssg_cluster_plan_destroy(plan)
            return eps, labels
    return Emu()


@pytest.mark.parametrize("name,args", [("shard_check_emu", ["150", "3"]), ("shard_check_emu", ["97", "5"]),
                                       ("sparse_check_emu", ["160", "32"])])
def test_c_harnesses_pass_under_emulation(emu, name, args):
    """tests/c/shard_check.c and sparse_check.c (the programs the GPU box runs) against the emulated library: rows of
    final_dist, sharded eps / DBSCAN, CSR form of final_dist, certified sparse eps, sparse DBSCAN -- all equal to the
    dense single-device results."""
    r = subprocess.run([emu.bins[name]] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr


def test_pair_exact_vec8_variant_is_byte_identical_under_emulation(emu):
    outs = []
    for v in ("0", "1"):
        r = subprocess.run([emu.bins["sparse_check_emu"], "120", "32"], capture_output=True, text=True, timeout=900,
                           env=dict(os.environ, SSG_PAIR_VEC8=v))
        assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr
        outs.append([l for l in r.stdout.splitlines() if l.startswith(("CSR", "eps", "dbscan"))])
    assert outs[0] == outs[1]               # same nnz, eps to 17 digits, clusters


@pytest.mark.parametrize("n,ns", [(2, 1), (3, 5), (10, 7), (21, 30), (22, 22), (64, 40)])
def test_re_ranking_kernels_against_oracle_incl_tiny_sets(emu, n, ns):
    """Exact mode, float32-mode oracle (pinned to the reference, tests/test_oracle_vs_reference.py): squared distances bit
    for bit, final_dist to 1e-4 -- including target sets smaller than the k1 + 1 = 21 rank columns and than k2 = 6
    (np.mean over the rows that exist, rerank.py:97)."""
    rng = np.random.RandomState(n * 31 + ns)
    tgt, src = rng.randn(n, 16).astype(np.float32), rng.randn(ns, 16).astype(np.float32)
    e_ref, f_ref = O.re_ranking(src, tgt, lambda_value=0.1, mode="f32")
    e, f = emu.re_ranking(src, tgt, lam=0.1)
    assert np.array_equal(e, e_ref)
    np.testing.assert_allclose(f, f_ref, rtol=0, atol=1e-4)
    assert np.array_equal(f, f.T)


def test_re_ranking_kernels_against_reference_golden_with_ties(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "rerank_n96_d64_ties.npz"))
    e, f = emu.re_ranking(g["src"], g["tgt"], lam=float(g["lam"]))
    assert np.array_equal(e, g["euclid_f32"])
    np.testing.assert_allclose(f, g["final_f32"], rtol=0, atol=1e-4)
    for bi in range(len(g["rhos"])):
        eps = float(g["eps_%d" % bi])
        if np.abs(g["final_f32"] - eps).min() < 1e-5:
            continue
        _, lab = emu.eps_and_labels(f, eps=eps)
        assert np.array_equal(lab, g["labels_%d" % bi])


@pytest.mark.parametrize("n,rho,dtype", [(40, 0.05, np.float64), (90, 0.3, np.float64), (70, 0.02, np.float32), (2, 0.5, np.float64)])
def test_eps_and_dbscan_kernels_against_numpy_and_sklearn(emu, n, rho, dtype):
    from sklearn.cluster import DBSCAN
    rng = np.random.RandomState(n)
    a = rng.rand(n, n)
    d = np.round((a + a.T) / 2, 2).astype(dtype)          # symmetric, heavy ties, non-zero diagonal
    want_eps = O.eps_estimate(d, rho)
    eps, labels = emu.eps_and_labels(d, rho=rho)
    if np.isnan(want_eps):
        assert np.isnan(eps)
        return
    # a float32 matrix: numpy takes the mean in float32, the kernel in float64
    np.testing.assert_allclose(eps, want_eps, rtol=1e-13 if dtype == np.float64 else 1e-6)
    e_cmp = float(dtype(eps)) if dtype == np.float32 else eps
    want = DBSCAN(eps=e_cmp, min_samples=4, metric="precomputed").fit_predict(d)
    _, got = emu.eps_and_labels(d, eps=e_cmp)
    assert np.array_equal(got, want)
