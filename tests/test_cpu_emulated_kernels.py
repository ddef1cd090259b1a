"""CPU: the library's plain-CUDA kernels EXECUTED on the host (tests/cpu_cuda: the .cu sources of the exact-mode
re-ranking, eps and DBSCAN -- dense, row-sharded and sparse -- compiled by g++ against a stand-in cuda_runtime.h whose
execution model is cooperative fibers), checked against the oracle, the reference's golden vectors and sklearn.

This is not the product (which needs an sm_100 GPU and has no CPU fallback): it is a second line of evidence for the
kernels' LOGIC -- index arithmetic, barrier and warp-collective protocols (a missed participant is reported as a
deadlock), tie handling, the certified sparse path -- that runs where no GPU is.  The tensor-core kernels (tcgen05 /
TMA) run in tests/test_cpu_emulated_tensor_kernels.py.  Data races are not detected as such, but their usual symptom is:
the tests at the end of this file run the kernels under reverse and pseudo-random thread / block schedules and demand
the same bytes (an injected missing barrier passes the default schedule and fails the others)."""
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


@pytest.mark.parametrize("name,args", [("shard_check_emu", ["300", "3"]), ("shard_check_emu", ["97", "5"]),
                                       ("shard_check_emu", ["64", "8"]), ("sparse_check_emu", ["400", "32"])])
def test_c_harnesses_pass_under_emulation(emu, name, args):
    """tests/c/shard_check.c and sparse_check.c (the programs the GPU box runs) against the emulated library: rows of
    final_dist, sharded eps / DBSCAN, CSR form of final_dist, certified sparse eps, sparse DBSCAN -- all equal to the
    dense single-device results."""
    r = subprocess.run([emu.bins[name]] + args, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout + r.stderr


def test_pair_exact_vec8_variant_is_byte_identical_under_emulation(emu):
    outs = []
    for v in ("0", "1"):
        r = subprocess.run([emu.bins["sparse_check_emu"], "200", "32"], capture_output=True, text=True, timeout=900,
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


def test_python_prototypes_of_the_sharded_and_sparse_entry_points(emu):
    """ssg_b200._lib.PROTOTYPES (the ctypes signatures the GPU wrappers use) applied to the emulated library, with the
    argument order of ssg_b200.cluster / rerank / dist: a swapped or mistyped argument would show as garbage here
    rather than on the GPU box.  Two ranks simulated one after another, collectives done by hand."""
    import build_emu
    from ssg_b200 import _lib as L
    lib = ctypes.CDLL(os.path.join(build_emu.OUT, "libssg_emu.so"))
    names = ["ssg_rerank_plan_create", "ssg_rerank_plan_destroy", "ssg_rerank_run", "ssg_rerank_distance_rows",
             "ssg_rerank_finish_rows", "ssg_rerank_finish_sparse", "ssg_rerank_sparse_view", "ssg_cluster_plan_create",
             "ssg_cluster_plan_destroy", "ssg_cluster_buffers", "ssg_eps_estimate", "ssg_dbscan", "ssg_eps_shard_begin",
             "ssg_eps_shard_hist", "ssg_eps_shard_pick", "ssg_eps_shard_gather", "ssg_eps_shard_finish",
             "ssg_dbscan_shard_count", "ssg_dbscan_shard_fill", "ssg_dbscan_shard_label", "ssg_eps_sparse",
             "ssg_dbscan_sparse", "ssg_last_error"]
    for nm in names:
        fn = getattr(lib, nm)
        fn.restype, fn.argtypes = L.PROTOTYPES[nm]

    def ok(rc):
        assert rc == 0, lib.ssg_last_error().decode()
    ptr = lambda a: a.ctypes.data                                                       # noqa: E731
    n, ns, d, lam, rho, world = 90, 50, 16, 0.1, 0.03, 2
    tgt, _ = O.synth_features(n, d, 1, per_cluster=10, noise=0.3)
    src, _ = O.synth_features(ns, d, 2, per_cluster=10, noise=0.4)
    rp, cp = ctypes.c_void_p(), ctypes.c_void_p()
    ok(lib.ssg_rerank_plan_create(ctypes.byref(rp), 0, n, ns, d))
    ok(lib.ssg_cluster_plan_create(ctypes.byref(cp), 0, n, 0))
    final = np.empty((n, n), np.float64)
    ok(lib.ssg_rerank_run(rp, ptr(src), ns, ptr(tgt), n, d, 20, 6, lam, L.DIST_EXACT, ptr(final), None, None))
    eps0, top0, ncl0 = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_int()
    ok(lib.ssg_eps_estimate(cp, ptr(final), L.F64, n, rho, ctypes.byref(eps0), ctypes.byref(top0), None))
    lab0 = np.empty(n, np.int64)
    ok(lib.ssg_dbscan(cp, ptr(final), L.F64, n, eps0.value, 4, ptr(lab0), ctypes.byref(ncl0), None))
    assert ncl0.value >= 2

    # ---- rows of final_dist (dist.CudaBackend.finish_rows)
    from ssg_b200.dist import shard_bounds
    blocks = []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        blk = np.empty((hi - lo, n), np.float64)
        ok(lib.ssg_rerank_distance_rows(rp, ptr(src), ns, ptr(tgt), n, d, 20, L.DIST_EXACT, lo, hi - lo, None, None))
        blocks.append(blk)
    for r in range(world):            # tables complete (one plan holds all rows here): finish every block
        lo, hi = shard_bounds(n, world, r)
        ok(lib.ssg_rerank_finish_rows(rp, ptr(tgt), n, d, 20, 6, lam, lo, hi - lo, ptr(blocks[r]), None))
    assert np.array_equal(np.concatenate(blocks), final)

    # ---- sharded eps (dist.sharded_eps / cluster.ClusterPlan.eps_shard_*), two plans = two ranks
    def buffers(plan):
        p = [ctypes.c_void_p() for _ in range(6)]
        ok(lib.ssg_cluster_buffers(plan, *[ctypes.byref(x) for x in p]))
        mk = lambda v, ct, cnt: np.ctypeslib.as_array(ctypes.cast(v.value, ctypes.POINTER(ct)), shape=(cnt,))   # noqa: E731
        return dict(hist=mk(p[0], ctypes.c_int64, L.EPS_BINS), state=mk(p[1], ctypes.c_int64, 8),
                    partial=mk(p[2], ctypes.c_double, n), list=mk(p[3], ctypes.c_double, L.EPS_LIST_CAP),
                    cnt=mk(p[4], ctypes.c_int32, n), nbr_ptr=p[5])
    plans = []
    for r in range(world):
        h = ctypes.c_void_p()
        ok(lib.ssg_cluster_plan_create(ctypes.byref(h), 0, n, 0))
        plans.append((h, buffers(h)))
        ok(lib.ssg_eps_shard_begin(h, None))
    for npass in (0, 1):
        for r, (h, b) in enumerate(plans):
            ok(lib.ssg_eps_shard_hist(h, ptr(blocks[r]), L.F64, n, world, r, npass, None))
        tot = sum(b["hist"].copy() for _, b in plans)
        for h, b in plans:
            b["hist"][:] = tot
            ok(lib.ssg_eps_shard_pick(h, npass, rho, None))
    counts, lists = [], []
    for r, (h, b) in enumerate(plans):
        c = ctypes.c_longlong()
        ok(lib.ssg_eps_shard_gather(h, ptr(blocks[r]), L.F64, n, world, r, 0, ctypes.byref(c), None))
        counts.append(c.value)
        lists.append(b["list"][:c.value].copy())
    assert min(counts) >= 0
    part = np.zeros(n)
    for r, (h, b) in enumerate(plans):
        lo, hi = shard_bounds(n, world, r)
        part[lo:hi] = b["partial"][lo:hi]
    merged = np.concatenate(lists)
    got = []
    for h, b in plans:
        b["partial"][:] = part
        b["list"][:len(merged)] = merged
        b["state"][5] = len(merged)
        e, t = ctypes.c_double(), ctypes.c_longlong()
        ok(lib.ssg_eps_shard_finish(h, n, 0, ctypes.byref(e), ctypes.byref(t), None))
        got.append((e.value, t.value))
    assert got[0] == got[1] and got[0][1] == top0.value
    np.testing.assert_allclose(got[0][0], eps0.value, rtol=1e-13)

    # ---- sharded DBSCAN (dist.sharded_dbscan)
    for r, (h, b) in enumerate(plans):
        lo, hi = shard_bounds(n, world, r)
        ok(lib.ssg_dbscan_shard_count(h, ptr(blocks[r]), L.F64, n, lo, hi - lo, eps0.value, None))
    cnt = np.zeros(n, np.int32)
    for r, (h, b) in enumerate(plans):
        lo, hi = shard_bounds(n, world, r)
        cnt[lo:hi] = b["cnt"][lo:hi]
    nbr_sum = None
    for r, (h, b) in enumerate(plans):
        lo, hi = shard_bounds(n, world, r)
        b["cnt"][:] = cnt
        total = ctypes.c_longlong()
        ok(lib.ssg_dbscan_shard_fill(h, ptr(blocks[r]), L.F64, n, lo, hi - lo, eps0.value, ctypes.byref(total), None))
        nb = np.ctypeslib.as_array(ctypes.cast(b["nbr_ptr"].value, ctypes.POINTER(ctypes.c_int32)), shape=(total.value,))
        nbr_sum = nb.copy() if nbr_sum is None else nbr_sum + nb
    for h, b in plans:
        nb = np.ctypeslib.as_array(ctypes.cast(b["nbr_ptr"].value, ctypes.POINTER(ctypes.c_int32)), shape=(len(nbr_sum),))
        nb[:] = nbr_sum
        lab, ncl = np.empty(n, np.int64), ctypes.c_int()
        ok(lib.ssg_dbscan_shard_label(h, n, 4, ptr(lab), ctypes.byref(ncl), None))
        assert np.array_equal(lab, lab0) and ncl.value == ncl0.value

    # ---- sparse form (rerank.RerankPlan.finish_sparse, cluster.ClusterPlan.eps_sparse / dbscan_sparse)
    nnz = ctypes.c_longlong()
    ok(lib.ssg_rerank_distance_rows(rp, ptr(src), ns, ptr(tgt), n, d, 20, L.DIST_EXACT, 0, n, None, None))
    ok(lib.ssg_rerank_finish_sparse(rp, ptr(tgt), n, d, 20, 6, lam, ctypes.byref(nnz), None))
    v = [ctypes.c_void_p() for _ in range(3)]
    cnt2, thr = ctypes.c_longlong(), ctypes.c_double()
    ok(lib.ssg_rerank_sparse_view(rp, ctypes.byref(v[0]), ctypes.byref(v[1]), ctypes.byref(v[2]), ctypes.byref(cnt2),
                                  ctypes.byref(thr)))
    assert cnt2.value == nnz.value and thr.value == float(np.float32(1 - lam))
    e, t, certified = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_int()
    ok(lib.ssg_eps_sparse(cp, n, v[0], v[1], v[2], thr.value, rho, ctypes.byref(e), ctypes.byref(t), ctypes.byref(certified), None))
    assert certified.value == 1 and t.value == top0.value
    np.testing.assert_allclose(e.value, eps0.value, rtol=1e-13)
    lab, ncl = np.empty(n, np.int64), ctypes.c_int()
    ok(lib.ssg_dbscan_sparse(cp, n, v[0], v[1], v[2], eps0.value, 4, ptr(lab), ctypes.byref(ncl), None))
    assert np.array_equal(lab, lab0) and ncl.value == ncl0.value
    for h, _ in plans:
        lib.ssg_cluster_plan_destroy(h)
    lib.ssg_cluster_plan_destroy(cp)
    lib.ssg_rerank_plan_destroy(rp)


def test_sparse_jaccard_beyond_256_bitmap_words(emu):
    """n = 9000 -> 282 bitmap words: the word-rank scan of jaccard_sparse_kernel takes two passes of its block-wide loop,
    as at the production size.  Last run: profiles/r01_emulated_checks.log."""
    r = subprocess.run([emu.bins["jaccard_big"], "9000"], capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0 and "JACCARD_BIG PASSED" in r.stdout, r.stdout + r.stderr


def test_k2_one_lambda_and_no_rerank_paths_under_emulation(emu):
    """rerank.py:65-66 (no_rerank returns after the distance stages), :94 (k2 == 1 skips the query expansion) and other
    lambda values, exact mode, against the oracle."""
    tgt, _ = O.synth_features(70, 24, 5, per_cluster=7)
    src, _ = O.synth_features(40, 24, 6, per_cluster=7)
    for k1, k2, lam in ((20, 1, 0.3), (20, 6, 0.0), (10, 3, 0.5)):
        _, f_ref = O.re_ranking(src, tgt, k1=k1, k2=k2, lambda_value=lam, mode="f32")
        _, f = emu.re_ranking(src, tgt, k1=k1, k2=k2, lam=lam)
        np.testing.assert_allclose(f, f_ref, rtol=0, atol=1e-4)


def test_dbscan_kernels_property_based_under_emulation(emu):
    """Random symmetric matrices with ties, several eps / min_samples: labels identical to sklearn (the order-free
    union-find formulation against the index-ordered DFS)."""
    from sklearn.cluster import DBSCAN
    rng = np.random.RandomState(7)
    for trial in range(6):
        n = int(rng.randint(5, 80))
        a = rng.rand(n, n)
        d = np.round((a + a.T) / 2, int(rng.randint(1, 3)))
        np.fill_diagonal(d, np.round(rng.rand(n) * 0.3, 2))          # the diagonal counts like any other entry
        for eps in (0.1, 0.3, 0.55):
            ms = int(rng.randint(1, 6))
            want = DBSCAN(eps=eps, min_samples=ms, metric="precomputed").fit_predict(d)
            _, got = emu.eps_and_labels(d, eps=eps, min_samples=ms)
            assert np.array_equal(got, want), (trial, n, eps, ms)


@pytest.mark.parametrize("n,ns,k,quant", [(60, 40, 20, None), (25, 10, 20, None), (20, 30, 20, None), (48, 48, 5, None),
                                          (70, 30, 20, 1), (40, 20, 3, 0)])
def test_plain_knn_set_reranker_against_oracle(emu, n, ns, k, quant):
    """ssg_rerank_plain (reid/rerank_plain.py:125-178) under emulation, exact mode, against the
    restatement that is pinned bit for bit to the reference (oracle/rerank_plain_oracle.py).  quant rounds the features so
    that many distances tie at the k-th neighbour: those rows go through the exact fallback scan and their sets grow
    beyond k, as `tem_vec <= kThreshold` does in the reference."""
    import build_emu
    from ssg_b200 import _lib as L
    from oracle import rerank_plain_oracle as P
    lib = ctypes.CDLL(os.path.join(build_emu.OUT, "libssg_emu.so"))
    for nm in ("ssg_rerank_plan_create", "ssg_rerank_plan_destroy", "ssg_rerank_plain", "ssg_last_error"):
        getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]
    rng = np.random.RandomState(n + k)
    tgt, src = rng.randn(n, 12).astype(np.float32), rng.randn(ns, 12).astype(np.float32)
    if quant is not None:
        tgt, src = np.round(tgt, quant), np.round(src, quant)
    want = P.re_ranking_plain(src, tgt, k=k, lambda_value=0.1, mode="f32")
    plan = ctypes.c_void_p()
    assert lib.ssg_rerank_plan_create(ctypes.byref(plan), 0, n, ns, 12) == 0
    got = np.empty((n, n), np.float64)
    rc = lib.ssg_rerank_plain(plan, src.ctypes.data, ns, tgt.ctypes.data, n, 12, k, 0.1, L.DIST_EXACT, got.ctypes.data, None)
    assert rc == 0, lib.ssg_last_error().decode()
    lib.ssg_rerank_plan_destroy(plan)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)          # the Jaccard part is exact; exp() of the source term
    assert np.array_equal(got, got.T)


@pytest.mark.parametrize("n,ns,d,k1,k2,lam,kind", [(40, 33, 3, 1, 8, 0.1, "quant"), (4, 33, 3, 2, 8, 0.7, "gauss"),
                                                   (7, 2, 8, 2, 6, 0.7, "dups"), (65, 9, 17, 1, 6, 0.1, "dups"),
                                                   (32, 9, 8, 20, 6, 0.7, "dups"), (33, 9, 3, 1, 8, 1.0, "quant"),
                                                   (31, 1, 8, 20, 1, 0.0, "gauss")])
def test_re_ranking_kernels_edge_parameters(emu, n, ns, d, k1, k2, lam, kind):
    """Cases a fuzz run over the emulated library turned up or passed through: k2 > k1 + 1 (the reference slices its
    full argsort; the rank table must then be wider than k1 + 1 columns -- it used to read stale columns), duplicate
    and quantised targets, a single source, lambda 0 and 1."""
    rng = np.random.RandomState(n * 7 + ns)
    tgt, src = rng.randn(n, d).astype(np.float32), rng.randn(ns, d).astype(np.float32)
    if kind == "quant":
        tgt, src = np.round(tgt), np.round(src)
    if kind == "dups":
        tgt[n // 2:] = tgt[: n - n // 2]
    _, want = O.re_ranking(src, tgt, k1=k1, k2=k2, lambda_value=lam, mode="f32")
    _, got = emu.re_ranking(src, tgt, k1=k1, k2=k2, lam=lam)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4, equal_nan=True)


def test_k1_beyond_the_row_capacity_is_refused(emu):
    rng = np.random.RandomState(0)
    tgt, src = rng.randn(40, 8).astype(np.float32), rng.randn(9, 8).astype(np.float32)
    with pytest.raises(AssertionError, match="capacity|may expand a row"):      # k-reciprocal row or expanded row too long
        emu.re_ranking(src, tgt, k1=21, k2=6, lam=0.1)
    with pytest.raises(AssertionError, match="capacity"):
        emu.re_ranking(src, tgt, k1=21, k2=1, lam=0.1)


@pytest.mark.parametrize("q,g,k1,k2,lam,quant", [(1, 22, 2, 6, 0.0, False), (30, 40, 5, 6, 1.0, True), (3, 22, 20, 1, 0.3, False),
                                                  (12, 5, 2, 3, 0.3, True)])
def test_re_ranking_init_kernels_against_oracle(emu, q, g, k1, k2, lam, quant):
    """reid/rerank_initial.py:40-99 (cosine k-reciprocal re-ranking on similarity blocks) under emulation, including
    k2 > k1 + 1 (the table must hold max(k1 + 1, k2) sorted columns) and quantised similarities."""
    import build_emu
    from ssg_b200 import _lib as L
    lib = ctypes.CDLL(os.path.join(build_emu.OUT, "libssg_emu.so"))
    for nm in ("ssg_rerank_plan_create", "ssg_rerank_plan_destroy", "ssg_rerank_init", "ssg_last_error"):
        getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]
    rng = np.random.RandomState(q * 100 + g)
    f = rng.randn(q + g, 16).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    if quant:
        f = np.round(f, 1)
    Q, G = f[:q], f[q:]
    qg, qq, gg = [np.ascontiguousarray(x, np.float32) for x in (Q @ G.T, Q @ Q.T, G @ G.T)]
    want = O.re_ranking_init(qg, qq, gg, k1=k1, k2=k2, lambda_value=lam)
    plan = ctypes.c_void_p()
    assert lib.ssg_rerank_plan_create(ctypes.byref(plan), 0, q + g, 1, 64) == 0
    out = np.empty((q, g), np.float32)
    rc = lib.ssg_rerank_init(plan, qg.ctypes.data, qq.ctypes.data, gg.ctypes.data, q, g, k1, k2, lam, out.ctypes.data, None)
    assert rc == 0, lib.ssg_last_error().decode()
    lib.ssg_rerank_plan_destroy(plan)
    np.testing.assert_allclose(out, want, rtol=0, atol=1e-5)


def test_triplet_loss_kernels_against_reference_goldens(emu, golden_dir):
    """Row f1: ssg_triplet_forward / backward (csrc/triplet.cu) under emulation against the golden vectors made by the
    reference's TripletLoss module (loss, precision, gradient; 1e-5 relative, the tolerance of that row)."""
    import build_emu
    from ssg_b200 import _lib as L
    lib = ctypes.CDLL(os.path.join(build_emu.OUT, "libssg_emu.so"))
    for nm in ("ssg_triplet_forward", "ssg_triplet_backward", "ssg_last_error"):
        getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]
    g = np.load(os.path.join(golden_dir, "triplet_cases.npz"))
    done = 0
    for ci, row in enumerate(g["cases"]):
        x, t = np.ascontiguousarray(g["x_%d" % ci], np.float32), np.ascontiguousarray(g["t_%d" % ci], np.int64)
        n, d = x.shape
        if n * n * d > 3e7:
            continue                                    # keep the emulated tier quick
        K, margin, semi = int(row[1]), float(row[4]), int(bool(row[5]))
        dist, coef = np.empty((n, n), np.float32), np.empty((n, n), np.float32)
        lp, status, grad = np.empty(2, np.float32), np.zeros(2, np.int32), np.empty((n, d), np.float32)
        rc = lib.ssg_triplet_forward(x.ctypes.data, t.ctypes.data, n, d, K, margin, semi, dist.ctypes.data, coef.ctypes.data,
                                     lp.ctypes.data, status.ctypes.data, None)
        assert rc == 0, lib.ssg_last_error().decode()
        assert status[0] == 0
        rc = lib.ssg_triplet_backward(x.ctypes.data, n, d, coef.ctypes.data, None, grad.ctypes.data, None)
        assert rc == 0, lib.ssg_last_error().decode()
        loss, ref = float(g["loss_%d" % ci]), g["grad_%d" % ci]
        assert abs(float(lp[0]) - loss) <= 1e-5 * max(1.0, abs(loss)), ci
        assert abs(float(lp[1]) - float(g["prec_%d" % ci])) < 1e-6, ci
        assert np.abs(grad - ref).max() <= 1e-5 * np.abs(ref).max(), ci
        done += 1
    assert done >= 2


@pytest.mark.parametrize("n,ns,k1,k2,lam", [(60, 40, 20, 6, 0.2), (25, 1, 5, 8, 0.5), (40, 33, 10, 1, 0.0)])
def test_re_ranking_lh_kernels_against_oracle(emu, n, ns, k1, k2, lam):
    """Row f4, re_ranking_lh (reid/rerank_plain.py:27-123): the float64 source term on un-squared distances, under
    emulation against the restatement pinned to the reference (tests/test_oracle_vs_reference.py)."""
    import build_emu
    from ssg_b200 import _lib as L
    from oracle import rerank_plain_oracle as P
    lib = ctypes.CDLL(os.path.join(build_emu.OUT, "libssg_emu.so"))
    for nm in ("ssg_rerank_plan_create", "ssg_rerank_plan_destroy", "ssg_rerank_lh", "ssg_last_error"):
        getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]
    rng = np.random.RandomState(n + ns)
    tgt, src = rng.randn(n, 12).astype(np.float32), rng.randn(ns, 12).astype(np.float32)
    want = P.re_ranking_lh(src, tgt, k1, k2, lam, "f32")
    plan = ctypes.c_void_p()
    assert lib.ssg_rerank_plan_create(ctypes.byref(plan), 0, n, ns, 12) == 0
    got = np.empty((n, n), np.float64)
    rc = lib.ssg_rerank_lh(plan, src.ctypes.data, ns, tgt.ctypes.data, n, 12, k1, k2, lam, L.DIST_EXACT, got.ctypes.data, None)
    assert rc == 0, lib.ssg_last_error().decode()
    lib.ssg_rerank_plan_destroy(plan)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4)     # J to float32 rounding (exp in the weights), v exact
    assert np.array_equal(got, got.T)


def _tensor_vs_exact(lib, L, src, tgt):
    def run(mode):
        n, d = tgt.shape
        plan = ctypes.c_void_p()
        assert lib.ssg_rerank_plan_create(ctypes.byref(plan), 0, n, src.shape[0], d) == 0
        f = np.empty((n, n))
        rc = lib.ssg_rerank_run(plan, src.ctypes.data, src.shape[0], tgt.ctypes.data, n, d, 20, 6, 0.1, mode, f.ctypes.data, None, None)
        assert rc == 0, lib.ssg_last_error().decode()
        rank, fl = np.empty((n, 32), np.int32), np.empty(1, np.int32)
        assert lib.ssg_rerank_get_stage(plan, L.STAGE_RANK, rank.ctypes.data, rank.nbytes) == 0
        assert lib.ssg_rerank_get_stage(plan, L.STAGE_FLAGGED, fl.ctypes.data, 4) == 0
        lib.ssg_rerank_plan_destroy(plan)
        return f, rank[:, :21], int(fl[0])
    fe, re_, _ = run(L.DIST_EXACT)
    ft, rt, flagged = run(L.DIST_TENSOR)
    return np.array_equal(re_, rt) and np.array_equal(fe, ft), flagged


def test_tensor_distance_mode_is_exact_for_any_approximation_within_the_bound(emu):
    """DESIGN.md §4 under emulation: the operand split, candidate selection, exact re-scoring, certification and
    fallback are the library's kernels; the approximate d2 matrix comes from a float stand-in that can be perturbed by
    up to f * E per entry (E = the certified bound).  For f < 1 the outputs must equal the exact mode bit for bit --
    tie-heavy integer features (rows that cannot be certified take the exact fallback) and clustered data alike; f = 30
    violates the error model and must be able to break it (the test has teeth)."""
    import build_emu
    from ssg_b200 import _lib as L
    lib = ctypes.CDLL(os.path.join(build_emu.OUT, "libssg_emu.so"))
    for nm in ("ssg_rerank_plan_create", "ssg_rerank_plan_destroy", "ssg_rerank_run", "ssg_rerank_get_stage", "ssg_last_error"):
        getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]
    rng = np.random.RandomState(0)
    ti = rng.randint(0, 3, (150, 8)).astype(np.float32)
    si = rng.randint(0, 3, (40, 8)).astype(np.float32)
    tc, _ = O.synth_features(130, 32, 3, per_cluster=12)
    sc, _ = O.synth_features(70, 32, 4, noise=0.6)
    saved = os.environ.get("SSG_EMU_GEMM_NOISE")
    try:
        for f in ("0", "0.5", "0.95"):
            os.environ["SSG_EMU_GEMM_NOISE"] = f
            same, flagged = _tensor_vs_exact(lib, L, si, ti)
            assert same and flagged > 0, (f, flagged)          # ties: some rows must take the exact fallback
            same, flagged = _tensor_vs_exact(lib, L, sc, tc)
            assert same, f
        os.environ["SSG_EMU_GEMM_NOISE"] = "30"
        same, _ = _tensor_vs_exact(lib, L, si, ti)
        assert not same                                         # an error 30x the bound is NOT covered, and shows
    finally:
        if saved is None:
            os.environ.pop("SSG_EMU_GEMM_NOISE", None)
        else:
            os.environ["SSG_EMU_GEMM_NOISE"] = saved


def test_symmetric_distance_matrix_leaves_the_outputs_unchanged_under_emulation(emu):
    """SSG_DIST_SYM=1 (read once per process -> subprocess): the mirrored approximate matrix (stand-in GEMM with the same
    upper-triangle-and-mirror semantics) feeds the same pipeline; outputs equal to the exact mode."""
    script = (
        "import sys, os, ctypes, numpy as np\n"
        "sys.path[:0] = %r\n"
        "import build_emu\n"
        "from ssg_b200 import _lib as L\n"
        "sys.path.insert(0, %r)\n"
        "import test_cpu_emulated_kernels as T\n"
        "from oracle import ssg_oracle as O\n"
        "lib = ctypes.CDLL(os.path.join(build_emu.OUT, 'libssg_emu.so'))\n"
        "for nm in ('ssg_rerank_plan_create', 'ssg_rerank_plan_destroy', 'ssg_rerank_run', 'ssg_rerank_get_stage', 'ssg_last_error'):\n"
        "    getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]\n"
        "t, _ = O.synth_features(120, 32, 3, per_cluster=12); s, _ = O.synth_features(60, 32, 4, noise=0.6)\n"
        "same, flagged = T._tensor_vs_exact(lib, L, s, t)\n"
        "print('SAME' if same else 'DIFFERENT', flagged)\n"
        % ([ROOT, os.path.join(ROOT, "self-similarity-grouping_b200"), os.path.join(ROOT, "tests", "cpu_cuda")],
           os.path.join(ROOT, "tests")))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, SSG_DIST_SYM="1", SSG_EMU_GEMM_NOISE="0.9"))
    assert r.returncode == 0 and "SAME" in r.stdout, r.stdout + r.stderr


# ---------------------------------------------------------------------------------------------------
# Thread / block schedules (tests/cpu_cuda/emu.cpp: SSG_EMU_SCHED, ssg_emu_set_sched).  The default schedule runs the
# threads of a block in index order between barriers and the blocks in grid order -- one of many orders the hardware
# may produce.  A kernel without races between its barriers, and whose outputs do not depend on the order in which
# blocks run or atomics land, computes the same bytes under every schedule.
# ---------------------------------------------------------------------------------------------------
SCHEDULES = {"forward": (0, 0), "reverse": (1, 0), "random-1": (2, 1), "random-2": (2, 2)}


def test_outputs_are_byte_identical_under_every_thread_and_block_schedule(emu):
    """Re-ranking in the exact and in the tensor distance mode, eps and DBSCAN on its result, the kNN-set re-ranker and
    re_ranking_lh: forward, reverse and two pseudo-random schedules (fresh permutation per sweep and per grid, the
    thread that completes a barrier or a warp collective gets no head start) give the same bytes."""
    import build_emu
    from ssg_b200 import _lib as L
    lib = ctypes.CDLL(build_emu.build()[0])
    for name, (res, args) in L.PROTOTYPES.items():
        if hasattr(lib, name):
            getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    n, ns, d = 150, 70, 32
    tgt, _ = O.synth_features(n, d, 3, per_cluster=12)
    src, _ = O.synth_features(ns, d, 4, noise=0.6)
    tgt[5] = tgt[6]                                         # duplicate rows: exact ties in the rank tables
    runs = {}
    try:
        for name, (mode, seed) in SCHEDULES.items():
            lib.ssg_emu_set_sched(mode, seed)
            out = []
            plan = ctypes.c_void_p()
            assert lib.ssg_rerank_plan_create(ctypes.byref(plan), 0, n, ns, d) == 0
            for dist_mode in (L.DIST_EXACT, L.DIST_TENSOR):
                f = np.empty((n, n))
                rc = lib.ssg_rerank_run(plan, src.ctypes.data, ns, tgt.ctypes.data, n, d, 20, 6, 0.1, dist_mode,
                                        f.ctypes.data, None, None)
                assert rc == 0, lib.ssg_last_error().decode()
                out.append(f)
            for fn, args in ((lib.ssg_rerank_plain, (20, 0.1)), (lib.ssg_rerank_lh, (20, 6, 0.1))):
                g = np.empty((n, n))
                rc = fn(plan, src.ctypes.data, ns, tgt.ctypes.data, n, d, *args, L.DIST_EXACT, g.ctypes.data, None)
                assert rc == 0, lib.ssg_last_error().decode()
                out.append(g)
            lib.ssg_rerank_plan_destroy(plan)
            eps, labels = emu.eps_and_labels(out[0], rho=0.06)
            out += [np.array([eps]), labels]
            runs[name] = out
    finally:
        lib.ssg_emu_set_sched(0, 0)
    assert runs["forward"][5].max() >= 1                    # there are clusters to get wrong
    for name, out in runs.items():
        for a, b in zip(out, runs["forward"]):
            assert a.tobytes() == b.tobytes(), name


def test_this_module_passes_under_other_schedules():
    """Everything above (oracle, golden-vector and sklearn comparisons; the C harnesses inherit the environment) once
    more under a reverse and under a pseudo-random schedule.  The slowest cases are left out."""
    if os.environ.get("SSG_EMU_SCHED"):
        pytest.skip("already inside a scheduled run")
    import build_emu
    from concurrent.futures import ThreadPoolExecutor
    build_emu.build()
    keep = "not (beyond_256 or property_based or schedule or within_the_bound or python_prototypes)"
    cmd = [sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-p", "no:cacheprovider", "-k", keep]
    with ThreadPoolExecutor(max_workers=2) as pool:
        runs = {s: pool.submit(subprocess.run, cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT,
                               env=dict(os.environ, SSG_EMU_SCHED=s)) for s in ("reverse", "random:7")}
        for s, fut in runs.items():
            r = fut.result()
            assert r.returncode == 0 and " passed" in r.stdout, (s, r.stdout[-3000:] + r.stderr[-2000:])


SCHED_FAULT_SCRIPT = r"""
import sys, ctypes, numpy as np
sys.path[:0] = %r
import build_emu
from ssg_b200 import _lib as L
from oracle import ssg_oracle as O
lib = ctypes.CDLL(build_emu.build(fault=sys.argv[1] if sys.argv[1] != "none" else None)[0])
for nm in ("ssg_cluster_plan_create", "ssg_cluster_plan_destroy", "ssg_eps_estimate_host"):
    getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]
rng = np.random.RandomState(0)
n = 200
a = rng.rand(n, n)
dmat = (a + a.T) / 2
np.fill_diagonal(dmat, 0)
want = O.eps_estimate(dmat, 0.05)
for name, mode in (("forward", 0), ("reverse", 1), ("random", 2)):
    lib.ssg_emu_set_sched(mode, 1)
    plan = ctypes.c_void_p()
    assert lib.ssg_cluster_plan_create(ctypes.byref(plan), 0, n, 0) == 0
    e, top = ctypes.c_double(), ctypes.c_longlong()
    rc = lib.ssg_eps_estimate_host(plan, dmat.ctypes.data, 1, n, 0.05, ctypes.byref(e), ctypes.byref(top))
    lib.ssg_cluster_plan_destroy(plan)
    print(name, "RIGHT" if rc == 0 and abs(e.value - want) <= 1e-12 * want else "WRONG")
"""


SCHED_FAULT_CASES = [
    ("none", {"forward": "RIGHT", "reverse": "RIGHT", "random": "RIGHT"}),
    # eps_pick_kernel without the barrier between thread 0 publishing the selection state and everybody reading it:
    # right whenever thread 0 happens to run first, as it does in the forward schedule
    ("eps_pick_no_barrier", {"forward": "RIGHT", "reverse": "WRONG", "random": "WRONG"}),
]


def test_schedules_catch_an_injected_race():
    import build_emu
    from concurrent.futures import ThreadPoolExecutor
    for fault, _ in SCHED_FAULT_CASES:
        build_emu.build(fault=None if fault == "none" else fault)
    paths = [ROOT, os.path.join(ROOT, "self-similarity-grouping_b200"), os.path.join(ROOT, "tests", "cpu_cuda")]
    env = {k: v for k, v in os.environ.items() if k != "SSG_EMU_SCHED"}
    with ThreadPoolExecutor(max_workers=2) as pool:
        runs = [pool.submit(subprocess.run, [sys.executable, "-c", SCHED_FAULT_SCRIPT % (paths,), fault],
                            capture_output=True, text=True, timeout=900, env=env) for fault, _ in SCHED_FAULT_CASES]
        for (fault, expect), fut in zip(SCHED_FAULT_CASES, runs):
            r = fut.result()
            assert r.returncode == 0, (fault, r.stdout + r.stderr)
            got = dict(l.split() for l in r.stdout.splitlines() if l.split() and l.split()[0] in expect)
            assert got == expect, (fault, r.stdout + r.stderr)
