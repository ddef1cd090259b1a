"""GPU parity: the CUDA re-ranking path (through the C ABI) against the oracle, stage by stage.

Tolerances: integer / index results are bit-exact; squared distances are bit-exact (the exact mode
restates cdist's float64 summation order); float32 values downstream of exp() within 2e-6; final_dist
within 1e-4 (BASELINE.json north_star) — in practice ~1e-7.
"""
import os

import numpy as np
import pytest

from oracle import ssg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ssg():
    import ssg_b200
    return ssg_b200


def _sparse_rows(cnt, idx, val, n):
    M = np.zeros((n, n), np.float32)
    for i in range(n):
        M[i, idx[i, :cnt[i]]] = val[i, :cnt[i]]
    return M


@pytest.mark.parametrize("nx,ny,d", [(1, 1, 4), (37, 91, 7), (130, 64, 2048), (200, 333, 515)])
def test_sqdist_exact_is_bit_identical_to_cdist(ssg, nx, ny, d):
    import torch
    from scipy.spatial.distance import cdist
    rng = np.random.RandomState(nx + ny + d)
    x = rng.randn(nx, d).astype(np.float32)
    y = rng.randn(ny, d).astype(np.float32)
    got = ssg.sqdist(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()).cpu().numpy()
    want = np.power(cdist(x, y).astype(np.float32), 2).astype(np.float32)   # rerank.py:61-62 in O-f32
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,ns,d,seed", [(64, 50, 128, 0), (257, 300, 2048, 1), (512, 400, 256, 2),
                                          (1000, 1000, 512, 3)])
def test_stages_against_oracle(ssg, n, ns, d, seed):
    import torch
    from ssg_b200 import _lib
    tgt, _ = O.synth_features(n, d, seed)
    src, _ = O.synth_features(ns, d, seed + 77, noise=0.6)
    st = {}
    e_ref, f_ref = O.re_ranking(src, tgt, lambda_value=0.1, mode="f32", stages=st)
    plan = ssg.RerankPlan(n, ns, d)
    e, f = plan.run(torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda(), lambda_value=0.1,
                    want_euclid=True)
    torch.cuda.synchronize()
    e, f = e.cpu().numpy(), f.cpu().numpy()
    # (ii) squared distances: bit exact
    assert np.array_equal(e, e_ref)
    # (i) source vector
    np.testing.assert_allclose(plan.stage(_lib.STAGE_VEC, n), st["vec"], rtol=0, atol=2e-6)
    # (iii) row normaliser, (iv) leading rank columns: exact
    assert np.array_equal(plan.stage(_lib.STAGE_ROWMAX, n), e_ref.max(axis=0))
    rank = plan.stage(_lib.STAGE_RANK, n)[:, :21]
    assert np.array_equal(rank, st["rank"][:, :21])
    np.testing.assert_array_equal(plan.stage(_lib.STAGE_RANK_VAL, n)[:, :21],
                                  np.take_along_axis(st["odn"], st["rank"][:, :21].astype(np.int64), 1))
    # (v) k-reciprocal rows: same support, values to float32 exp accuracy
    V = _sparse_rows(plan.stage(_lib.STAGE_V_CNT, n), plan.stage(_lib.STAGE_V_IDX, n),
                     plan.stage(_lib.STAGE_V_VAL, n), n)
    assert np.array_equal(V != 0, st["V"] != 0)
    np.testing.assert_allclose(V, st["V"], rtol=0, atol=2e-6)
    # (vi) query expansion
    Vq = _sparse_rows(plan.stage(_lib.STAGE_VQ_CNT, n), plan.stage(_lib.STAGE_VQ_IDX, n),
                      plan.stage(_lib.STAGE_VQ_VAL, n), n)
    assert np.array_equal(Vq != 0, st["Vq"] != 0)
    np.testing.assert_allclose(Vq, st["Vq"], rtol=0, atol=2e-6)
    # (vii)-(viii) final distance
    assert f.dtype == np.float64 and np.array_equal(f, f.T)
    np.testing.assert_allclose(f, f_ref, rtol=0, atol=1e-4)
    assert np.abs(f - f_ref).max() < 5e-6


@pytest.mark.parametrize("case", ["rerank_n160_d256.npz", "rerank_n257_d2048.npz", "rerank_n96_d64_ties.npz"])
def test_drop_in_re_ranking_against_reference_goldens(ssg, golden_dir, case):
    g = np.load(os.path.join(golden_dir, case))
    e, f = ssg.re_ranking(g["src"], g["tgt"], lambda_value=float(g["lam"]))
    assert isinstance(f, np.ndarray) and f.dtype == np.float64 and f.flags["C_CONTIGUOUS"]
    assert np.array_equal(e, g["euclid_f32"])
    np.testing.assert_allclose(f, g["final_f32"], rtol=0, atol=1e-4)
    # pseudo-labels from our matrix with the reference's eps == the reference's labels
    for bi in range(len(g["rhos"])):
        eps = float(g["eps_%d" % bi])
        if np.abs(np.abs(g["final_f32"] - eps).min()) < 1e-5:
            continue     # an entry sits on the eps boundary: labels are not stable to 1e-7 noise
        lab = ssg.DBSCAN(eps=eps, min_samples=4, metric="precomputed", n_jobs=8).fit_predict(f)
        assert np.array_equal(lab, g["labels_%d" % bi])


def test_no_rerank_and_k2_one(ssg):
    tgt, _ = O.synth_features(120, 64, 5)
    src, _ = O.synth_features(80, 64, 6)
    e, f = ssg.re_ranking(src, tgt, no_rerank=True)
    assert f is None and np.array_equal(e, O.original_distance(tgt))
    _, f1 = ssg.re_ranking(src, tgt, k2=1, lambda_value=0.3)
    _, f1_ref = O.re_ranking(src, tgt, k2=1, lambda_value=0.3)
    np.testing.assert_allclose(f1, f1_ref, rtol=0, atol=5e-6)


def test_argument_errors(ssg):
    import torch
    plan = ssg.RerankPlan(32, 32, 16)
    x = torch.zeros(40, 16, device="cuda")
    with pytest.raises(ValueError):
        plan.run(x, x)                       # n > n_max
    with pytest.raises(ValueError):
        plan.run(x[:8], x[:8], k1=40)        # k1 out of range


def test_full_size_properties(ssg):
    """N = 16 702 (Market-1501 shape, BASELINE.json configs[1]): size-independent properties."""
    import torch
    from ssg_b200 import _lib
    n, d, lam = 16702, 2048, 0.1
    tgt, _ = O.synth_features(n, d, 0)
    src, _ = O.synth_features(n, d, 1, noise=0.6)
    plan = ssg.RerankPlan(n, n, d)
    _, f = plan.run(torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda(), lambda_value=lam)
    torch.cuda.synchronize()
    assert torch.equal(f, f.t())                                         # exactly symmetric
    vec = torch.from_numpy(plan.stage(_lib.STAGE_VEC, n)).cuda()
    assert float(vec.max()) == 1.0 and float(vec.min()) >= 0.0
    # diagonal: J_ii = 1 - S/(2-S) with S = sum(Vq_i) ~ 1  ->  final_ii ~ 2*lambda*v_i  (App. B.3)
    np.testing.assert_allclose(f.diagonal().cpu().numpy(), 2 * lam * vec.double().cpu().numpy(), atol=1e-5)
    rank = plan.stage(_lib.STAGE_RANK, n)
    assert np.array_equal(rank[:, 0], np.arange(n))                      # self ranks first
    base = 1.0 - lam
    assert float(f.max()) <= np.float32(base) + 2 * lam + 1e-12
    # rows of the expanded matrix are distributions
    cnt = plan.stage(_lib.STAGE_VQ_CNT, n)
    val = plan.stage(_lib.STAGE_VQ_VAL, n)
    sums = np.array([val[i, :cnt[i]].sum() for i in range(0, n, 97)])
    np.testing.assert_allclose(sums, 1.0, atol=1e-5)
    # spot-check 64 rows of the final matrix against the oracle stages computed on those rows only
    rows = np.arange(0, n, n // 64)[:64]
    from scipy.spatial.distance import cdist
    od = np.power(cdist(tgt[rows], tgt).astype(np.float32), 2).astype(np.float32)
    odn = od / od.max(axis=1, keepdims=True)
    assert np.array_equal(np.argsort(odn, kind="stable")[:, :21], rank[rows, :21])


def test_re_ranking_init_against_reference_golden(ssg, golden_dir):
    """reid/rerank_initial.py:40 (blocks) and reid/rerank.py:171 (features) through the drop-in modules."""
    import reid.rerank
    import reid.rerank_initial
    g = np.load(os.path.join(golden_dir, "rerank_init_q40_g90.npz"))
    qf, gf = g["qf"], g["gf"]
    out = reid.rerank_initial.re_ranking_init(qf @ gf.T, qf @ qf.T, gf @ gf.T)
    assert out.shape == (40, 90) and out.dtype == np.float32
    np.testing.assert_allclose(out, g["final"], rtol=0, atol=1e-5)
    out2 = reid.rerank.re_ranking_init(qf, gf)
    np.testing.assert_allclose(out2, g["final"], rtol=0, atol=1e-4)   # GPU GEMM vs np.dot in the similarities
    for lam, k2 in ((0.1, 6), (0.5, 1)):
        f, _ = O.synth_features(300, 128, 9, per_cluster=10)
        q_, g_ = f[:100], f[100:]
        want = O.re_ranking_init(q_ @ g_.T, q_ @ q_.T, g_ @ g_.T, k2=k2, lambda_value=lam)
        got = reid.rerank_initial.re_ranking_init(q_ @ g_.T, q_ @ q_.T, g_ @ g_.T, k2=k2, lambda_value=lam)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)
