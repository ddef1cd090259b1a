"""GPU parity: eps estimate and DBSCAN (through the C ABI) against numpy / scikit-learn.

Labels are compared bit-for-bit; eps to 1e-12 relative (the mean is taken in a different but
deterministic summation order).
"""
import numpy as np
import pytest

from oracle import ssg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ssg():
    import ssg_b200
    return ssg_b200


def _sym(n, seed, diag_scale=0.5, dtype=np.float64):
    rng = np.random.RandomState(seed)
    A = rng.rand(n, n)
    D = np.minimum(A, A.T)
    np.fill_diagonal(D, rng.rand(n) * diag_scale)
    return D.astype(dtype)


@pytest.mark.parametrize("n,eps", [(1, 0.5), (5, 0.9), (50, 0.2), (200, 0.08), (333, 0.03), (1024, 0.01),
                                   (2000, 0.004)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dbscan_matches_sklearn_bit_exact(ssg, n, eps, dtype):
    from sklearn.cluster import DBSCAN
    D = _sym(n, n, dtype=dtype)
    eps = np.float64(eps)
    want = DBSCAN(eps=eps, min_samples=4, metric="precomputed").fit(D)
    got = ssg.DBSCAN(eps=eps, min_samples=4, metric="precomputed", n_jobs=8).fit(D)
    assert got.labels_.dtype == np.int64
    assert np.array_equal(got.labels_, want.labels_)
    assert np.array_equal(got.core_sample_indices_, want.core_sample_indices_)


def test_dbscan_on_device_tensor_and_reuse(ssg):
    import torch
    from sklearn.cluster import DBSCAN
    est = ssg.DBSCAN(eps=0.05, min_samples=4, metric="precomputed")
    for seed in (1, 2):                      # the reference re-uses the estimator (selftraining.py:296-298)
        D = _sym(600, seed)
        want = DBSCAN(eps=0.05, min_samples=4, metric="precomputed").fit_predict(D)
        assert np.array_equal(est.fit_predict(torch.from_numpy(D).cuda()), want)
        assert np.array_equal(est.fit_predict(D), want)


def test_dbscan_everything_is_a_neighbour(ssg):
    """eps above the maximum: one cluster, neighbour list = n*n entries (capacity growth path)."""
    D = _sym(300, 3)
    lab = ssg.DBSCAN(eps=2.0, min_samples=4, metric="precomputed").fit_predict(D)
    assert np.array_equal(lab, np.zeros(300, np.int64))
    # nothing is a neighbour: the estimator rejects eps <= 0 as sklearn does; the kernel-level entry labels all noise
    with pytest.raises(ValueError):
        ssg.DBSCAN(eps=-1.0, min_samples=4, metric="precomputed").fit_predict(D)
    assert np.array_equal(ssg.dbscan_labels(D, -1.0, 4), -np.ones(300, np.int64))


def test_dbscan_on_rerank_output(ssg):
    from sklearn.cluster import DBSCAN
    tgt, _ = O.synth_features(1500, 256, 4)
    src, _ = O.synth_features(900, 256, 5, noise=0.6)
    _, f = ssg.re_ranking(src, tgt, lambda_value=0.1)
    for rho in (1.6e-3, 1e-2):
        eps = O.eps_estimate(f, rho)
        want = DBSCAN(eps=eps, min_samples=4, metric="precomputed", n_jobs=8).fit_predict(f)
        assert np.array_equal(ssg.DBSCAN(eps=eps, min_samples=4, metric="precomputed").fit_predict(f), want)
        assert want.max() >= 1


@pytest.mark.parametrize("n,rho", [(2, 0.5), (40, 0.05), (300, 1.6e-3), (300, 0.3), (1111, 1.6e-3), (64, 1.0)])
def test_eps_matches_numpy(ssg, n, rho):
    D = _sym(n, 100 + n)
    D[D < 0.02] = 0.0                       # exact zeros are dropped by np.nonzero (selftraining.py:290)
    D = np.minimum(D, D.T)
    tri = np.triu(D, 1)
    tri = np.sort(tri[np.nonzero(tri)], axis=None)
    top = int(np.round(rho * tri.size))
    got = ssg.eps_estimate(D, rho)
    if top == 0:
        assert np.isnan(got)
    else:
        want = tri[:top].mean()
        assert abs(got - want) <= 1e-12 * abs(want)


def test_eps_heavy_ties_and_float32(ssg):
    import torch
    rng = np.random.RandomState(0)
    vals = np.array([0.25, 0.5, 0.75, 0.899999976158142, 1.0])
    A = vals[rng.randint(0, 5, (500, 500))]
    D = np.minimum(A, A.T)
    for rho in (1e-3, 0.2, 0.41, 0.9):
        assert abs(ssg.eps_estimate(D, rho) - O.eps_estimate(D, rho)) < 1e-12
    D32 = D.astype(np.float32)
    got = ssg.eps_estimate(torch.from_numpy(D32).cuda(), 0.3)
    assert abs(got - O.eps_estimate(D32.astype(np.float64), 0.3)) < 1e-9


def test_full_size_cycle_labels(ssg):
    """N = 16 702: eps and labels from the device-resident matrix equal numpy / sklearn on its host copy."""
    import torch
    from sklearn.cluster import DBSCAN
    n, d = 16702, 2048
    tgt, _ = O.synth_features(n, d, 0)
    src, _ = O.synth_features(n, d, 1, noise=0.6)
    _, f = ssg.re_ranking_device(torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda(), lambda_value=0.1)
    plan = ssg.ClusterPlan(n)
    eps, top = plan.eps(f, 1.6e-3)
    labels, ncl = plan.dbscan(f, eps, 4)
    fh = f.cpu().numpy()
    assert top == 223152                                   # round(rho*M)
    want_eps = O.eps_estimate(fh, 1.6e-3)
    assert abs(eps - want_eps) <= 1e-12 * want_eps
    want = DBSCAN(eps=eps, min_samples=4, metric="precomputed", n_jobs=8).fit_predict(fh)
    assert np.array_equal(labels.cpu().numpy(), want)
    assert ncl == want.max() + 1 and ncl > 100


@pytest.mark.parametrize("m,n,nid,quant", [(60, 300, 25, False), (120, 500, 40, True)])
def test_cmc_and_mean_ap_match_reference_restatement(m, n, nid, quant):
    """reid/evaluation_metrics/ranking.py on the GPU vs the numpy + sklearn restatement (quantised distances = ties)."""
    import torch
    from reid.evaluation_metrics import cmc, mean_ap
    rng = np.random.RandomState(m)
    qid, gid = rng.randint(0, nid, m), rng.randint(0, nid, n)
    qcam, gcam = rng.randint(0, 3, m), rng.randint(0, 3, n)
    d = rng.rand(m, n).astype(np.float32) + 0.5 * (qid[:, None] != gid[None, :])
    if quant:
        d = np.round(d * 20) / 20
    dist = torch.from_numpy(d)
    want_map = O.mean_ap(d, qid, gid, qcam, gcam)
    assert abs(mean_ap(dist, qid, gid, qcam, gcam) - want_map) < 1e-9
    if not quant:      # with tied distances the rank order of ties is unspecified in the reference (np.argsort)
        for fmb in (True, False):
            got = cmc(dist, qid, gid, qcam, gcam, topk=50, first_match_break=fmb)
            np.testing.assert_allclose(got, O.cmc(d, qid, gid, qcam, gcam, topk=50, first_match_break=fmb), atol=1e-12)
    # the library's own kernel (ssg_rank_metrics) orders ties by gallery index = the stable argsort of the restatement
    for sep in (False, True):
        for fmb in (True, False):
            got = cmc(dist.cuda().double(), qid, gid, qcam, gcam, topk=50, first_match_break=fmb, separate_camera_set=sep)
            want = O.cmc(d, qid, gid, qcam, gcam, topk=50, first_match_break=fmb, separate_camera_set=sep)
            np.testing.assert_allclose(got, want, atol=1e-12)


def test_rank_metrics_kernel_at_evaluation_size():
    """Market-1501-shaped evaluation (3 368 queries x 15 913 gallery entries, 6 cameras, ~750 ids): the kernel against
    the restatement on a sample of the queries (the restatement is a per-query Python loop)."""
    import torch
    from reid.evaluation_metrics import ranking
    rng = np.random.RandomState(7)
    m, n, nid = 3368, 15913, 750
    qid, gid = rng.randint(0, nid, m), rng.randint(0, nid + 50, n)
    qcam, gcam = rng.randint(0, 6, m), rng.randint(0, 6, n)
    d = torch.rand(m, n, generator=torch.Generator().manual_seed(1)) + 0.3 * torch.from_numpy(qid[:, None] != gid[None, :])
    ap, nm, slots = ranking.rank_metrics(d.cuda(), qid, gid, qcam, gcam)
    sel = rng.choice(m, 40, replace=False)
    dn = d.numpy()
    for i in sel:
        one = O.mean_ap(dn[i:i + 1], qid[i:i + 1], gid, qcam[i:i + 1], gcam)
        assert abs(ap[i] - one) < 1e-12
        c = O.cmc(dn[i:i + 1], qid[i:i + 1], gid, qcam[i:i + 1], gcam, topk=n, first_match_break=True)
        assert int(np.argmax(c > 0)) == int(slots[i, :nm[i]].min())
    assert abs(ranking.mean_ap(d.cuda(), qid, gid, qcam, gcam) - float(np.mean(ap[nm > 0]))) < 1e-15


def test_dbscan_and_eps_property_based(ssg):
    """Property test: random symmetric matrices with heavy ties (quantised values, non-zero diagonal),
    random eps / min_samples / rho — labels, core samples and eps must equal sklearn / numpy every time."""
    from hypothesis import given, settings, strategies as st
    from sklearn.cluster import DBSCAN

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(n=st.integers(1, 150), seed=st.integers(0, 10 ** 6), levels=st.integers(2, 40),
           eps_q=st.integers(1, 40), min_samples=st.integers(1, 7), rho=st.floats(0.001, 1.0))
    def run(n, seed, levels, eps_q, min_samples, rho):
        rng = np.random.RandomState(seed)
        A = rng.randint(0, levels, (n, n)).astype(np.float64) / levels
        D = np.minimum(A, A.T)
        np.fill_diagonal(D, rng.randint(0, levels, n) / levels)
        eps = eps_q / 40.0
        want = DBSCAN(eps=eps, min_samples=min_samples, metric="precomputed").fit(D)
        got = ssg.DBSCAN(eps=eps, min_samples=min_samples, metric="precomputed").fit(D)
        assert np.array_equal(got.labels_, want.labels_)
        assert np.array_equal(got.core_sample_indices_, want.core_sample_indices_)
        tri = np.triu(D, 1)
        tri = np.sort(tri[np.nonzero(tri)], axis=None)
        top = int(np.round(rho * tri.size))
        e = ssg.eps_estimate(D, rho)
        if top == 0:
            assert np.isnan(e)
        else:
            assert abs(e - tri[:top].mean()) <= 1e-12 * max(abs(tri[:top].mean()), 1e-300)
    run()
    D = np.zeros((4, 4))
    for bad in (0.0, -1.0, float("nan")):            # sklearn: InvalidParameterError (a ValueError) at fit()
        with pytest.raises(ValueError):
            DBSCAN(eps=bad, min_samples=4, metric="precomputed").fit(D)
        with pytest.raises(ValueError):
            ssg.DBSCAN(eps=bad, min_samples=4, metric="precomputed").fit(D)
