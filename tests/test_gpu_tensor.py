"""GPU: the tcgen05 distance GEMM (bf16x3 split) and the tensor distance mode of the re-ranking path.

The GEMM output is an approximation used only to pick candidates, so it is checked against the error
bound the candidate verification relies on (tensor_eps_rel(d) * (|x|^2 + |y|^2), api.cu); the
re-ranking results in tensor mode must equal the exact mode bit for bit (ranks) / to 1e-6 (values).
"""
import numpy as np
import pytest

from oracle import ssg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ssg():
    import ssg_b200
    return ssg_b200


@pytest.mark.parametrize("nx,ny,d", [(128, 128, 64), (130, 257, 64), (1000, 777, 2048), (64, 3000, 512),
                                      (2500, 2500, 2048)])
def test_tensor_sqdist_within_error_bound(ssg, nx, ny, d):
    import torch
    from ssg_b200 import _lib
    rng = np.random.RandomState(nx + ny)
    x = rng.randn(nx, d).astype(np.float32)
    y = rng.randn(ny, d).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    y[: min(nx, ny) // 2] = x[: min(nx, ny) // 2]            # exact duplicates: distance 0
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    approx = ssg.sqdist(xt, yt, mode=_lib.DIST_TENSOR)
    exact = ssg.sqdist(xt, yt, mode=_lib.DIST_EXACT)
    err = float((approx - exact).abs().max())
    bound = (1.15e-5 + 2.24e-8 * d) * 2.0                     # api.cu tensor_eps_rel(d) at unit norms
    print("max |approx-exact| = %.3e, certified bound %.3e" % (err, bound))
    assert err < 0.8 * bound, err


def test_tensor_sqdist_unnormalised_scales_with_norms(ssg):
    import torch
    from ssg_b200 import _lib
    rng = np.random.RandomState(5)
    x = (rng.randn(300, 256) * 7).astype(np.float32)
    y = (rng.randn(500, 256) * 0.3).astype(np.float32)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    approx = ssg.sqdist(xt, yt, mode=_lib.DIST_TENSOR)
    exact = ssg.sqdist(xt, yt, mode=_lib.DIST_EXACT)
    bound = (1.15e-5 + 2.24e-8 * 256) * ((xt * xt).sum(1, keepdim=True) + (yt * yt).sum(1).max())
    assert bool(((approx - exact).abs() < 0.8 * bound).all())


@pytest.mark.parametrize("n,ns,d,seed", [(257, 300, 2048, 1), (1000, 900, 512, 3), (3000, 2000, 2048, 4)])
def test_tensor_mode_equals_exact_mode(ssg, n, ns, d, seed):
    import torch
    from ssg_b200 import _lib
    tgt, _ = O.synth_features(n, d, seed)
    src, _ = O.synth_features(ns, d, seed + 77, noise=0.6)
    t, s = torch.from_numpy(tgt).cuda(), torch.from_numpy(src).cuda()
    plan = ssg.RerankPlan(n, ns, d)
    _, f_ex = plan.run(s, t, lambda_value=0.1, dist_mode=_lib.DIST_EXACT)
    torch.cuda.synchronize()
    st_ex = {k: plan.stage(k, n) for k in (_lib.STAGE_VEC, _lib.STAGE_ROWMAX, _lib.STAGE_RANK, _lib.STAGE_RANK_VAL)}
    f_ex = f_ex.clone()
    _, f_tc = plan.run(s, t, lambda_value=0.1, dist_mode=_lib.DIST_TENSOR)
    torch.cuda.synchronize()
    assert np.array_equal(plan.stage(_lib.STAGE_ROWMAX, n), st_ex[_lib.STAGE_ROWMAX])
    assert np.array_equal(plan.stage(_lib.STAGE_RANK, n)[:, :21], st_ex[_lib.STAGE_RANK][:, :21])
    assert np.array_equal(plan.stage(_lib.STAGE_RANK_VAL, n)[:, :21], st_ex[_lib.STAGE_RANK_VAL][:, :21])
    assert np.array_equal(plan.stage(_lib.STAGE_VEC, n), st_ex[_lib.STAGE_VEC])
    assert torch.equal(f_tc, f_ex)
    flagged = int(plan.stage(_lib.STAGE_FLAGGED, n)[0])
    assert flagged < n // 4, flagged


def test_tensor_mode_heavy_ties_take_the_fallback(ssg, golden_dir):
    """Duplicate features: the error bound cannot certify ties, every row must fall back and still match."""
    import os
    from ssg_b200 import _lib
    g = np.load(os.path.join(golden_dir, "rerank_n96_d64_ties.npz"))
    e, f = ssg.re_ranking(g["src"], g["tgt"], lambda_value=float(g["lam"]), dist_mode=_lib.DIST_TENSOR)
    np.testing.assert_allclose(f, g["final_f32"], rtol=0, atol=1e-4)


def test_tensor_mode_full_size_matches_exact(ssg):
    import torch
    from ssg_b200 import _lib
    n, d = 16702, 2048
    tgt, _ = O.synth_features(n, d, 0)
    src, _ = O.synth_features(n, d, 1, noise=0.6)
    t, s = torch.from_numpy(tgt).cuda(), torch.from_numpy(src).cuda()
    plan = ssg.RerankPlan(n, n, d)
    _, f_ex = plan.run(s, t, lambda_value=0.1, dist_mode=_lib.DIST_EXACT)
    torch.cuda.synchronize()
    rank_ex = plan.stage(_lib.STAGE_RANK, n)
    f_ex = f_ex.clone()
    _, f_tc = plan.run(s, t, lambda_value=0.1, dist_mode=_lib.DIST_TENSOR)
    torch.cuda.synchronize()
    assert np.array_equal(plan.stage(_lib.STAGE_RANK, n)[:, :21], rank_ex[:, :21])
    assert torch.equal(f_tc, f_ex)
    print("flagged rows:", int(plan.stage(_lib.STAGE_FLAGGED, n)[0]))


def test_duke_size_row_blocked_path(ssg):
    """N = 36 411 (DukeMTMC shape, BASELINE.json configs[3]): the fp32 distance matrix (5.3 GB) exceeds the 4 GiB
    distance block, so the distance stages run in row blocks; size-independent properties + sampled rows vs cdist."""
    import torch
    from scipy.spatial.distance import cdist
    from ssg_b200 import _lib
    n, d, lam = 36411, 2048, 0.1
    g = torch.Generator(device="cuda").manual_seed(3)
    c = n // 20
    centres = torch.randn(c, d, generator=g, device="cuda")
    lab = torch.randint(0, c, (n,), generator=g, device="cuda")
    t = centres[lab] + 0.5 * torch.randn(n, d, generator=g, device="cuda")
    t = (t / t.norm(dim=1, keepdim=True)).contiguous()
    s = centres[torch.randint(0, c, (n,), generator=g, device="cuda")] + 0.6 * torch.randn(n, d, generator=g, device="cuda")
    s = (s / s.norm(dim=1, keepdim=True)).contiguous()
    plan = ssg.RerankPlan(n, n, d)
    _, f = plan.run(s, t, lambda_value=lam, dist_mode=_lib.DIST_TENSOR)
    torch.cuda.synchronize()
    rank = plan.stage(_lib.STAGE_RANK, n)
    assert np.array_equal(rank[:, 0], np.arange(n))
    rows = np.arange(0, n, n // 48)[:48]
    th = t.cpu().numpy()
    od = np.power(cdist(th[rows], th).astype(np.float32), 2).astype(np.float32)
    odn = od / od.max(axis=1, keepdims=True)
    assert np.array_equal(np.argsort(odn, kind="stable")[:, :21], rank[rows, :21])
    assert np.array_equal(plan.stage(_lib.STAGE_ROWMAX, n)[rows], od.max(axis=1))
    # symmetric on a sampled block (a full transpose of the 10.6 GB matrix is not needed)
    blk = torch.arange(0, n, 37, device="cuda")
    sub = f.index_select(0, blk).index_select(1, blk)
    assert torch.equal(sub, sub.t())
    cplan = ssg.ClusterPlan(n)
    eps, top = cplan.eps(f, 1.6e-3)
    assert top == 1060580                                  # round(rho * M) at N = 36 411
    labels, ncl = cplan.dbscan(f, eps, 4)
    assert ncl > 500 and int((labels >= 0).sum()) > n // 2
    del f, plan, cplan
    torch.cuda.empty_cache()
