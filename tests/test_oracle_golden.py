"""CPU: the oracle restatement against the committed reference outputs (tests/golden/).

The goldens were produced by the unmodified reference in the build container
(oracle/make_goldens.py).  Float outputs are compared with a small tolerance rather than bit-for-bit
because numpy's float32 exp / pairwise sums are SIMD-dispatch dependent across host CPUs; integer
outputs (ranks on tie-free data, DBSCAN labels on the golden matrix) are exact.
"""
import os

import numpy as np
import pytest

from oracle import ssg_oracle as O

CASES = ["rerank_n160_d256.npz", "rerank_n257_d2048.npz", "rerank_n96_d64_ties.npz"]


@pytest.mark.parametrize("case", CASES)
def test_re_ranking_f32_matches_reference(golden_dir, case):
    g = np.load(os.path.join(golden_dir, case))
    st = {}
    e, f = O.re_ranking(g["src"], g["tgt"], lambda_value=float(g["lam"]), mode="f32", stages=st)
    assert f.dtype == np.float64 and e.dtype == np.float32
    np.testing.assert_allclose(e, g["euclid_f32"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(st["vec"], g["vec_f32"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(f, g["final_f32"], rtol=0, atol=5e-6)
    assert np.array_equal(f, f.T)


@pytest.mark.parametrize("case", CASES[:2])
def test_re_ranking_fp16_mode_close_to_reference(golden_dir, case):
    g = np.load(os.path.join(golden_dir, case))
    e, f = O.re_ranking(g["src"], g["tgt"], lambda_value=float(g["lam"]), mode="ref")
    assert e.dtype == np.float16
    # fp16 exp / unstable argsort are CPU dependent: statistical agreement only
    assert np.mean(np.abs(f - g["final_ref"]) < 2e-3) > 0.97


@pytest.mark.parametrize("case", CASES)
def test_eps_and_dbscan_on_golden_matrix(golden_dir, case):
    from sklearn.cluster import DBSCAN
    g = np.load(os.path.join(golden_dir, case))
    D = g["final_f32"]
    for bi, rho in enumerate(g["rhos"]):
        eps = O.eps_estimate(D, rho)
        assert abs(eps - float(g["eps_%d" % bi])) < 1e-12
        eps = float(g["eps_%d" % bi])
        lab = g["labels_%d" % bi]
        assert np.array_equal(O.dbscan_dfs(D, eps), lab)
        assert np.array_equal(O.dbscan_components(D, eps), lab)
        assert np.array_equal(DBSCAN(eps=eps, min_samples=4, metric="precomputed").fit_predict(D), lab)


def test_dbscan_restatements_random_symmetric():
    from sklearn.cluster import DBSCAN
    rng = np.random.RandomState(1)
    for n, thr in [(50, 0.2), (200, 0.08), (333, 0.03)]:
        A = rng.rand(n, n)
        D = np.minimum(A, A.T)
        np.fill_diagonal(D, rng.rand(n) * 0.5)      # the reference's diagonal is non-zero (App. B.3)
        ref = DBSCAN(eps=thr, min_samples=4, metric="precomputed").fit_predict(D)
        assert np.array_equal(O.dbscan_dfs(D, thr), ref)
        assert np.array_equal(O.dbscan_components(D, thr), ref)


def test_re_ranking_init_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "rerank_init_q40_g90.npz"))
    out = O.re_ranking_init_features(g["qf"], g["gf"])
    np.testing.assert_allclose(out, g["final"], rtol=0, atol=5e-6)


def test_embed_oracle_matches_reference(golden_dir):
    import torch
    from oracle import resnet_oracle as R
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = np.load(os.path.join(golden_dir, "embed_4img.npz"))
    n = int(g["n_img"])
    imgs = R.synth_images(n, int(g["seed_img"]))
    names = ["im%03d" % i for i in range(n)]
    batches = [(imgs, names, list(range(n)), [0] * n)]
    for S in (1, 3):
        m = R.build_model(S, int(g["weight_seed"]))
        fl, _ = R.extract_features(m, batches, for_eval=False)
        fe, _ = R.extract_features(m, batches, for_eval=True)
        gl = g["list_S%d" % S]
        for b in range(gl.shape[0]):
            mine = torch.stack([fl[k] if S == 1 else fl[k][b] for k in names]).numpy()
            np.testing.assert_allclose(mine, gl[b], rtol=0, atol=2e-5)
        np.testing.assert_allclose(torch.stack([fe[k] for k in names]).numpy(), g["eval_S%d" % S],
                                   rtol=0, atol=2e-5)


def test_whole_path_golden_is_pinned_to_the_oracle(golden_dir):
    """tests/golden/cycle_n512_S2.npz (images -> labels through the unmodified reference): the oracle's ResNet
    restatement reproduces the stored head of the reference features from the same seeded images, and the stored
    eps / labels are what the oracle's eps estimator and DBSCAN give on the stored final_dist."""
    import torch
    from sklearn.metrics import adjusted_rand_score
    from oracle import resnet_oracle as R
    g = np.load(os.path.join(golden_dir, "cycle_n512_S2.npz"))
    n, S = int(g["n"]), int(g["num_split"])
    head = g["feat_tgt_head"]
    k = 8
    # the generator draws the images in order, so a prefix needs the whole set's generator state: draw the set
    imgs, ident = R.synth_identity_images(n, int(g["seed_tgt"]), int(g["per_identity"]), float(g["noise"]))
    assert np.array_equal(ident.numpy(), g["identity"])
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    model = R.build_model(S, int(g["weight_seed"]))
    names = ["i%d" % i for i in range(k)]
    f, _ = R.extract_features(model, [(imgs[:k], names, [0] * k, [0] * k)], for_eval=False)
    for b in range(S + 1):
        got = torch.stack([f[nm][b] for nm in names]).numpy()
        np.testing.assert_allclose(got, head[b, :k], rtol=0, atol=2e-6)
    iu = np.triu_indices(n)
    for b in range(S + 1):
        tri = g["final_f32_b%d" % b].astype(np.float64)
        D = np.zeros((n, n))
        D[iu] = tri
        D = D + np.triu(D, 1).T
        for ri, rho in enumerate(g["rhos"]):
            eps = O.eps_estimate(D, float(rho))
            want = float(g["eps_f32_r%d_b%d" % (ri, b)])
            assert abs(eps - want) <= 1e-6 * want              # the stored triangle is float32-rounded
            lab = O.dbscan_dfs(D, want, 4)
            assert adjusted_rand_score(g["labels_f32_r%d_b%d" % (ri, b)], lab) >= 0.995
