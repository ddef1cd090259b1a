"""CPU, build container only: the restatement against the UNMODIFIED reference, live, bit-for-bit."""
import numpy as np
import pytest

from oracle import refshim, ssg_oracle as O

pytestmark = pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")


@pytest.mark.parametrize("mode", ["f32", "ref"])
@pytest.mark.parametrize("n,ns,d,seed", [(120, 90, 128, 0), (200, 200, 512, 1)])
def test_re_ranking_bit_exact(mode, n, ns, d, seed):
    tgt, _ = O.synth_features(n, d, seed)
    src, _ = O.synth_features(ns, d, seed + 50)
    e0, f0 = refshim.ref_re_ranking(src, tgt, mode=mode, lambda_value=0.1)
    e1, f1 = O.re_ranking(src, tgt, lambda_value=0.1, mode=mode)
    assert np.array_equal(e0, e1) and np.array_equal(f0, f1)


@pytest.mark.parametrize("n,ns", [(2, 1), (3, 5), (10, 7), (21, 30), (22, 22)])
def test_re_ranking_bit_exact_below_k1_targets(n, ns):
    """Fewer targets than the k1 + 1 = 21 rank columns reid/rerank.py:76 slices: the reference just works with the
    shorter rows; pin that edge (the CUDA parity case for it is tests/test_gpu_next_variants.py)."""
    rng = np.random.RandomState(n * 31 + ns)
    tgt, src = rng.randn(n, 64).astype(np.float32), rng.randn(ns, 64).astype(np.float32)
    _, f0 = refshim.ref_re_ranking(src, tgt, mode="f32", lambda_value=0.1)
    _, f1 = O.re_ranking(src, tgt, lambda_value=0.1, mode="f32")
    assert np.array_equal(f0, f1)


def test_no_rerank_returns_none():
    tgt, _ = O.synth_features(40, 32, 0)
    e0, f0 = refshim.ref_re_ranking(tgt, tgt, mode="f32", no_rerank=True)
    e1, f1 = O.re_ranking(tgt, tgt, mode="f32", no_rerank=True)
    assert f0 is None and f1 is None and np.array_equal(e0, e1)


def test_re_ranking_init_bit_exact():
    f, _ = O.synth_features(100, 128, 2, per_cluster=10)
    q, g = f[:30], f[30:]
    a = refshim.ref_re_ranking_init(q @ g.T, q @ q.T, g @ g.T, stable=True)
    b = O.re_ranking_init(q @ g.T, q @ q.T, g @ g.T)
    assert np.array_equal(a, b)


def test_reference_model_matches_oracle_model():
    import torch, warnings
    from oracle import resnet_oracle as R
    ref = refshim.load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ref.models.create("resnet50", num_classes=0, num_split=2, pretrained=False)
    m.base.load_state_dict(R.make_state_dict(0), strict=False)
    m.eval()
    o = R.build_model(2, 0)
    x = R.synth_images(2, 99)
    with torch.no_grad():
        a = m(x, False)[0]
        b = o(x, False)[0]
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_evaluation_metrics_restatement_matches_reference():
    """oracle cmc / mean_ap (row f2) against reid/evaluation_metrics/ranking.py of the unmodified reference."""
    ref = refshim.load_reference()
    with refshim._reference_on_path():
        import reid.evaluation_metrics as rem
    rng = np.random.RandomState(0)
    m, n = 40, 200
    qid, gid = rng.randint(0, 20, m), rng.randint(0, 20, n)
    qc, gc = rng.randint(0, 3, m), rng.randint(0, 3, n)
    d = rng.rand(m, n).astype(np.float32) + 0.5 * (qid[:, None] != gid[None, :])
    assert rem.mean_ap(d, qid, gid, qc, gc) == O.mean_ap(d, qid, gid, qc, gc)
    for fmb in (True, False):
        a = rem.cmc(d, qid, gid, qc, gc, first_match_break=fmb)
        b = O.cmc(d, qid, gid, qc, gc, first_match_break=fmb)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", ["f32", "ref"])
@pytest.mark.parametrize("n,ns,d,seed,noise", [(90, 70, 64, 0, 0.5), (150, 120, 256, 1, 0.5), (64, 40, 32, 2, 0.0)])
def test_rerank_plain_restatement_bit_exact(mode, n, ns, d, seed, noise):
    """Row f4 (oracle only so far): reid/rerank_plain.py:127-178 through neighbour lists / set intersections, against
    the unmodified reference (dense boolean matrix + scipy's boolean Jaccard); noise 0 = duplicate features = ties at
    the k-th distance."""
    import contextlib
    import os
    from oracle import rerank_plain_oracle as P
    refshim.load_reference()
    with refshim._reference_on_path():
        import reid.rerank_plain as RP
    tgt, _ = O.synth_features(n, d, seed, noise=noise)
    src, _ = O.synth_features(ns, d, seed + 9)
    ctx = refshim.f32_stable(RP) if mode == "f32" else contextlib.nullcontext()
    with ctx, contextlib.redirect_stdout(open(os.devnull, "w")):
        want, again = RP.re_ranking(src, tgt, k=20, lambda_value=0.1)
    assert want is again
    got = P.re_ranking_plain(src, tgt, 20, 0.1, mode)
    assert got.dtype == want.dtype and np.array_equal(got, want)


@pytest.mark.parametrize("mode", ["f32", "ref"])
@pytest.mark.parametrize("n,ns,d,seed", [(90, 70, 64, 0), (64, 40, 32, 2)])
def test_rerank_lh_restatement_bit_exact(mode, n, ns, d, seed):
    """Row f4, second function: reid/rerank_plain.py:27-123 re_ranking_lh against the unmodified reference."""
    import contextlib
    import os
    from oracle import rerank_plain_oracle as P
    refshim.load_reference()
    with refshim._reference_on_path():
        import reid.rerank_plain as RP
    tgt, _ = O.synth_features(n, d, seed)
    src, _ = O.synth_features(ns, d, seed + 9)
    ctx = refshim.f32_stable(RP) if mode == "f32" else contextlib.nullcontext()
    with ctx, contextlib.redirect_stdout(open(os.devnull, "w")):
        want = RP.re_ranking_lh(src, tgt, k1=20, k2=6, lambda_value=0.2)
    want = want[-1] if isinstance(want, tuple) else want
    got = P.re_ranking_lh(src, tgt, 20, 6, 0.2, mode)
    assert got.dtype == want.dtype and np.array_equal(got, want)


@pytest.mark.parametrize("sep", [False, True])
def test_cmc_and_mean_ap_oracle_against_the_reference_ranking_module(sep):
    """oracle cmc / mean_ap vs reid/evaluation_metrics/ranking.py:18-115 as the reference's evaluators module imports
    them (distinct distances: the reference's argsort is unstable under ties)."""
    ref = refshim.load_reference()
    rng = np.random.RandomState(5)
    m, n, nid = 70, 400, 30
    qid, gid = rng.randint(0, nid, m), rng.randint(0, nid, n)
    qcam, gcam = rng.randint(0, 3, m), rng.randint(0, 3, n)
    d = rng.rand(m, n) + 0.5 * (qid[:, None] != gid[None, :])
    assert abs(ref.evaluators.mean_ap(d, qid, gid, qcam, gcam) - O.mean_ap(d, qid, gid, qcam, gcam)) < 1e-15
    for fmb in (False, True):
        want = ref.evaluators.cmc(d, qid, gid, qcam, gcam, topk=40, separate_camera_set=sep, first_match_break=fmb)
        got = O.cmc(d, qid, gid, qcam, gcam, topk=40, separate_camera_set=sep, first_match_break=fmb)
        assert np.array_equal(want, got)
