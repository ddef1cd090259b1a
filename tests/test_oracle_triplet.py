"""CPU: the triplet-loss restatement (oracle/triplet_oracle.py, row f1) against the reference's golden vectors and,
in the build container, against the unmodified reference module live."""
import os
import warnings

import numpy as np
import pytest

from oracle import refshim, triplet_oracle as TO

RTOL = 1e-5   # float32 GEMM rounding of the reference (reid/loss/triplet.py:27-31) vs float64 accumulation


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "triplet_cases.npz"))
    for ci, row in enumerate(g["cases"]):
        P, K, d, seed, margin, semi, extra, sep = row
        yield ci, g, int(K), float(margin), bool(semi)


def test_oracle_matches_reference_goldens(golden_dir):
    seen = 0
    for ci, g, K, margin, semi in _cases(golden_dir):
        x, t = g["x_%d" % ci], g["t_%d" % ci]
        loss, prec, grad = TO.triplet_loss(x, t, K, margin, semi, with_grad=True)
        assert abs(loss - float(g["loss_%d" % ci])) <= RTOL * max(1.0, abs(loss))
        assert abs(prec - float(g["prec_%d" % ci])) < 1e-6
        ref = g["grad_%d" % ci]
        assert np.abs(grad - ref).max() <= RTOL * np.abs(ref).max()
        seen += 1
    assert seen == 6


def test_oracle_error_cases():
    x, t = TO.synth_batch(2, 4, 16, 0)
    with pytest.raises(ValueError):
        TO.triplet_loss(x, np.zeros_like(t), 4, 0.5)          # no negatives anywhere (reference: min() of empty)
    with pytest.raises(ValueError):
        TO.triplet_loss(x, t, 1, 0.5)                         # K = 1: no pairs (reference: cat of an empty list)


def test_fined_trainer2_aggregation():
    x2, t = TO.synth_batch(4, 4, 64, 1)
    banks = [TO.synth_batch(4, 4, 64, 2 + b)[0] for b in range(3)]
    loss, prec = TO.fined_trainer2_loss(x2, banks, [t, t, t], 4, 0.5)
    parts = [TO.triplet_loss(x2, t, 4, 0.5)] + [TO.triplet_loss(b, t, 4, 0.5) for b in banks]
    assert loss == pytest.approx(sum(p[0] for p in parts)) and prec == parts[0][1]


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("P,K,d,seed,margin,semi,extra", [(8, 4, 128, 10, 0.5, True, 0), (6, 5, 300, 11, 0.2, True, 4),
                                                          (10, 4, 64, 12, 0.4, False, 0)])
def test_oracle_matches_reference_live(P, K, d, seed, margin, semi, extra):
    import torch
    refshim.load_reference()
    with refshim._reference_on_path():
        import reid.loss.triplet as T
    x, t = TO.synth_batch(P, K, d, seed, 0.3, extra)
    xt = torch.from_numpy(x).requires_grad_(True)
    crit = T.TripletLoss(margin=margin, num_instances=K, use_semi=semi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss, prec = crit(xt, torch.from_numpy(t), 0)
    loss.backward()
    l, p, g = TO.triplet_loss(x, t, K, margin, semi, with_grad=True)
    assert abs(l - loss.item()) <= RTOL * max(1.0, abs(l)) and abs(p - float(prec)) < 1e-6
    assert np.abs(g - xt.grad.numpy()).max() <= RTOL * np.abs(g).max()


def test_oracle_gradient_is_the_derivative_of_the_oracle_loss():
    """Central finite differences of the restated loss (float64 evaluation of the same expression) against the analytic
    gradient the CUDA kernels are compared with: the mining is piecewise constant, so away from ties they must agree."""
    x, t = TO.synth_batch(4, 3, 6, 21, 0.4, extra=1)

    def loss64(xx):
        sq = (xx * xx).sum(1)
        d = np.sqrt(np.maximum(sq[:, None] + sq[None, :] - 2.0 * xx @ xx.T, 1e-12))
        a, p, m = TO.mine(d, t, 3, True)
        return np.maximum(d[a, p] - d[a, m] + 0.5, 0.0).mean()

    _, _, g = TO.triplet_loss(x, t, 3, 0.5, True, with_grad=True)
    x64 = x.astype(np.float64)
    h = 1e-6
    num = np.zeros_like(x64)
    for i in range(x64.shape[0]):
        for j in range(x64.shape[1]):
            xp, xm = x64.copy(), x64.copy()
            xp[i, j] += h
            xm[i, j] -= h
            num[i, j] = (loss64(xp) - loss64(xm)) / (2 * h)
    assert np.abs(num - g).max() <= 1e-5 * max(np.abs(num).max(), 1.0)
