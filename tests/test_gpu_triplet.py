"""GPU: the CUDA triplet loss (csrc/triplet.cu through ssg_triplet_forward/backward, row f1) against the oracle, the
reference's golden vectors, and torch autograd of the same expression; then one FinedTrainer2 step."""
import os

import numpy as np
import pytest

from oracle import triplet_oracle as TO

pytestmark = pytest.mark.gpu
RTOL = 1e-5   # stated tolerance of this row (float32 distances; see oracle/triplet_oracle.py)


@pytest.fixture(scope="module")
def torch_():
    import torch
    return torch


def _run(torch, x, t, K, margin, semi):
    import ssg_b200
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    loss, prec, dist = ssg_b200.triplet_loss(xt, torch.from_numpy(t).cuda(), K, margin, semi, return_dist=True)
    loss.backward()
    return float(loss), float(prec), xt.grad.cpu().numpy(), dist.cpu().numpy()


def test_against_reference_goldens(torch_, golden_dir):
    g = np.load(os.path.join(golden_dir, "triplet_cases.npz"))
    for ci, row in enumerate(g["cases"]):
        K, margin, semi = int(row[1]), float(row[4]), bool(row[5])
        loss, prec, grad, _ = _run(torch_, g["x_%d" % ci], g["t_%d" % ci], K, margin, semi)
        assert abs(loss - float(g["loss_%d" % ci])) <= RTOL * max(1.0, abs(loss)), ci
        assert abs(prec - float(g["prec_%d" % ci])) < 1e-6, ci
        ref = g["grad_%d" % ci]
        assert np.abs(grad - ref).max() <= RTOL * np.abs(ref).max(), ci


@pytest.mark.parametrize("P,K,d,seed,margin,semi,extra", [
    (1, 2, 1, 0, 0.5, True, 1), (32, 4, 2048, 1, 0.5, True, 0), (16, 8, 2048, 2, 0.3, True, 0),
    (7, 3, 130, 3, 0.0, True, 5), (64, 4, 2048, 4, 0.5, True, 0), (33, 4, 257, 5, 0.5, False, 0),
    (256, 4, 512, 6, 0.5, True, 0), (1024, 4, 64, 7, 0.5, True, 0)])
def test_against_oracle(torch_, P, K, d, seed, margin, semi, extra):
    x, t = TO.synth_batch(P, K, d, seed, 0.3, extra)
    loss, prec, grad, dist = _run(torch_, x, t, K, margin, semi)
    dref, _ = TO.pairwise_dist(x)
    off = ~np.eye(len(t), dtype=bool)
    assert np.abs(dist - dref)[off].max() <= 1e-6 * dref.max()
    assert np.all(np.diag(dist) == np.float32(1e-6))          # clamped diagonal (triplet.py:31)
    l, p, gr = TO.triplet_loss(x, t, K, margin, semi, with_grad=True)
    assert abs(loss - l) <= RTOL * max(1.0, abs(l))
    assert abs(prec - p) < 1e-6
    assert np.abs(grad - gr).max() <= RTOL * max(np.abs(gr).max(), 1e-30)


def test_matches_torch_autograd_and_is_deterministic(torch_):
    torch = torch_
    import ssg_b200
    x, t = TO.synth_batch(32, 4, 2048, 9, 0.3)
    xt = torch.from_numpy(x).cuda().double().requires_grad_(True)
    tt = torch.from_numpy(t).cuda()
    dist = (xt[:, None, :] - xt[None, :, :]).pow(2).sum(-1).clamp(min=1e-12).sqrt()
    mask = tt[None, :] == tt[:, None]
    an = dist.masked_fill(mask, float("inf")).min(dim=1).values
    K = 4
    aps, ans = [], []
    for a in range(len(t)):
        for p in range(a + 1, (a // K + 1) * K):
            aps.append(dist[a, p]); ans.append(an[a])
    ref = torch.relu(torch.stack(aps) - torch.stack(ans) + 0.5).mean()
    ref.backward()
    outs = []
    for _ in range(2):
        x32 = torch.from_numpy(x).cuda().requires_grad_(True)
        loss, prec = ssg_b200.triplet_loss(x32, tt, K, 0.5)
        (3.0 * loss).backward()                                 # upstream gradient is honoured
        outs.append((loss.detach().clone(), x32.grad.clone()))
    assert abs(float(outs[0][0]) - float(ref)) <= RTOL * float(ref)
    g64 = xt.grad.float() * 3.0
    assert float((outs[0][1] - g64).abs().max()) <= RTOL * float(g64.abs().max())
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])   # bit-reproducible


def test_error_behaviour(torch_):
    torch = torch_
    import ssg_b200
    x, t = TO.synth_batch(2, 4, 16, 0)
    xc = torch.from_numpy(x).cuda()
    with pytest.raises(RuntimeError):     # the reference raises at neg_examples.min() of an empty tensor
        ssg_b200.triplet_loss(xc, torch.zeros(8, dtype=torch.int64).cuda(), 4, 0.5)
    with pytest.raises(ValueError):       # the reference raises at torch.cat([]) (num_instances = 1)
        ssg_b200.triplet_loss(xc, torch.from_numpy(t).cuda(), 1, 0.5)
    with pytest.raises(ValueError):
        ssg_b200.triplet_loss(xc, torch.from_numpy(t[:5]).cuda(), 4, 0.5)
    loss, _ = ssg_b200.triplet_loss(xc, torch.zeros(8, dtype=torch.int64).cuda(), 4, 0.5, check=False)
    assert torch.isnan(loss)


def test_fined_trainer2_step(torch_):
    """One fine-tune step (trainers.py:204-271) on a tiny batch: loss equals the oracle's aggregation evaluated on the
    model's own outputs, and SGD moves the trunk weights."""
    torch = torch_
    from reid import models
    from reid.loss import TripletLoss
    from reid.trainers import FinedTrainer2
    torch.manual_seed(0)
    model = models.create("resnet50", num_classes=0, num_split=2, pretrained=False).cuda()
    for m in model.modules():                       # the reference init (std 1e-3) gives ~0 features; use a live init
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight)
    crit = [TripletLoss(margin=0.5, num_instances=2).cuda(), TripletLoss(margin=0.5, num_instances=2).cuda()]
    trainer = FinedTrainer2(model, crit)
    imgs = torch.randn(8, 3, 256, 128)
    pids = [torch.arange(4).repeat_interleave(2) for _ in range(3)]
    batch = (imgs, ["f%d" % i for i in range(8)], pids, torch.ones(8))
    model.train()
    inputs, tg, _ = trainer._parse_data(batch)
    with torch.no_grad():
        x1, x2 = model(*inputs)
    want, want_prec = TO.fined_trainer2_loss(x2.cpu().numpy(), [b.cpu().numpy() for b in x1],
                                             [p.numpy() for p in pids], 2, 0.5)
    loss, prec = trainer._forward(inputs, tg, 0)
    assert abs(float(loss) - want) <= 1e-4 * max(1.0, abs(want))      # BN batch statistics: same batch, same mode
    assert abs(float(prec) - want_prec) < 1e-6
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    w0 = model.base.conv1.weight.detach().clone()
    trainer.train(0, [batch], opt, print_freq=1)
    assert not torch.equal(w0, model.base.conv1.weight.detach())
