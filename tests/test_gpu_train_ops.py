"""B200: the training-side convolution operators (csrc/train.cu) at the shapes of the fine-tune
step -- reid/trainers.py:204-271 runs loss.backward() through the ResNet-50 convolutions of reid/models/resnet.py:52-70.
A floating-point kernel: the reference is torch (cuDNN, fp32, TF32 off) on the same bf16-rounded operands, and the
tolerances are written out below.  Called through the C ABI (ssg_b200._lib) and through the autograd wrappers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-similarity-grouping_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

# (B, H, W, cin, cout, k, stride): H, W = input map of the convolution for 256 x 128 images
SHAPES = [
    (8, 64, 32, 64, 64, 3, 1),        # layer1 conv2 (kernel-row-sharing kernel in the data gradient)
    (8, 64, 32, 64, 256, 1, 1),       # layer1 conv3
    (8, 64, 32, 256, 64, 1, 1),       # layer1 conv1 of the later blocks
    (8, 64, 32, 128, 128, 3, 2),      # layer2.0 conv2 (stride 2)
    (8, 64, 32, 256, 512, 1, 2),      # layer2.0 downsample (1x1 stride 2)
    (8, 32, 16, 128, 128, 3, 1),      # layer2 conv2
    (8, 16, 8, 256, 256, 3, 1),       # layer3 conv2
    (8, 16, 8, 1024, 256, 1, 1),      # layer3 conv1
    (8, 16, 8, 512, 512, 3, 2),       # layer4.0 conv2 (stride 2)
    (8, 8, 4, 512, 2048, 1, 1),       # layer4 conv3
    (5, 8, 4, 512, 512, 3, 1),        # layer4 conv2, odd batch (tiles of several images, ragged last tile)
]


@pytest.mark.parametrize("B,H,W,cin,cout,k,stride", SHAPES)
def test_dgrad_and_wgrad_against_cudnn_fp32(B, H, W, cin, cout, k, stride):
    import torch
    from ssg_b200 import train
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + cin + cout + k + stride)
    x = torch.randn(B, H, W, cin, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / float(np.sqrt(cin * k * k))
    dy = torch.randn(B, H // stride, W // stride, cout, device="cuda", generator=g).to(torch.bfloat16)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.to(torch.bfloat16).float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, None, stride, k // 2)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    xo = x.clone().requires_grad_(True)
    wo = w.clone().requires_grad_(True)
    yo = train.conv2d_nhwc(xo, wo, stride)
    yo.backward(dy)
    torch.cuda.synchronize()
    want_y, want_dx = yr.detach().permute(0, 2, 3, 1), xr.grad.permute(0, 2, 3, 1)
    # outputs rounded to bf16: half an ulp of the largest magnitude (2^-9 relative) + fp32 accumulation-order noise
    assert float((yo.float() - want_y).abs().max()) <= 2.0 ** -8 * float(want_y.abs().max())
    assert float((xo.grad.float() - want_dx).abs().max()) <= 2.0 ** -8 * float(want_dx.abs().max())
    assert float((xo.grad.float() - want_dx).norm() / want_dx.norm()) < 3e-3
    # fp32 weight gradient: exact bf16 products, fp32 sums in a different order than cuDNN's
    assert float((wo.grad - wr.grad).abs().max()) <= 2e-4 * float(wr.grad.abs().max())
    assert float((wo.grad - wr.grad).norm() / wr.grad.norm()) < 2e-5
    # deterministic: the split-K partial products are summed in a fixed order
    wo.grad = None
    xo.grad = None
    y2 = train.conv2d_nhwc(xo, wo, stride)
    y2.backward(dy)
    first = wo.grad.clone()
    wo.grad = None
    train.conv2d_nhwc(xo, wo, stride).backward(dy)
    assert torch.equal(first, wo.grad)


def test_fine_tune_step_through_own_convolutions_matches_autograd():
    """One FinedTrainer2 forward / backward (reid/trainers.py:257-271) of the reference-style ResNet-50 (random init,
    train mode: BatchNorm on batch statistics) with every convolution swapped for the library's operators, against
    torch autograd (a) with bf16 rounding at the same points (tests/train_ref.py) and (b) in plain fp32.  A 53-layer
    network amplifies single-ulp differences (ReLU masks flip), so the whole-network bounds are looser than the
    per-operator ones above: loss to 2e-3 relative of (a), every weight gradient with cosine > 0.99 and the median
    relative error < 2e-2 against (a) -- the distance between (a) and (b), printed, is the scale to read them against."""
    import torch
    import train_ref
    from reid.loss import TripletLoss
    from reid.trainers import FinedTrainer2
    from ssg_b200 import synth, train
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    model = synth.build_model(num_split=2, seed=3).to(dev).train()
    P, K = 4, 4
    imgs, _ = synth.synth_images(P * K, seed=9, device=dev, per_identity=K)
    pids = [torch.arange(P, device=dev).repeat_interleave(K) for _ in range(3)]
    crit = [TripletLoss(0.5, K, True).to(dev), TripletLoss(0.5, K, True).to(dev)]
    trainer = FinedTrainer2(model, crit)

    def run():
        model.zero_grad()
        loss, _ = trainer._forward([imgs], pids, 0)
        loss.backward()
        return float(loss.detach()), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    fp32 = run()
    with train_ref.bf16_rounding_convs(model) as n_ref:
        ref = run()
    with train.own_convs(model) as swapped:
        got = run()
    assert swapped == n_ref >= 53
    assert abs(got[0] - ref[0]) <= 2e-3 * abs(ref[0]), (got[0], ref[0], fp32[0])
    rel, rel_ref, cos = [], [], []
    for name in ref[1]:
        if not name.endswith("weight") or ref[1][name].dim() != 4:
            continue
        g, r, w = got[1][name].flatten(), ref[1][name].flatten(), fp32[1][name].flatten()
        rel.append(float((g - r).norm() / r.norm()))
        rel_ref.append(float((r - w).norm() / w.norm()))
        cos.append(float(torch.dot(g, r) / (g.norm() * r.norm())))
    print("conv weight gradients: own vs bf16-rounded autograd: median rel %.2e max %.2e, min cosine %.5f; "
          "bf16-rounded vs fp32 autograd: median rel %.2e max %.2e; loss own %.6f / bf16 ref %.6f / fp32 %.6f"
          % (np.median(rel), max(rel), min(cos), np.median(rel_ref), max(rel_ref), got[0], ref[0], fp32[0]))
    assert len(rel) >= 53 and min(cos) > 0.99 and np.median(rel) < 2e-2
