"""CPU: the training-side convolution operators (csrc/train.cu: ssg_op_conv_pack_weight / _dgrad / _wgrad, SURVEY.md §8 row
f1) EXECUTED on the host under the functional tcgen05 / TMA emulation of tests/cpu_cuda, against torch's float64
autograd formulas (torch.nn.grad.conv2d_input / conv2d_weight) on the same bf16-rounded operands -- a floating-point
kernel, so the reference is a plain torch one and the tolerance is written out: the data gradient is rounded to bf16 on
output (relative 2^-9), the weight gradient is an fp32 sum of exact bf16 products."""
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_cuda"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")

from test_cpu_emulated_tensor_kernels import from_bf16, to_bf16  # noqa: E402


@pytest.fixture(scope="module")
def tc_lib():
    import build_emu
    from ssg_b200 import _lib as L
    lib = ctypes.CDLL(build_emu.build_tc())
    for name, (res, args) in L.PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    lib.ssg_last_error.restype = ctypes.c_char_p
    return lib


def _case(B, H, W, cin, cout, k, stride, seed):
    rng = np.random.RandomState(seed)
    x = from_bf16(to_bf16(rng.randn(B, H, W, cin)))
    w = (rng.randn(cout, cin, k, k) / np.sqrt(cin * k * k)).astype(np.float32)
    dy = from_bf16(to_bf16(rng.randn(B, H // stride, W // stride, cout)))
    return x, w, dy


def _torch_grads(x, w, dy, k, stride):
    import torch
    xt = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)
    wt = torch.from_numpy(from_bf16(to_bf16(w)).astype(np.float64).reshape(w.shape))
    dyt = torch.from_numpy(dy.astype(np.float64)).permute(0, 3, 1, 2)
    dx = torch.nn.grad.conv2d_input(xt.shape, wt, dyt, stride=stride, padding=k // 2)
    dw = torch.nn.grad.conv2d_weight(xt, wt.shape, dyt, stride=stride, padding=k // 2)
    y = torch.nn.functional.conv2d(xt, wt, stride=stride, padding=k // 2)
    return dx.permute(0, 2, 3, 1).numpy(), dw.numpy(), y.permute(0, 2, 3, 1).numpy()


CASES = [
    (2, 8, 16, 64, 64, 1, 1),         # 1x1: plain GEMMs either way
    (1, 16, 16, 128, 64, 3, 1),       # 3x3: implicit-GEMM data gradient with mirrored taps, cin != cout
    (1, 8, 32, 64, 64, 3, 1),         # 3x3 64 -> 64: the kernel-row-sharing kernel computes the data gradient
    (2, 16, 32, 128, 128, 3, 2),      # 3x3 stride 2: gradient spread onto the even pixels, then a stride-1 convolution
    (2, 16, 16, 256, 128, 1, 2),      # 1x1 stride 2 (downsample branch)
    (3, 8, 16, 192, 64, 1, 1),        # the stem as a 1x1 operator on its im2col (K = 192)
    (4, 16, 32, 64, 64, 1, 1),        # 2 048 pixels: three split-K batches on the emulator's three SMs
]


@pytest.mark.parametrize("late", [0, 1], ids=["eager", "late"])
@pytest.mark.parametrize("B,H,W,cin,cout,k,stride", CASES)
def test_dgrad_wgrad_and_packed_forward_against_torch_formulas(tc_lib, B, H, W, cin, cout, k, stride, late):
    tc_lib.ssg_emu_set_async(late)
    x, w, dy = _case(B, H, W, cin, cout, k, stride, seed=B * 100 + cin + cout + k + stride)
    want_dx, want_dw, want_y = _torch_grads(x, w, dy, k, stride)
    xb, dyb = to_bf16(x), to_bf16(dy)

    def ck(rc):
        assert rc == 0, tc_lib.ssg_last_error().decode()
    # forward through ssg_op_conv on the packed weights (zero bias, no ReLU)
    wp = np.zeros(cout * k * k * cin, np.uint16)
    ck(tc_lib.ssg_op_conv_pack_weight(w.ctypes.data, cout, cin, k, 0, wp.ctypes.data, None))
    assert np.array_equal(wp.reshape(cout, k, k, cin), to_bf16(w.transpose(0, 2, 3, 1)))
    if cin % 64 == 0:
        y = np.zeros((B, H // stride, W // stride, cout), np.uint16)
        zero = np.zeros(cout, np.float32)
        scratch = np.zeros(x.size + 64, np.uint16)
        ck(tc_lib.ssg_op_conv(xb.ctypes.data, B, H, W, cin, k, stride, wp.ctypes.data, zero.ctypes.data, cout, None, 0,
                              y.ctypes.data, scratch.ctypes.data, None))
        assert np.abs(from_bf16(y) - want_y).max() <= 2.0 ** -8 * np.abs(want_y).max() + 1e-6
    # data gradient
    if cin % 64 == 0:
        dx = np.zeros((B, H, W, cin), np.uint16)
        ck(tc_lib.ssg_op_conv_dgrad(dyb.ctypes.data, B, H, W, cout, k, stride, w.ctypes.data, cin, dx.ctypes.data, None))
        err = np.abs(from_bf16(dx) - want_dx).max()
        assert err <= 2.0 ** -8 * np.abs(want_dx).max() + 1e-6, err
        if stride == 2 and k == 1:
            assert not from_bf16(dx)[:, 1::2].any() and not from_bf16(dx)[:, :, 1::2].any()
    # weight gradient
    dw = np.full((cout, cin, k, k), np.nan, np.float32)
    ck(tc_lib.ssg_op_conv_wgrad(xb.ctypes.data, B, H, W, cin, dyb.ctypes.data, cout, k, stride, dw.ctypes.data, None))
    err = np.abs(dw - want_dw).max()
    assert err <= 2e-5 * np.abs(want_dw).max() + 1e-6, err


def test_wgrad_is_deterministic_and_rejects_bad_shapes(tc_lib):
    tc_lib.ssg_emu_set_async(0)
    x, w, dy = _case(2, 8, 16, 64, 64, 3, 1, seed=11)
    xb, dyb = to_bf16(x), to_bf16(dy)
    outs = []
    for _ in range(2):
        dw = np.zeros(w.shape, np.float32)
        assert tc_lib.ssg_op_conv_wgrad(xb.ctypes.data, 2, 8, 16, 64, dyb.ctypes.data, 64, 3, 1, dw.ctypes.data, None) == 0
        outs.append(dw)
    assert np.array_equal(outs[0], outs[1])
    assert tc_lib.ssg_op_conv_wgrad(xb.ctypes.data, 2, 8, 16, 64, dyb.ctypes.data, 64, 5, 1, outs[0].ctypes.data, None) != 0
    assert tc_lib.ssg_op_conv_dgrad(dyb.ctypes.data, 2, 8, 16, 64, 3, 1, w.ctypes.data, 48, xb.ctypes.data, None) != 0
    assert b"not supported" in tc_lib.ssg_last_error()


def test_split_k_batches_stacked_along_m_with_two_m_blocks_per_split(tmp_path):
    """cout = 256 (two 128-row blocks per split) with the split count forced to 2 and 5 (SSG_WGRAD_SPLITS is read once
    per process): the producer's per-split K offset (AOperand::ksplit_mblks), ragged last split, same result."""
    import subprocess
    script = (
        "import sys, ctypes, numpy as np\n"
        "sys.path[:0] = %r\n"
        "import build_emu\n"
        "from ssg_b200 import _lib as L\n"
        "from test_cpu_emulated_train_ops import _case, _torch_grads, to_bf16\n"
        "lib = ctypes.CDLL(build_emu.build_tc())\n"
        "lib.ssg_op_conv_wgrad.restype, lib.ssg_op_conv_wgrad.argtypes = L.PROTOTYPES['ssg_op_conv_wgrad']\n"
        "x, w, dy = _case(3, 16, 16, 64, 256, 3, 2, seed=5)\n"
        "_, want, _ = _torch_grads(x, w, dy, 3, 2)\n"
        "dw = np.zeros(w.shape, np.float32); xb, dyb = to_bf16(x), to_bf16(dy)\n"
        "assert lib.ssg_op_conv_wgrad(xb.ctypes.data, 3, 16, 16, 64, dyb.ctypes.data, 256, 3, 2, dw.ctypes.data, None) == 0\n"
        "print('OK' if np.abs(dw - want).max() <= 2e-5 * np.abs(want).max() else 'BAD', np.abs(dw - want).max())\n"
        % ([ROOT, os.path.join(ROOT, "self-similarity-grouping_b200"), os.path.join(ROOT, "tests", "cpu_cuda"),
            os.path.join(ROOT, "tests")],))
    for splits in ("2", "5"):
        r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=1200,
                           env=dict(os.environ, SSG_WGRAD_SPLITS=splits, SSG_EMU_ASYNC="late"))
        assert r.returncode == 0 and r.stdout.startswith("OK"), (splits, r.stdout + r.stderr)


def test_autograd_wrappers_through_a_small_network_against_torch_autograd():
    """ssg_b200.train.own_convs on a conv3x3 -> BN(batch statistics) -> ReLU -> conv1x1/2 stack: the swapped modules run
    the library's forward / dgrad / wgrad (emulated), torch differentiates the rest.  Against torch autograd with bf16
    rounding at the same points (tests/train_ref.py) the loss and every gradient agree to 1e-3 relative; against plain
    fp32 autograd to bf16 accuracy (a few 1e-2: ReLU masks flip under rounding)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "self-similarity-grouping_b200"))
    import emu_device
    import train_ref
    undo = emu_device.install(tc=True)
    try:
        from ssg_b200 import train
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(64, 64, 3, padding=1, bias=False), torch.nn.BatchNorm2d(64), torch.nn.ReLU(),
                                  torch.nn.Conv2d(64, 128, 1, stride=2, bias=False)).train()
        x = torch.randn(2, 64, 8, 16, requires_grad=True)
        tgt = torch.randn(2, 128, 4, 8)

        def run():
            x.grad = None
            loss, grads = train_ref.grads_of(net, x, lambda y: ((y - tgt) ** 2).mean())
            return loss, [x.grad.clone()] + grads
        fp32 = run()
        with train_ref.bf16_rounding_convs(net) as n_ref:
            ref = run()
        with train.own_convs(net) as swapped:
            assert swapped == 2 == n_ref
            got = run()
        assert "forward" not in net[0].__dict__                      # restored
        assert abs(got[0] - ref[0]) <= 1e-4 * abs(ref[0]) and abs(got[0] - fp32[0]) <= 5e-3 * abs(fp32[0])
        for g, r, w in zip(got[1], ref[1], fp32[1]):
            assert float((g - r).norm() / r.norm()) < 1e-3
            assert float((g - w).norm() / w.norm()) < 5e-2
    finally:
        undo()


def test_stem_forward_and_weight_gradient_through_the_im2col_operator():
    """The 7x7/2 stem as im2col (ssg_op_stem_im2col) + 1x1 operator, forward and weight gradient, on one 256 x 128 image."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "self-similarity-grouping_b200"))
    import emu_device
    import train_ref
    undo = emu_device.install(tc=True)
    try:
        from ssg_b200 import train
        torch.manual_seed(1)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)).train()
        img = torch.randn(1, 3, 256, 128)
        tgt = torch.randn(1, 64, 128, 64)
        loss_fn = lambda y: ((y - tgt) ** 2).mean()      # noqa: E731
        with train_ref.bf16_rounding_convs(net) as n_ref:
            ref = train_ref.grads_of(net, img, loss_fn)
        with train.own_convs(net) as swapped:
            assert swapped == 1 == n_ref
            got = train_ref.grads_of(net, img, loss_fn)
        assert abs(got[0] - ref[0]) <= 1e-4 * abs(ref[0])
        assert float((got[1][0] - ref[1][0]).norm() / ref[1][0].norm()) < 1e-3
    finally:
        undo()


@pytest.mark.parametrize("cin,mid,stride", [(64, 64, 2), (256, 64, 1)])
def test_bottleneck_block_through_own_convolutions(cin, mid, stride):
    """A torchvision Bottleneck in train mode (conv1x1 - BN - ReLU - conv3x3[/2] - BN - ReLU - conv1x1 - BN (+ 1x1[/2]
    downsample - BN) - add - ReLU): the block structure of reid/models/resnet.py:52-70, every convolution on the
    library's operators (emulated), against torch autograd with bf16 rounding at the same points: every gradient to 5e-3
    relative (two valid bf16 evaluations of such a block differ by ~1e-3: ReLU masks flip under single-ulp changes)."""
    import torch
    from torchvision.models.resnet import Bottleneck
    sys.path.insert(0, os.path.join(ROOT, "self-similarity-grouping_b200"))
    import emu_device
    import train_ref
    undo = emu_device.install(tc=True)
    try:
        from ssg_b200 import train
        torch.manual_seed(cin + stride)
        ds = None
        if stride != 1 or cin != 4 * mid:
            ds = torch.nn.Sequential(torch.nn.Conv2d(cin, 4 * mid, 1, stride=stride, bias=False), torch.nn.BatchNorm2d(4 * mid))
        net = Bottleneck(cin, mid, stride, ds).train()
        x = torch.randn(2, cin, 8, 16, requires_grad=True)
        with torch.no_grad():
            tgt = torch.randn(net(x).shape)

        def run():
            x.grad = None
            loss, grads = train_ref.grads_of(net, x, lambda y: (y * tgt).sum() / float(tgt.numel()) ** 0.5)
            return loss, [x.grad.clone()] + grads
        with train_ref.bf16_rounding_convs(net) as n_ref:
            ref = run()
        with train.own_convs(net) as swapped:
            got = run()
        assert swapped == n_ref == (4 if ds is not None else 3)
        worst = max(float((g - r).norm() / r.norm()) for g, r in zip(got[1], ref[1]))
        assert worst < 5e-3, worst
    finally:
        undo()


def test_bf16_activation_mode_keeps_the_tensors_between_the_convolutions_in_bf16():
    """own_convs(activations="bf16"): the swapped convolutions hand their bf16 channels-last output on as a view, BatchNorm /
    ReLU / the residual add run on bf16 tensors, the next convolution consumes them without a copy and ``cast_back``
    returns fp32 to the caller.  Against the fp32-activation mode the gradients agree to the rounding of the
    activations (a few 1e-2); parameters' gradients stay fp32."""
    import torch
    from torchvision.models.resnet import Bottleneck
    sys.path.insert(0, os.path.join(ROOT, "self-similarity-grouping_b200"))
    import emu_device
    import train_ref
    undo = emu_device.install(tc=True)
    try:
        from ssg_b200 import train
        torch.manual_seed(3)
        ds = torch.nn.Sequential(torch.nn.Conv2d(64, 256, 1, stride=2, bias=False), torch.nn.BatchNorm2d(256))
        net = Bottleneck(64, 64, 2, ds).train()
        x = torch.randn(2, 64, 8, 16)
        with torch.no_grad():
            tgt = torch.randn(net(x).shape)
        loss_fn = lambda y: (y.float() * tgt).sum() / float(tgt.numel()) ** 0.5      # noqa: E731
        with train.own_convs(net) as swapped:
            want = train_ref.grads_of(net, x, loss_fn)
        seen = []
        probe = net.bn2.register_forward_hook(lambda m, i, o: seen.append((i[0].dtype, o.dtype, i[0].is_contiguous(
            memory_format=torch.channels_last))))
        with train.own_convs(net, activations="bf16", cast_back=net) as swapped2:
            out = net(x)
            got = train_ref.grads_of(net, x, loss_fn)
        probe.remove()
        assert swapped == swapped2 == 4 and out.dtype == torch.float32
        assert seen and all(s == (torch.bfloat16, torch.bfloat16, True) for s in seen)
        assert abs(got[0] - want[0]) < 3e-2 * max(1.0, abs(want[0]))
        for g, w in zip(got[1], want[1]):
            assert g.dtype == torch.float32 and float((g - w).norm() / w.norm()) < 6e-2
        assert not net._forward_hooks                                     # the cast-back hook is gone
    finally:
        undo()
