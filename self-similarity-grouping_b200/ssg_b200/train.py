"""Training-side convolutions of the fine-tune step.

The reference fine-tunes the ResNet-50 on the pseudo-labels with ``loss.backward()`` through torch autograd
(reid/trainers.py:204-271 FinedTrainer2.train / _forward; the convolutions of reid/models/resnet.py:52-70 run on cuDNN).
Here every 1x1 / 3x3 convolution of the trunk -- forward, data gradient and weight gradient -- and the 7x7 stem
(forward and weight gradient) run on the library's tcgen05 GEMM kernels (csrc/train.cu, include/ssg_b200.h
``ssg_op_conv*``) behind ``torch.autograd.Function``s; BatchNorm (batch statistics), ReLU, pooling and the optimiser
stay with torch.

Numerics: activations and gradients cross the kernels as bf16 (fp32 accumulation), weights are rounded to bf16 per
call from the fp32 master copy, the weight gradient is fp32.  Against fp32 autograd the relative error of a gradient
is a few 1e-3 per convolution (tests/test_gpu_train_ops.py).

``own_convs(model)`` swaps the forward of the eligible ``nn.Conv2d`` modules (a context manager / undo handle); the
model, its parameters and the optimiser are untouched, so the reference's trainer code runs as it is.
``own_convs(model, activations="bf16", cast_back=...)`` additionally keeps the tensors between the convolutions in bf16
channels-last (zero-copy in and out of the kernels), and ``GraphedStep`` captures the whole optimisation step in one
CUDA graph.  Measured on B200s at batch 64: 19.3-19.8 ms per FinedTrainer2 step on eager cuDNN autograd
(the reference's path); here 23-34 ms eager (launch / host bound, box dependent) and 18.1 ms as a graph.
"""
import contextlib

from . import _lib


def _f():
    import torch
    return torch


class _ConvNHWC(object):
    """Namespace for the autograd function (defined lazily: importing ssg_b200 must not import torch.autograd eagerly)."""
    fn = None
    zeros = {}


def _zero_bias(n, dev):
    """fp32 zeros [n] on ``dev`` (training-mode convolutions carry no bias; cached per device and length)."""
    key = (str(dev), int(n))
    z = _ConvNHWC.zeros.get(key)
    if z is None:
        z = _ConvNHWC.zeros[key] = _f().zeros(int(n), dtype=_f().float32, device=dev)
    return z


def _conv_fn():
    if _ConvNHWC.fn is not None:
        return _ConvNHWC.fn
    torch = _f()

    class ConvNHWC(torch.autograd.Function):
        """y [B,H/s,W/s,cout] bf16 = conv_kxk(x [B,H,W,cin] bf16, weight fp32 [cout,cin,k,k]), padding k//2."""

        @staticmethod
        def forward(ctx, x, weight, stride):
            lib = _lib.load()
            x = x.contiguous()
            weight = weight.contiguous()
            B, H, W, cin = x.shape
            cout, _, k, _ = weight.shape
            dev = x.device
            st = _lib.stream_ptr(dev)
            wp = torch.empty(cout * k * k * cin, dtype=torch.bfloat16, device=dev)
            _lib.check(lib.ssg_op_conv_pack_weight(weight.data_ptr(), cout, cin, k, 0, wp.data_ptr(), st))
            y = torch.empty((B, H // stride, W // stride, cout), dtype=torch.bfloat16, device=dev)
            zero = _zero_bias(cout, dev)
            scratch = torch.empty(x.numel() + 64, dtype=torch.bfloat16, device=dev) if stride == 2 else None
            _lib.check(lib.ssg_op_conv(x.data_ptr(), B, H, W, cin, k, stride, wp.data_ptr(), zero.data_ptr(), cout, None, 0,
                                       y.data_ptr(), scratch.data_ptr() if scratch is not None else None, st))
            ctx.save_for_backward(x, weight)
            ctx.stride = stride
            return y

        @staticmethod
        def backward(ctx, dy):
            lib = _lib.load()
            x, weight = ctx.saved_tensors
            B, H, W, cin = x.shape
            cout, _, k, _ = weight.shape
            dy = dy.contiguous()
            st = _lib.stream_ptr(x.device)
            dx = dw = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                _lib.check(lib.ssg_op_conv_dgrad(dy.data_ptr(), B, H, W, cout, k, ctx.stride, weight.data_ptr(), cin,
                                                 dx.data_ptr(), st))
            if ctx.needs_input_grad[1]:
                dw = torch.empty_like(weight)
                _lib.check(lib.ssg_op_conv_wgrad(x.data_ptr(), B, H, W, cin, dy.data_ptr(), cout, k, ctx.stride,
                                                 dw.data_ptr(), st))
            return dx, dw, None

    class StemNHWC(torch.autograd.Function):
        """y [n,128,64,64] bf16 = conv_7x7/2(images fp32 [n,3,256,128], weight fp32 [64,3,7,7]), padding 3: the im2col
        (ssg_op_stem_im2col, K padded 147 -> 192) followed by a 1x1 operator; no gradient reaches the images."""

        @staticmethod
        def forward(ctx, images, weight):
            lib = _lib.load()
            images = images.contiguous().float()
            n = images.shape[0]
            dev = images.device
            st = _lib.stream_ptr(dev)
            col = torch.empty((n * 8192, 192), dtype=torch.bfloat16, device=dev)
            _lib.check(lib.ssg_op_stem_im2col(images.data_ptr(), n, 0, col.data_ptr(), st))
            w192 = torch.zeros((64, 192), dtype=torch.float32, device=dev)
            w192[:, :147] = weight.permute(0, 2, 3, 1).reshape(64, 147)          # (kh, kw, ci) order
            wp = w192.to(torch.bfloat16).contiguous()
            y = torch.empty((n, 128, 64, 64), dtype=torch.bfloat16, device=dev)
            zero = _zero_bias(64, dev)
            _lib.check(lib.ssg_op_conv(col.data_ptr(), n, 128, 64, 192, 1, 1, wp.data_ptr(), zero.data_ptr(), 64, None, 0,
                                       y.data_ptr(), None, st))
            ctx.save_for_backward(col)
            return y

        @staticmethod
        def backward(ctx, dy):
            lib = _lib.load()
            (col,) = ctx.saved_tensors
            n = col.shape[0] // 8192
            dy = dy.contiguous()
            dw = None
            if ctx.needs_input_grad[1]:
                dw192 = torch.empty((64, 192, 1, 1), dtype=torch.float32, device=col.device)
                _lib.check(lib.ssg_op_conv_wgrad(col.data_ptr(), n, 128, 64, 192, dy.data_ptr(), 64, 1, 1, dw192.data_ptr(),
                                                 _lib.stream_ptr(col.device)))
                dw = dw192.reshape(64, 192)[:, :147].reshape(64, 7, 7, 3).permute(0, 3, 1, 2).contiguous()
            return None, dw

    _ConvNHWC.fn = (ConvNHWC, StemNHWC)
    return _ConvNHWC.fn


def conv2d_nhwc(x, weight, stride=1):
    """NHWC bf16 in / out (see ``ConvNHWC``); differentiable in ``x`` and ``weight``."""
    _lib.require_cuda()
    return _conv_fn()[0].apply(x, weight, int(stride))


def conv2d(x, weight, stride=1, out_dtype=None):
    """torch-layout adapter: x NCHW (fp32 or bf16, any memory format) -> NCHW view of the NHWC result, through
    ``conv2d_nhwc``.  A bf16 channels-last input is consumed as it is (no copy); ``out_dtype`` defaults to x.dtype -- a bf16
    result is the zero-copy view of the kernel's output."""
    torch = _f()
    xh = x.to(dtype=torch.bfloat16, memory_format=torch.channels_last).permute(0, 2, 3, 1)
    y = conv2d_nhwc(xh, weight, stride).permute(0, 3, 1, 2)
    want = x.dtype if out_dtype is None else out_dtype
    return y if want == torch.bfloat16 else y.to(want)


def stem_conv2d(images, weight, out_dtype=None):
    """The 7x7/2 stem on fp32 NCHW images of 256 x 128 pixels -> NCHW [n,64,128,64] (fp32 unless ``out_dtype``)."""
    torch = _f()
    _lib.require_cuda()
    if tuple(images.shape[1:]) != (3, 256, 128) or tuple(weight.shape) != (64, 3, 7, 7):
        raise ValueError("stem_conv2d: images [n,3,256,128] and weight [64,3,7,7] expected, got %s / %s"
                         % (tuple(images.shape), tuple(weight.shape)))
    y = _conv_fn()[1].apply(images, weight).permute(0, 3, 1, 2)
    want = torch.float32 if out_dtype is None else out_dtype
    return y if want == torch.bfloat16 else y.to(want)


def _eligible(conv, stem_too):
    k = conv.kernel_size
    if conv.groups != 1 or conv.bias is not None or conv.dilation != (1, 1) or k[0] != k[1]:
        return None
    if k[0] in (1, 3) and conv.padding == (k[0] // 2, k[0] // 2) and conv.stride in ((1, 1), (2, 2)) \
            and conv.in_channels % 64 == 0 and conv.out_channels % 64 == 0:
        return "conv"
    if stem_too and k[0] == 7 and conv.padding == (3, 3) and conv.stride == (2, 2) and conv.in_channels == 3 \
            and conv.out_channels == 64:
        return "stem"
    return None


@contextlib.contextmanager
def own_convs(model, stem=True, activations=None, cast_back=None):
    """Inside the block every eligible ``nn.Conv2d`` of ``model`` (1x1 / 3x3, stride 1 / 2, channels multiples of 64, no
    bias; and the 7x7/2 stem on 256 x 128 images) computes forward AND backward on the library's kernels.  Yields the
    number of swapped modules.

    activations=None   : every swapped convolution returns its input's dtype -- an fp32 model keeps fp32 activations, each
                         layer pays an fp32 NCHW <-> bf16 NHWC conversion, and the only roundings are those of the
                         convolution operands / results (what tests/train_ref.py reproduces).
    activations="bf16" : the swapped convolutions return the kernels' bf16 channels-last output as a zero-copy view, so the
                         BatchNorm / ReLU / pooling / residual adds between them run on bf16 channels-last tensors and the
                         next convolution consumes them without a copy (mixed-precision training as under autocast:
                         parameters, BatchNorm statistics and weight gradients stay fp32).  ``cast_back``: a module whose
                         output is cast back to fp32 (e.g. ``model.base.layer4`` in front of fp32 heads)."""
    import types
    torch = _f()
    if activations not in (None, "bf16"):
        raise ValueError("own_convs: activations must be None or 'bf16'")
    out_dtype = torch.bfloat16 if activations == "bf16" else None
    swapped = []
    for mod in model.modules():
        if not isinstance(mod, torch.nn.Conv2d):
            continue
        kind = _eligible(mod, stem)
        if kind is None:
            continue
        if kind == "conv":
            def fwd(self, x):
                return conv2d(x, self.weight, self.stride[0], out_dtype)
        else:
            def fwd(self, x):
                if tuple(x.shape[1:]) != (3, 256, 128):
                    y = torch.nn.Conv2d.forward(self, x)
                    return y if out_dtype is None else y.to(out_dtype)
                return stem_conv2d(x, self.weight, out_dtype)
        mod.forward = types.MethodType(fwd, mod)
        swapped.append(mod)
    hook = None
    if cast_back is not None and out_dtype is not None:
        hook = cast_back.register_forward_hook(lambda m, i, o: o.float())
    try:
        yield len(swapped)
    finally:
        if hook is not None:
            hook.remove()
        for mod in swapped:
            del mod.forward


class GraphedStep(object):
    """One optimisation step -- forward, loss, backward, ``optimizer.step()`` -- captured ONCE in a CUDA graph and replayed.

    At the fine-tune batch size (64 images) the step is bound by launch count and host overhead, not by the GPU (DESIGN.md
    §3.5b): the graph removes both.  ``step_fn(*inputs) -> loss`` (a 0-dim tensor) must be free of host synchronisation
    (``TripletLoss.check = False``: the no-negative check of reid/loss/triplet.py:55 reads a flag on the host) and is traced
    with whatever convolution implementation is active at construction time (``with own_convs(model): GraphedStep(...)``);
    replays do not need the context manager.  ``__call__(*inputs)`` copies the inputs into the captured buffers, replays
    and returns the captured loss tensor (valid until the next call).  SGD / momentum are capture-safe as they are; the
    library's own launches go to torch's current stream, which is the capturing stream, and its scratch memory comes
    from the stream-ordered allocator, which CUDA graphs record as allocation nodes.  Drop every reference to losses /
    outputs of earlier EAGER steps of the same model first: their autograd graphs keep AccumulateGrad nodes bound to the
    default stream alive, and a capture that reaches one is invalidated (cudaErrorStreamCaptureInvalidated)."""

    def __init__(self, step_fn, optimizer, example_inputs, warmup=3):
        torch = _f()
        _lib.require_cuda()
        self.static_inputs = [t.clone() for t in example_inputs]
        self.optimizer = optimizer
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):                      # lazy state (momentum buffers, kernel attributes)
                optimizer.zero_grad(set_to_none=True)
                step_fn(*self.static_inputs).backward()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(*self.static_inputs)
            self.loss.backward()
            optimizer.step()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
