"""Test helper: torch autograd with bf16 rounding at exactly the points where the library's training-side convolutions
round (operands of every convolution in the forward pass, its output, and the gradients that cross it in the backward
pass), fp32 arithmetic otherwise -- the "torch autograd in bf16" reference of the f1 parity tests.  What is left between
this and ssg_b200.train.own_convs is accumulation order (1 ulp of bf16 here and there)."""
import contextlib
import types


def _round_fn():
    import torch

    class Round(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return t.to(torch.bfloat16).float()

        @staticmethod
        def backward(ctx, g):
            return g.to(torch.bfloat16).float()

    class RoundFwd(torch.autograd.Function):       # weights: rounded on the way in, fp32 gradient on the way out
        @staticmethod
        def forward(ctx, t):
            return t.to(torch.bfloat16).float()

        @staticmethod
        def backward(ctx, g):
            return g
    return Round, RoundFwd


@contextlib.contextmanager
def bf16_rounding_convs(model, stem=True):
    import torch
    Round, RoundFwd = _round_fn()
    swapped = []
    for mod in model.modules():
        if not isinstance(mod, torch.nn.Conv2d) or mod.bias is not None:
            continue
        k = mod.kernel_size[0]
        if not (k in (1, 3) and mod.in_channels % 64 == 0 and mod.out_channels % 64 == 0) and not (stem and k == 7):
            continue

        def fwd(self, x):
            xin = Round.apply(x) if x.requires_grad else x.to(torch.bfloat16).float()
            y = torch.nn.functional.conv2d(xin, RoundFwd.apply(self.weight), None, self.stride, self.padding)
            return Round.apply(y)
        mod.forward = types.MethodType(fwd, mod)
        swapped.append(mod)
    try:
        yield len(swapped)
    finally:
        for mod in swapped:
            del mod.forward


def grads_of(model, inputs, loss_fn):
    """(loss, [gradient of every parameter]) of one forward / backward pass."""
    model.zero_grad()
    loss = loss_fn(model(inputs))
    loss.backward()
    return float(loss.detach()), [None if p.grad is None else p.grad.detach().clone() for p in model.parameters()]
