"""Fine-tune loss of the SSG iteration on the GPU: reid/loss/triplet.py:11-77.

``triplet_loss(x, targets, num_instances, margin, use_semi)`` is a ``torch.autograd.Function`` over the two C-ABI
entry points ``ssg_triplet_forward`` / ``ssg_triplet_backward`` (csrc/triplet.cu): pairwise distances, negative
mining, hinge and the gradient run in three kernels instead of the reference's O(P*K^2) Python loop of ``.view(1)``
cats.  There is no CPU path.
"""
import torch

from . import _lib


class _TripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, targets, num_instances, margin, use_semi, check):
        _lib.require_cuda()
        if not (x.is_cuda and x.dim() == 2):
            raise ValueError("ssg_b200.triplet_loss: inputs must be a CUDA tensor [n, d]")
        xc = x.detach().contiguous().float()
        if xc.data_ptr() % 16:
            xc = xc.clone()
        tg = targets.detach().to(device=x.device, dtype=torch.int64).contiguous()
        n, d = xc.shape
        if tg.numel() != n:
            raise ValueError("ssg_b200.triplet_loss: %d targets for %d rows" % (tg.numel(), n))
        dist = torch.empty((n, n), dtype=torch.float32, device=x.device)
        coef = torch.empty((n, n), dtype=torch.float32, device=x.device)
        out = torch.empty((2,), dtype=torch.float32, device=x.device)
        status = torch.empty((2,), dtype=torch.int32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ssg_triplet_forward(xc.data_ptr(), tg.data_ptr(), n, d, int(num_instances),
                                                       float(margin), int(bool(use_semi)), dist.data_ptr(),
                                                       coef.data_ptr(), out.data_ptr(), status.data_ptr(),
                                                       _lib.stream_ptr()))
        if check:
            st = status.tolist()       # one sync; the reference raises at neg_examples.min() (triplet.py:55)
            if st[0]:
                raise RuntimeError("TripletLoss: anchor %d has no sample with a different label in the batch" % st[1])
        ctx.save_for_backward(xc, coef)
        ctx.in_dtype = x.dtype
        ctx.mark_non_differentiable(out, dist)
        return out[0].clone(), out, dist

    @staticmethod
    def backward(ctx, g_loss, _g_out, _g_dist):
        xc, coef = ctx.saved_tensors
        n, d = xc.shape
        g = g_loss.detach().to(torch.float32).contiguous()
        gx = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.load().ssg_triplet_backward(xc.data_ptr(), n, d, coef.data_ptr(), g.data_ptr(),
                                                        gx.data_ptr(), _lib.stream_ptr()))
        return gx.to(ctx.in_dtype), None, None, None, None, None


def triplet_loss(inputs, targets, num_instances, margin=0.0, use_semi=True, check=True, return_dist=False):
    """-> (loss, prec) as 0-dim CUDA tensors (loss differentiable w.r.t. ``inputs``)."""
    loss, out, dist = _TripletFn.apply(inputs, targets, num_instances, margin, use_semi, check)
    if return_dist:
        return loss, out[1], dist.detach()
    return loss, out[1]
