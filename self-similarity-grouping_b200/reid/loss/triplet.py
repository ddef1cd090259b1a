"""Drop-in for reid/loss/triplet.py:11-77 of the reference: ``TripletLoss(margin, num_instances, use_semi)`` with
``forward(inputs, targets, epoch, w=None) -> (loss, prec)``.  Distances, mining, hinge and gradient are CUDA kernels
(ssg_b200.triplet -> csrc/triplet.cu); ``epoch`` is accepted and unused, as in the reference (its curriculum branch
is ``if False``, triplet.py:34)."""
import torch
from torch import nn
from torch.nn import functional as F

from ssg_b200.triplet import triplet_loss


class TripletLoss(nn.Module):
    def __init__(self, margin=0, num_instances=0, use_semi=True):
        super(TripletLoss, self).__init__()
        self.margin = margin
        self.use_semi = use_semi
        self.K = num_instances
        # True: raise like the reference (triplet.py:55) when an anchor has no other-label row -- one host read of a
        # device flag per call; False: no host synchronisation (the loss is NaN in that case), e.g. under CUDA-graph capture
        self.check = True

    def forward(self, inputs, targets, epoch=0, w=None):
        if w is not None:
            # triplet.py:68-71: every negative distance against every positive one (an O(T^2) variant no driver
            # reaches: FinedTrainer2 drops `w`, trainers.py:250-258)
            raise NotImplementedError("TripletLoss(w=...) is not on the self-training path")
        return triplet_loss(inputs, targets, self.K, self.margin, self.use_semi, check=getattr(self, "check", True))


class FocalLoss(nn.Module):
    """reid/loss/triplet.py:79-107 (imported, never constructed, by the drivers): -alpha_t (1 - p_t)^gamma log p_t.
    The reference's constructor tests ``isinstance(alpha, (float, int, long))`` and therefore raises NameError on
    Python 3; this restatement accepts the same arguments and works."""

    def __init__(self, gamma=2.0, alpha=None, size_average=True):
        super(FocalLoss, self).__init__()
        self.gamma = gamma
        if isinstance(alpha, (float, int)):
            alpha = torch.tensor([alpha, 1 - alpha])
        elif isinstance(alpha, (list, tuple)):
            alpha = torch.tensor(alpha)
        self.alpha = alpha
        self.size_average = size_average

    def forward(self, input, target, epoch=0):
        if input.dim() > 2:                      # N,C,H,W -> N*H*W,C
            input = input.flatten(2).transpose(1, 2).reshape(-1, input.size(1))
        target = target.view(-1, 1)
        logpt = F.log_softmax(input, dim=1).gather(1, target).view(-1)
        pt = logpt.detach().exp()
        if self.alpha is not None:
            logpt = logpt * self.alpha.to(input).gather(0, target.view(-1))
        loss = -((1 - pt) ** self.gamma) * logpt
        return loss.mean() if self.size_average else loss.sum()
