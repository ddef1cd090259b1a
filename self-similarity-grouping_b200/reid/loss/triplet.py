"""Drop-in for reid/loss/triplet.py:11-77 of the reference: ``TripletLoss(margin, num_instances, use_semi)`` with
``forward(inputs, targets, epoch, w=None) -> (loss, prec)``.  Distances, mining, hinge and gradient are CUDA kernels
(ssg_b200.triplet -> csrc/triplet.cu); ``epoch`` is accepted and unused, as in the reference (its curriculum branch
is ``if False``, triplet.py:34)."""
from torch import nn

from ssg_b200.triplet import triplet_loss


class TripletLoss(nn.Module):
    def __init__(self, margin=0, num_instances=0, use_semi=True):
        super(TripletLoss, self).__init__()
        self.margin = margin
        self.use_semi = use_semi
        self.K = num_instances

    def forward(self, inputs, targets, epoch=0, w=None):
        if w is not None:
            # triplet.py:68-71: every negative distance against every positive one (an O(T^2) variant no driver
            # reaches: FinedTrainer2 drops `w`, trainers.py:250-258)
            raise NotImplementedError("TripletLoss(w=...) is not on the self-training path")
        return triplet_loss(inputs, targets, self.K, self.margin, self.use_semi)
