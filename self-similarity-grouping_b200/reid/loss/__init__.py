"""Losses (reid/loss/__init__.py:3-5 of the reference).  ``TripletLoss`` -- the loss the self-training driver builds
(selftraining.py:149-150) -- runs on this repo's CUDA kernels; ``FocalLoss`` is a plain-torch restatement so that the
driver's ``from reid.loss import TripletLoss,FocalLoss`` keeps working; OIM / WeightCE are the reference's own files
when the reference is reachable."""
from .. import _reference
from .triplet import TripletLoss, FocalLoss  # noqa: F401

__all__ = ['TripletLoss', 'FocalLoss']
if _reference.extend_path(__path__, "loss"):
    try:
        from .oim import oim, OIM, OIMLoss  # noqa: F401  (reference files)
        from .weight_cross_entropy import WeightCE  # noqa: F401
        __all__ += ['oim', 'OIM', 'OIMLoss', 'WeightCE']
    except ImportError:
        pass
