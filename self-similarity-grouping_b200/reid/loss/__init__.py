"""Losses of the fine-tune step (reid/loss/__init__.py:3-5 of the reference).  Only the loss the self-training
driver uses is provided (selftraining.py:149-150): TripletLoss, on the GPU."""
from .triplet import TripletLoss  # noqa: F401

__all__ = ['TripletLoss']
