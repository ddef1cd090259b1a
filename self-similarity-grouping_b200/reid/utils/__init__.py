"""Host utilities with the reference's names (reid/utils/__init__.py:6-23: ``to_numpy`` / ``to_torch``); the
sub-modules this drop-in does not carry (``data``, ``logging``, ``serialization``, ``osutils``) are the reference's."""
import numpy as np
import torch

from .. import _reference

_reference.extend_path(__path__, "utils")


def to_numpy(tensor):
    if torch.is_tensor(tensor):
        return tensor.cpu().numpy()
    if not isinstance(tensor, (np.ndarray, np.generic)):
        raise ValueError("Cannot convert {} to numpy array".format(type(tensor)))
    return tensor


def to_torch(ndarray):
    if isinstance(ndarray, (np.ndarray, np.generic)):
        return torch.from_numpy(np.asarray(ndarray))
    if not torch.is_tensor(ndarray):
        raise ValueError("Cannot convert {} to torch tensor".format(type(ndarray)))
    return ndarray
