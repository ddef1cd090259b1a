"""Fine-tune step of the SSG iteration: ``FinedTrainer2`` of reid/trainers.py:204-292, the
trainer ``iter_trainer`` builds (selftraining.py:239-253).  The model forward/backward is PyTorch autograd over the
``reid.models.ResNet`` module (cuDNN convolutions: library code, named as such in DESIGN.md); the per-bank triplet
losses and their gradients are this repo's CUDA kernels (reid.loss.TripletLoss).  One host synchronisation per
step (``loss.item()``, as in the reference) -- the losses themselves do not synchronise (``check=False`` is NOT
used: a batch without negatives must raise like the reference does)."""
import time

import torch

from .utils.meters import AverageMeter


class FinedTrainer2(object):
    def __init__(self, model, criterions, beta=0.5):
        super(FinedTrainer2, self).__init__()
        self.model = model
        self.criterions = criterions
        self.beta = beta

    def train(self, epoch, train_loader, optimizer, print_freq=10):
        self.model.train()
        batch_time, data_time = AverageMeter(), AverageMeter()
        losses, precisions = AverageMeter(), AverageMeter()
        end = time.time()
        for i, batch in enumerate(train_loader):
            data_time.update(time.time() - end)
            inputs, pids, _ = self._parse_data(batch)
            loss, prec = self._forward(inputs, pids, epoch)
            bs = inputs[0].size(0)
            losses.update(loss.item(), bs)
            precisions.update(float(prec), bs)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            batch_time.update(time.time() - end)
            end = time.time()
            if (i + 1) % print_freq == 0:
                print('Epoch: [{}][{}/{}]\t'
                      'Time {:.3f} ({:.3f})\t'
                      'Data {:.3f} ({:.3f})\t'
                      'Loss {:.3f} ({:.3f})\t'
                      'Prec {:.2%} ({:.2%})\t'
                      .format(epoch, i + 1, len(train_loader), batch_time.val, batch_time.avg,
                              data_time.val, data_time.avg, losses.val, losses.avg,
                              precisions.val, precisions.avg))

    def _parse_data(self, inputs):
        """trainers.py:250-255: (imgs, fnames, [labels per bank], weight) -> ([imgs], [labels on the GPU], weight)."""
        imgs, _, pids, w = inputs
        dev = next(self.model.parameters()).device
        return [imgs.to(dev)], [p.to(dev) for p in pids], torch.as_tensor(w).float().to(dev)

    def _forward(self, inputs, pids, epoch):
        """trainers.py:257-271 (models with the DEC head, ``len(outputs) == 3``, are outside the hot path)."""
        outputs = self.model(*inputs)
        if len(outputs) == 3:
            raise NotImplementedError("the DEC head (--dce-loss) is outside the pseudo-label hot path")
        loss, prec = self.criterions[1](outputs[1], pids[0], epoch)
        if isinstance(outputs[0], list):
            for i, bank in enumerate(outputs[0]):
                loss = loss + self.criterions[0](bank, pids[i], epoch)[0]
        else:
            loss = loss + self.criterions[0](outputs[0], pids[0], epoch)[0]
        return loss, prec


# The other trainers of the reference (Trainer, FinedTrainer, DistillTrainer, JointTrainer*: imported by the drivers,
# selftraining.py:19 / semitraining.py:19 / eug.py:4, built only on paths outside the pseudo-label cycle) are the
# reference's own classes when the reference is reachable.
from . import _reference  # noqa: E402

_ref = None
try:
    _ref = _reference.load_shadowed("trainers.py", "_reference_trainers")
except Exception as exc:  # a reference that does not import (missing third-party module): keep the hot path usable
    import warnings
    warnings.warn("reid.trainers: the reference's trainers could not be imported (%s)" % (exc,))
if _ref is not None:
    for _name in ("BaseTrainer", "Trainer", "DistillTrainer", "FinedTrainer", "JointTrainer", "JointTrainer2"):
        if hasattr(_ref, _name):
            globals()[_name] = getattr(_ref, _name)
