"""Progress-line statistics for the trainers: the latest value of a quantity and its sample-weighted running mean.
Attribute surface of the reference's meter (reid/utils/meters.py:4-23: ``val``, ``avg``, ``sum``, ``count``,
``reset()``, ``update(val, n)``), as the trainers and evaluators print them."""


class AverageMeter(object):
    __slots__ = ("val", "sum", "count")

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.sum, self.count = 0.0, 0.0, 0

    @property
    def avg(self):
        """Weighted mean of everything seen since reset() (0 before the first update)."""
        return self.sum / self.count if self.count else 0.0

    def update(self, val, n=1):
        self.val = val
        self.count += n
        self.sum += n * val
