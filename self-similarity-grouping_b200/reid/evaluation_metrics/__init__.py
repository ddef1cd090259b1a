"""CMC / mAP on the GPU (reid/evaluation_metrics/ranking.py of the reference); ``accuracy`` (classification.py) is
the reference's own when it is reachable."""
from .. import _reference
from .ranking import cmc, mean_ap  # noqa: F401

__all__ = ['cmc', 'mean_ap']
if _reference.extend_path(__path__, "evaluation_metrics"):
    try:
        from .classification import accuracy  # noqa: F401  (reference file)
        __all__.append('accuracy')
    except ImportError:
        pass
