"""Drop-in ``reid`` package for the pseudo-label hot path of SHI-Labs/Self-Similarity-Grouping.

Put ``self-similarity-grouping_b200/`` ahead of the reference on ``sys.path`` and the unmodified
``selftraining.py`` picks up these modules for everything on the hot path:

    from reid import models                                   -> reid/models (same factory, same module tree)
    from reid.evaluators import Evaluator, extract_features   -> CUDA trunk (libssg_b200)
    from sklearn.cluster import DBSCAN
    from reid.rerank import *                                 -> CUDA re_ranking, and a CUDA ``DBSCAN`` that
                                                                 shadows sklearn's in the driver's namespace

    from reid.loss import TripletLoss                         -> CUDA triplet loss (row f1)
    from reid.trainers import FinedTrainer2                   -> the fine-tune step over it

Only the hot path and its "next" rows live here; datasets, samplers, the other trainers/losses and
checkpoint I/O are the reference's own and out of scope.
"""
from . import _reference

# everything this package does not define (datasets, dist_metric, metric_learning, utils.data ...) resolves to the
# reference's own files when the reference is on sys.path behind us / at SSG_REFERENCE_ROOT
_reference.extend_path(__path__)

from . import evaluation_metrics  # noqa: F401,E402
from . import feature_extraction  # noqa: F401,E402
from . import models  # noqa: F401,E402
from . import evaluators  # noqa: F401,E402
from . import rerank  # noqa: F401,E402
from . import rerank_initial  # noqa: F401,E402
from . import cluster  # noqa: F401,E402
from . import eug  # noqa: F401,E402
from . import loss  # noqa: F401,E402
from . import trainers  # noqa: F401,E402

__version__ = '0.2.0'
