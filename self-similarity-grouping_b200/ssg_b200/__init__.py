"""ssg_b200 — B200-native pseudo-label hot path of Self-Similarity Grouping (host-side Python).

Thin layer over libssg_b200.so (hand-written sm_100a CUDA behind the C ABI of include/ssg_b200.h).
PyTorch is used for device memory, streams and torch.distributed only.
"""
from . import _lib  # noqa: F401
from .rerank import RerankPlan, re_ranking, re_ranking_device, sqdist  # noqa: F401
from .cluster import ClusterPlan, DBSCAN, eps_estimate, dbscan_labels  # noqa: F401

from .embed import EmbedPlan, embed_images, extract_features  # noqa: F401
from .triplet import triplet_loss  # noqa: F401
from . import train  # noqa: F401  (own_convs, GraphedStep: the fine-tune step on the library's convolutions)
from .cycle import pseudo_label_cycle, compute_dist, generate_selflabel, generate_keep_mask  # noqa: F401

__version__ = "0.1.0"
