"""``reid.cluster.DBSCAN`` — the GPU DBSCAN with sklearn's constructor / fit_predict surface
(call sites: selftraining.py:295,303,306)."""
from ssg_b200.cluster import DBSCAN, eps_estimate, dbscan_labels  # noqa: F401
