"""Drop-in for reid/rerank_plain.py of the reference (both drivers carry the import commented out,
selftraining.py:29): same names, positional order and defaults, both functions on the GPU.

``re_ranking``    -- the plain kNN-set Jaccard variant (rerank_plain.py:125-178)              -> ssg_rerank_plain
``re_ranking_lh`` -- reid.rerank.re_ranking with a float64, un-squared source term (:27-123)  -> ssg_rerank_lh
"""
from ssg_b200.rerank import re_ranking_plain as re_ranking  # noqa: F401
from ssg_b200.rerank import re_ranking_lh  # noqa: F401
