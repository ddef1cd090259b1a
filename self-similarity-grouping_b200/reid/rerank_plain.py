"""Drop-in for reid/rerank_plain.py of the reference (both drivers carry the import commented out,
selftraining.py:29).

``re_ranking`` -- the plain kNN-set Jaccard variant (rerank_plain.py:125-178) -- runs on the GPU (ssg_rerank_plain).
``re_ranking_lh`` (rerank_plain.py:27-123) is ``reid.rerank.re_ranking`` with an un-squared, un-exponentiated source
term; it has no CUDA path: the reference's own function is used when the reference is importable.
"""
from ssg_b200.rerank import re_ranking_plain as re_ranking  # noqa: F401


def re_ranking_lh(input_feature_source, input_feature, k1=20, k2=6, lambda_value=0.2, MemorySave=False, Minibatch=2000):
    from ._reference import load_shadowed
    ref = load_shadowed("rerank_plain.py", "_reference_rerank_plain")
    if ref is None:
        raise NotImplementedError("re_ranking_lh has no CUDA path; put the reference on sys.path (or set "
                                  "SSG_REFERENCE_ROOT) to use its CPU implementation, or call reid.rerank.re_ranking")
    return ref.re_ranking_lh(input_feature_source, input_feature, k1, k2, lambda_value, MemorySave, Minibatch)
