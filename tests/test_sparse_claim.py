"""CPU: the arithmetic fact the sparse form of final_dist rests on (include/ssg_b200.h, "Sparse form of final_dist"),
checked on the oracle's float32-mode restatement of reid/rerank.py:27-127.

A pair whose expanded k-reciprocal rows share no column has Jaccard distance exactly 1 (rerank.py:108-115), hence
final_dist = fl32(1*fl32(1-lambda)) + fl32(v_i+v_m)*lambda >= fl32(1-lambda) because v >= 0 (rerank.py:38-40,122).  So
the rho-quantile of selftraining.py:289-293 can be taken over the entries below that bound alone, provided the slice
fits inside them, with M counted as all pairs minus the exact zeros."""
import numpy as np
import pytest

from oracle import ssg_oracle as O


@pytest.mark.parametrize("n,lam,rho", [(160, 0.1, 1.6e-2), (220, 0.3, 5e-3), (120, 0.0, 2e-2)])
def test_untouched_entries_sit_above_the_bound_and_eps_needs_only_the_rest(n, lam, rho):
    tgt, _ = O.synth_features(n, 64, 0, per_cluster=10, noise=0.3)
    src, _ = O.synth_features(n // 2, 64, 1, per_cluster=10, noise=0.4)
    st = {}
    _, final = O.re_ranking(src, tgt, lambda_value=lam, mode="f32", stages=st)
    bound = float(np.float32(1.0 - lam))
    untouched = (st["Vq"] @ st["Vq"].T) == 0                      # no common non-zero column
    assert untouched.any() and (~untouched).any()
    assert (st["J"][untouched] == 1).all()
    assert (st["vec"] >= 0).all()
    assert (final[untouched] >= bound).all()
    # eps from the entries below the bound only
    iu = np.triu_indices(n, 1)
    vals = final[iu]
    m_total = vals.size - int((vals == 0).sum())
    top = int(np.round(rho * m_total))
    low = np.sort(vals[(vals != 0) & (vals < bound)])
    assert 0 < top <= low.size                                     # certified
    np.testing.assert_allclose(low[:top].mean(), O.eps_estimate(final, rho), rtol=1e-13)
    # and DBSCAN's region queries at that eps never see an untouched entry
    eps = O.eps_estimate(final, rho)
    assert eps < bound and not (final[untouched] <= eps).any()
