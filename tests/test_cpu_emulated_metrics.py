"""CPU: the retrieval-metrics kernel (csrc/metrics.cu, ssg_rank_metrics) EXECUTED on the host under tests/cpu_cuda and
driven through the real drop-in ``reid.evaluation_metrics`` wrappers, against the numpy + sklearn restatement of
reid/evaluation_metrics/ranking.py:18-115 (which tests/test_oracle_vs_reference.py pins to the reference module)."""
import os
import shutil
import sys

import numpy as np
import pytest

from oracle import ssg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "self-similarity-grouping_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def metrics():
    import emu_device
    undo = emu_device.install()
    from reid.evaluation_metrics import ranking
    yield ranking
    undo()


def _case(m, n, nid, ncam, seed, quant=None, dtype=np.float32):
    rng = np.random.RandomState(seed)
    qid, gid = rng.randint(0, nid, m), rng.randint(0, nid, n)
    qcam, gcam = rng.randint(0, ncam, m), rng.randint(0, ncam, n)
    d = (rng.rand(m, n) + 0.5 * (qid[:, None] != gid[None, :])).astype(dtype)
    if quant:
        d = (np.round(d * quant) / quant).astype(dtype)
    return d, qid, gid, qcam, gcam


@pytest.mark.parametrize("m,n,nid,ncam,quant,dtype", [(40, 300, 25, 3, None, np.float32), (25, 700, 12, 2, None, np.float64),
                                                      (30, 260, 10, 3, 20, np.float32), (9, 33, 4, 1, 4, np.float64),
                                                      (5, 1, 2, 2, None, np.float32)])
def test_cmc_and_mean_ap_kernel_against_the_restatement(metrics, m, n, nid, ncam, quant, dtype):
    """Distinct and heavily tied distances (quantised: ties inside and across match / non-match entries), one camera
    (every same-id entry filtered), a 1-entry gallery; both CMC flavours and separate_camera_set."""
    d, qid, gid, qcam, gcam = _case(m, n, nid, ncam, seed=m + n, quant=quant, dtype=dtype)
    try:
        want_map = O.mean_ap(d, qid, gid, qcam, gcam)
    except RuntimeError:
        with pytest.raises(RuntimeError, match="No valid query"):
            metrics.mean_ap(d, qid, gid, qcam, gcam)
        return
    assert abs(metrics.mean_ap(d, qid, gid, qcam, gcam) - want_map) < 1e-12
    for sep in (False, True):
        for fmb in (False, True):
            try:
                want = O.cmc(d, qid, gid, qcam, gcam, topk=20, first_match_break=fmb, separate_camera_set=sep)
            except RuntimeError:
                with pytest.raises(RuntimeError, match="No valid query"):
                    metrics.cmc(d, qid, gid, qcam, gcam, topk=20, first_match_break=fmb, separate_camera_set=sep)
                continue
            got = metrics.cmc(d, qid, gid, qcam, gcam, topk=20, first_match_break=fmb, separate_camera_set=sep)
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


def test_per_match_counts_and_per_query_ap(metrics):
    """The raw kernel outputs: slots = the reference's k - j per match (stable order), AP per query = sklearn's."""
    from sklearn.metrics import average_precision_score
    d, qid, gid, qcam, gcam = _case(20, 150, 6, 3, seed=3, quant=10)
    ap, nm, slots = metrics.rank_metrics(d, qid, gid, qcam, gcam)
    for i in range(d.shape[0]):
        valid = (gid != qid[i]) | (gcam != qcam[i])
        match = valid & (gid == qid[i])
        assert nm[i] == match.sum()
        if nm[i] == 0:
            continue
        order = np.argsort(d[i], kind="stable")
        order = order[valid[order]]
        pos = {g: k for k, g in enumerate(order)}
        k_minus_j = sorted(pos[g] - j for j, g in enumerate(sorted(np.nonzero(match)[0], key=lambda g: pos[g])))
        assert sorted(slots[i, :nm[i]].tolist()) == k_minus_j
        assert abs(ap[i] - average_precision_score(match[valid], -d[i][valid])) < 1e-12


def test_too_many_matches_is_refused(metrics):
    d = np.random.RandomState(0).rand(2, metrics.MAX_MATCHES + 5).astype(np.float32)
    with pytest.raises(RuntimeError, match="more than"):
        metrics.mean_ap(d, np.zeros(2, int), np.zeros(d.shape[1], int), np.zeros(2, int), np.ones(d.shape[1], int))
