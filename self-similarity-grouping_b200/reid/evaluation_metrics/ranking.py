"""Drop-in for reid/evaluation_metrics/ranking.py (cmc 18-79, mean_ap 82-115) on the GPU.

Same signatures and return values.  Both metrics come from ONE launch of the library's own kernel
(``ssg_rank_metrics``, csrc/metrics.cu) which never sorts the q x g matrix: for every match of every query it counts the
valid entries ranked before it -- the reference's ``k - j`` for CMC -- and sums sklearn's ``average_precision_score``
(thresholds at distinct scores: tied distances form one threshold).  Ties in the ranking itself are ordered by gallery
index (the reference's np.argsort leaves them unspecified).  The host only histograms the per-match counts and averages
the per-query AP (numpy, fixed order).  ``single_gallery_shot=True`` (random gallery sampling, unused by the drivers and
broken under numpy 2 in the reference: ``np.bool``) is not provided.
"""
import numpy as np

MAX_MATCHES = 1024          # SSG_RANK_MAX_MATCHES (include/ssg_b200.h)


def _prepare(distmat, query_ids, gallery_ids, query_cams, gallery_cams):
    import torch
    from ssg_b200 import _lib
    dev = _lib.require_cuda()
    d = distmat if hasattr(distmat, "is_cuda") else torch.as_tensor(np.asarray(distmat))
    d = d.to(dev)
    if d.dtype not in (torch.float32, torch.float64):
        d = d.float()
    d = d.contiguous()
    m, n = d.shape

    def ids(x, default):
        return torch.as_tensor(np.asarray(default if x is None else x)).to(dev).long().contiguous()
    q = ids(query_ids, np.arange(m))
    g = ids(gallery_ids, np.arange(n))
    qc = ids(query_cams, np.zeros(m, dtype=np.int32))
    gc = ids(gallery_cams, np.ones(n, dtype=np.int32))
    if q.numel() != m or qc.numel() != m or g.numel() != n or gc.numel() != n:
        raise ValueError("ids / cams do not match the %d x %d distance matrix" % (m, n))
    return d, q, g, qc, gc


def rank_metrics(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None,
                 separate_camera_set=False):
    """(ap [m] float64, nmatch [m] int32, slots [m, MAX_MATCHES] int32) as numpy arrays -- see include/ssg_b200.h
    ``ssg_rank_metrics``.  ``slots[i, :nmatch[i]]`` are the reference's ``k - j`` of the matches of query i."""
    import torch
    from ssg_b200 import _lib
    d, q, g, qc, gc = _prepare(distmat, query_ids, gallery_ids, query_cams, gallery_cams)
    m, n = d.shape
    dev = d.device
    ap = torch.empty(m, dtype=torch.float64, device=dev)
    nmatch = torch.empty(m, dtype=torch.int32, device=dev)
    slots = torch.empty((m, MAX_MATCHES), dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().ssg_rank_metrics(d.data_ptr(), _lib.F64 if d.dtype == torch.float64 else _lib.F32, m, n,
                                            q.data_ptr(), g.data_ptr(), qc.data_ptr(), gc.data_ptr(),
                                            1 if separate_camera_set else 0, ap.data_ptr(), nmatch.data_ptr(),
                                            slots.data_ptr(), flags.data_ptr(), _lib.stream_ptr(dev)))
    nm = nmatch.cpu().numpy()
    if int(flags.cpu()[0]) or (nm < 0).any():
        raise RuntimeError("a query has more than %d matching gallery entries (SSG_RANK_MAX_MATCHES)" % MAX_MATCHES)
    width = max(int(nm.max()), 1)
    return ap.cpu().numpy(), nm, slots[:, :width].cpu().numpy()


def cmc(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None, topk=100,
        separate_camera_set=False, single_gallery_shot=False, first_match_break=False):
    if single_gallery_shot:
        raise NotImplementedError("single_gallery_shot CMC is not provided (unused by the drivers)")
    _, nm, slots = rank_metrics(distmat, query_ids, gallery_ids, query_cams, gallery_cams, separate_camera_set)
    has = nm > 0
    if not has.any():
        raise RuntimeError("No valid query")
    live = np.arange(slots.shape[1])[None, :] < nm[:, None]
    ret = np.zeros(topk, dtype=np.float64)
    if first_match_break:
        # the first match in ranking order is the one with the fewest valid non-matches before it (ranking.py:67-70)
        first = np.where(live, slots, np.iinfo(np.int32).max).min(axis=1)
        first = first[has & (first < topk)]
        np.add.at(ret, first, 1.0)
    else:
        # every match adds 1 / #matches at its k - j (ranking.py:71-75)
        delta = np.broadcast_to((1.0 / np.maximum(nm, 1))[:, None], slots.shape)
        sel = live & (slots < topk)
        np.add.at(ret, slots[sel], delta[sel])
    return ret.cumsum() / float(has.sum())


def mean_ap(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None):
    ap, nm, _ = rank_metrics(distmat, query_ids, gallery_ids, query_cams, gallery_cams, False)
    has = nm > 0
    if not has.any():
        raise RuntimeError("No valid query")
    return float(np.mean(ap[has]))
