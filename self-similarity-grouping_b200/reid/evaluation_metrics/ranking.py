"""Drop-in for reid/evaluation_metrics/ranking.py (cmc 18-79, mean_ap 82-115) on the GPU.

Same signatures and return values.  All queries are ranked at once with torch CUDA ops (sort / cumsum / scatter) instead
of the reference's per-query numpy loop + sklearn ``average_precision_score``; the average precision reproduces
sklearn's definition (thresholds at distinct scores: tied distances form one threshold).  Ties in the ranking itself are
ordered by gallery index (the reference's np.argsort leaves them unspecified).  ``single_gallery_shot=True`` (random
gallery sampling, unused by the drivers and broken under numpy 2 in the reference: ``np.bool``) is not provided.
"""
import numpy as np


def _prepare(distmat, query_ids, gallery_ids, query_cams, gallery_cams):
    import torch
    from ssg_b200 import _lib
    dev = _lib.require_cuda()
    d = distmat if hasattr(distmat, "is_cuda") else torch.as_tensor(np.asarray(distmat))
    d = d.to(dev)
    m, n = d.shape
    def ids(x, default):
        return torch.as_tensor(np.asarray(default if x is None else x)).to(dev).long()
    q = ids(query_ids, np.arange(m))
    g = ids(gallery_ids, np.arange(n))
    qc = ids(query_cams, np.zeros(m, dtype=np.int32))
    gc = ids(gallery_cams, np.ones(n, dtype=np.int32))
    return d, q, g, qc, gc


def _ranked(d, q, g, qc, gc, separate_camera_set=False):
    """Sort every query's gallery by distance with the invalid entries (same id AND same camera) pushed to the end."""
    import torch
    valid = (g[None, :] != q[:, None]) | (gc[None, :] != qc[:, None])
    if separate_camera_set:
        valid &= (gc[None, :] != qc[:, None])
    dd = torch.where(valid, d, torch.full_like(d, float("inf")))
    ds, idx = torch.sort(dd, dim=1, stable=True)
    match = (g[idx] == q[:, None]) & torch.isfinite(ds)
    nvalid = valid.sum(1)
    return ds, match, nvalid


def cmc(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None, topk=100,
        separate_camera_set=False, single_gallery_shot=False, first_match_break=False):
    import torch
    if single_gallery_shot:
        raise NotImplementedError("single_gallery_shot CMC is not provided (unused by the drivers)")
    d, q, g, qc, gc = _prepare(distmat, query_ids, gallery_ids, query_cams, gallery_cams)
    ds, match, nvalid = _ranked(d, q, g, qc, gc, separate_camera_set)
    m, n = ds.shape
    pos = torch.arange(n, device=ds.device)[None, :]
    in_valid = pos < nvalid[:, None]
    nmatch = match.sum(1)
    has = nmatch > 0
    if int(has.sum()) == 0:
        raise RuntimeError("No valid query")
    # k - j of the reference = number of valid non-matching entries ranked before the j-th match
    nonmatch_before = torch.cumsum((in_valid & ~match).long(), 1)
    ret = torch.zeros(topk, dtype=torch.float64, device=ds.device)
    if first_match_break:
        first = torch.argmax(match.long(), dim=1)                   # position of the first match
        slot = nonmatch_before.gather(1, first[:, None]).squeeze(1)
        ok = has & (slot < topk)
        ret.scatter_add_(0, slot[ok], torch.ones(int(ok.sum()), dtype=torch.float64, device=ds.device))
    else:
        delta = (1.0 / nmatch.clamp(min=1).double())[:, None].expand(m, n)
        sel = match & (nonmatch_before < topk) & has[:, None]
        ret.scatter_add_(0, nonmatch_before[sel], delta[sel])
    return (ret.cumsum(0) / float(int(has.sum()))).cpu().numpy()


def mean_ap(distmat, query_ids=None, gallery_ids=None, query_cams=None, gallery_cams=None):
    import torch
    d, q, g, qc, gc = _prepare(distmat, query_ids, gallery_ids, query_cams, gallery_cams)
    ds, match, nvalid = _ranked(d, q, g, qc, gc)
    m, n = ds.shape
    pos = torch.arange(n, device=ds.device)[None, :]
    in_valid = pos < nvalid[:, None]
    tp = torch.cumsum(match.double(), 1)
    total = tp[:, -1]
    has = total > 0
    if int(has.sum()) == 0:
        raise RuntimeError("No valid query")
    # sklearn.average_precision_score: one threshold per distinct score -> evaluate at the END of each tie group
    nxt = torch.cat([ds[:, 1:], torch.full((m, 1), float("inf"), device=ds.device, dtype=ds.dtype)], 1)
    group_end = in_valid & ((nxt != ds) | (pos == (nvalid[:, None] - 1)))
    prec = tp / (pos + 1).double()
    tp_end = torch.where(group_end, tp, torch.zeros_like(tp))
    prev_end = torch.cummax(tp_end, 1).values
    prev_end = torch.cat([torch.zeros((m, 1), dtype=tp.dtype, device=tp.device), prev_end[:, :-1]], 1)
    ap = (torch.where(group_end, prec * (tp - prev_end), torch.zeros_like(tp))).sum(1) / total.clamp(min=1)
    return float(ap[has].mean().item())
