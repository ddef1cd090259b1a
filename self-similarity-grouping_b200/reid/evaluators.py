"""Drop-in for reid/evaluators.py of the reference: extract_features (18-60), pairwise_distance (63-85),
fliplr (12-16), evaluate_all (88-133) and the Evaluator wrapper (183-192); CMC / mAP are scored on the GPU by
reid/evaluation_metrics/ranking.py."""
from collections import OrderedDict  # noqa: F401

import torch

from ssg_b200.embed import extract_features  # noqa: F401  (evaluators.py:18)


def fliplr(img):
    '''flip horizontal (evaluators.py:12-16)'''
    inv_idx = torch.arange(img.size(3) - 1, -1, -1).long()
    return img.index_select(3, inv_idx)


def pairwise_distance(features, query=None, gallery=None, metric=None):
    """evaluators.py:63-85.  Returns a CPU float32 tensor like the reference.

    The distances come from the library's own kernel (``ssg_sqdist``, float64 direct difference: the exact
    ``sum_k (x_k - y_k)^2``) instead of the reference's float32 ``|x|^2 + |y|^2 - 2 x.y`` GEMM, whose cancellation
    noise (~1e-6 at unit norms) is the only difference -- well inside the 1e-4 distance tolerance.  The all-pairs
    branch of the reference (evaluators.py:64-72) computes ``2 |x_i|^2 - 2 x_i.x_j``, which equals the squared
    distance only for equal norms; that expression is reproduced as
    ``d2(i,j) + |x_i|^2 - |x_j|^2``."""
    from ssg_b200 import _lib
    from ssg_b200.rerank import sqdist
    dev = _lib.require_cuda()
    if query is None and gallery is None:
        n = len(features)
        x = torch.cat(list(features.values())).view(n, -1)
        if metric is not None:
            x = metric.transform(x)
        x = x.to(dev, dtype=torch.float32).contiguous()
        sq = torch.pow(x, 2).sum(dim=1, keepdim=True)
        dist = sqdist(x, x, _lib.DIST_EXACT) + (sq - sq.t())
        return dist.cpu()
    x = torch.cat([features[f].unsqueeze(0) for f, _, _ in query], 0)
    y = torch.cat([features[f].unsqueeze(0) for f, _, _ in gallery], 0)
    m, n = x.size(0), y.size(0)
    x, y = x.view(m, -1), y.view(n, -1)
    if metric is not None:
        x, y = metric.transform(x), metric.transform(y)
    x, y = x.to(dev, dtype=torch.float32).contiguous(), y.to(dev, dtype=torch.float32).contiguous()
    return sqdist(x, y, _lib.DIST_EXACT).cpu()


def evaluate_all(distmat, query=None, gallery=None, query_ids=None, gallery_ids=None, query_cams=None,
                 gallery_cams=None, cmc_topk=(1, 5, 10)):
    """evaluators.py:88-133: mean AP + the 'market1501' CMC protocol, same prints, returns CMC top-1."""
    from .evaluation_metrics import cmc, mean_ap
    if query is not None and gallery is not None:
        query_ids = [pid for _, pid, _ in query]
        gallery_ids = [pid for _, pid, _ in gallery]
        query_cams = [cam for _, _, cam in query]
        gallery_cams = [cam for _, _, cam in gallery]
    else:
        assert (query_ids is not None and gallery_ids is not None
                and query_cams is not None and gallery_cams is not None)
    mAP = mean_ap(distmat, query_ids, gallery_ids, query_cams, gallery_cams)
    print('Mean AP: {:4.1%}'.format(mAP))
    cmc_configs = {'market1501': dict(separate_camera_set=False, single_gallery_shot=False, first_match_break=True)}
    cmc_scores = {name: cmc(distmat, query_ids, gallery_ids, query_cams, gallery_cams, **params)
                  for name, params in cmc_configs.items()}
    print('CMC Scores{:>12}'.format('market1501'))
    for k in cmc_topk:
        print('top-{:<4}{:12.1%}'.format(k, cmc_scores['market1501'][k - 1]))
    return cmc_scores['market1501'][0]


class Evaluator(object):
    def __init__(self, model, print_freq):
        super(Evaluator, self).__init__()
        self.model = model
        self.print_freq = print_freq

    def distmat(self, data_loader, query, gallery, metric=None):
        features, _ = extract_features(self.model, data_loader, print_freq=self.print_freq)
        return pairwise_distance(features, query, gallery, metric=metric)

    def evaluate(self, data_loader, query, gallery, metric=None):
        """evaluators.py:189-192."""
        return evaluate_all(self.distmat(data_loader, query, gallery, metric), query=query, gallery=gallery)
