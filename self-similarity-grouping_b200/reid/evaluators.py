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
    """evaluators.py:63-85.  Returns a CPU float32 tensor like the reference; the GEMM runs on the GPU."""
    from ssg_b200 import _lib
    dev = _lib.require_cuda()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        if query is None and gallery is None:
            n = len(features)
            x = torch.cat(list(features.values())).view(n, -1)
            if metric is not None:
                x = metric.transform(x)
            x = x.to(dev)
            dist = torch.pow(x, 2).sum(dim=1, keepdim=True) * 2
            dist = dist.expand(n, n) - 2 * torch.mm(x, x.t())
            return dist.cpu()
        x = torch.cat([features[f].unsqueeze(0) for f, _, _ in query], 0)
        y = torch.cat([features[f].unsqueeze(0) for f, _, _ in gallery], 0)
        m, n = x.size(0), y.size(0)
        x, y = x.view(m, -1), y.view(n, -1)
        if metric is not None:
            x, y = metric.transform(x), metric.transform(y)
        x, y = x.to(dev), y.to(dev)
        dist = torch.pow(x, 2).sum(dim=1, keepdim=True).expand(m, n) + \
            torch.pow(y, 2).sum(dim=1, keepdim=True).expand(n, m).t()
        dist = torch.addmm(dist, x, y.t(), beta=1, alpha=-2)
        return dist.cpu()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


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
