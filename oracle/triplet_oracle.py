"""CPU restatement of the fine-tune loss of the SSG iteration.  TEST INFRASTRUCTURE ONLY.

Follows ``reid/loss/triplet.py:11-77`` (``TripletLoss.forward`` with ``w=None``; the curriculum branch at
``:34-48`` is dead code, ``if False``) and the loss aggregation of ``reid/trainers.py:257-271``
(``FinedTrainer2._forward``).  Pinned against the unmodified reference module executed on CPU:
``tests/golden/triplet_*.npz`` (made by ``oracle/make_goldens.py``) and, in the build container, live in
``tests/test_oracle_triplet.py``.

Arithmetic: the reference works in float32 (``pow/sum``, ``addmm_``, ``clamp(min=1e-12).sqrt()``); a float32 GEMM is
not reproducible bit for bit across BLAS back ends, so the parity bar for this row is a tolerance (1e-5 relative on
the loss and the gradient, stated in the tests), not bit equality.  The restatement accumulates in float64 and
rounds the distance matrix to float32, which is within that tolerance of any float32 GEMM.

The gradient is the analytic derivative of the same expression (what ``loss.backward()`` produces):
  d loss / d dist[a,b] = c[a,b];   d dist / d d2 = 1/(2 dist) where d2 >= 1e-12, else 0 (``clamp`` backward);
  d d2[a,b] / d x_a = 2 (x_a - x_b),  d d2[a,b] / d x_b = 2 (x_b - x_a).
Ties in ``neg_examples.min()`` (``triplet.py:55``): torch spreads the gradient evenly over tied minima; ties between
float32 distances of distinct samples do not occur on real features, and both the oracle and the CUDA path take the
first minimum (documented deviation, exercised nowhere on the hot path).
"""
import numpy as np


def pairwise_dist(x):
    """triplet.py:27-31: dist = sqrt(clamp(|xi|^2 + |xj|^2 - 2 xi.xj, 1e-12)), float32 result."""
    x64 = np.asarray(x, dtype=np.float64)
    sq = (x64 * x64).sum(1)
    d2 = (sq[:, None] + sq[None, :] - 2.0 * (x64 @ x64.T)).astype(np.float32)
    return np.sqrt(np.maximum(d2, np.float32(1e-12))), d2


def mine(dist, targets, K, use_semi=True):
    """triplet.py:33,49-61: lists of (anchor, positive) pairs and the chosen negative of each pair.

    semi (default, ``:49-56``): anchors are taken by POSITION (P = n // K groups of K consecutive rows); each anchor
    a = i*K+j is paired with the later rows of its group, a < p < (i+1)*K; the negative is the closest row whose
    label differs from the anchor's (``mask[a] == 0``).  Rows beyond P*K are never anchors.
    OHEM (``:57-60``): per row the farthest same-label row (itself included) and the closest other-label row.
    Raises ValueError where the reference raises (``min()`` of an empty tensor / ``cat`` of an empty list).
    """
    n = dist.shape[0]
    t = np.asarray(targets)
    mask = t[None, :] == t[:, None]
    an_idx = np.full(n, -1, dtype=np.int64)
    anchors, pos = [], []
    rows = range((n // K) * K) if use_semi else range(n)
    for a in rows:
        neg = np.nonzero(~mask[a])[0]
        if neg.size == 0:
            raise ValueError("anchor %d has no negative in the batch" % a)
        an_idx[a] = neg[np.argmin(dist[a, neg])]          # first minimum
        if use_semi:
            for p in range(a + 1, (a // K + 1) * K):
                anchors.append(a)
                pos.append(p)
        else:
            same = np.nonzero(mask[a])[0]
            anchors.append(a)
            pos.append(same[np.argmax(dist[a, same])])    # first maximum
    if not anchors:
        raise ValueError("no triplets (num_instances < 2)")
    anchors, pos = np.array(anchors), np.array(pos)
    return anchors, pos, an_idx[anchors]


def triplet_loss(x, targets, K, margin=0.0, use_semi=True, with_grad=False):
    """TripletLoss(margin, num_instances=K, use_semi).forward(x, targets, epoch) -> (loss, prec[, dloss/dx]).

    triplet.py:62-76: MarginRankingLoss(margin)(dist_an, dist_ap, y=1) = mean(max(0, dist_ap - dist_an + margin));
    prec = mean(dist_an > dist_ap).
    """
    x = np.asarray(x, dtype=np.float32)
    dist, d2 = pairwise_dist(x)
    a, p, m = mine(dist, targets, K, use_semi)
    ap, an = dist[a, p].astype(np.float64), dist[a, m].astype(np.float64)
    hinge = ap - an + float(margin)
    loss = float(np.maximum(hinge, 0.0).mean())
    prec = float((an > ap).mean())
    if not with_grad:
        return loss, prec
    n = x.shape[0]
    coef = np.zeros((n, n), dtype=np.float64)              # d loss / d dist
    act = hinge > 0
    np.add.at(coef, (a[act], p[act]), 1.0 / a.size)
    np.add.at(coef, (a[act], m[act]), -1.0 / a.size)
    coef = np.where(d2 >= np.float32(1e-12), coef / dist.astype(np.float64), 0.0)   # chain through sqrt(clamp)
    w = coef + coef.T
    x64 = x.astype(np.float64)
    grad = w.sum(1)[:, None] * x64 - w @ x64
    return loss, prec, grad.astype(np.float32)


def fined_trainer2_loss(x2, banks, pids, K, margin):
    """reid/trainers.py:257-271 FinedTrainer2._forward for a model without the DEC head (``len(outputs) == 2``):
    loss = triplet(x2, pids[0]) + sum_i triplet(banks[i], pids[i]); prec = the global (x2) precision."""
    loss, prec = triplet_loss(x2, pids[0], K, margin)
    for i, b in enumerate(banks):
        loss += triplet_loss(b, pids[i], K, margin)[0]
    return loss, prec


def synth_batch(P, K, d, seed, sep=0.2, extra=0, normalise=False):
    """P identities x K instances (consecutive, as RandomIdentitySampler delivers them) + `extra` unpaired rows.
    `sep` scales the identity centres against unit noise: small values give overlapping identities (active hinges)."""
    rng = np.random.RandomState(seed)
    centres = rng.randn(P + extra, d)
    lab = np.concatenate([np.repeat(np.arange(P), K), P + np.arange(extra)])
    x = sep * centres[lab] + rng.randn(lab.size, d)
    if normalise:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32), lab.astype(np.int64)
