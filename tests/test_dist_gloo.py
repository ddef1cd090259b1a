"""CPU, world_size 2 over gloo: the multi-GPU sharding logic of ssg_b200.dist (shard bounds, padded all-gather of
uneven row blocks, row-block distance stage + table gather, bank ownership, label broadcast) with the compute
replaced by the oracle, checked against the single-process oracle cycle."""
import os
import tempfile

import numpy as np
import pytest

from oracle import ssg_oracle as O


def test_shard_bounds_cover_and_balance():
    from ssg_b200.dist import shard_bounds, max_shard
    for n in (0, 1, 7, 16702, 36411, 126441):
        for w in (1, 2, 3, 4, 8):
            cuts = [shard_bounds(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == (max_shard(n, w) if n else 0)


class FakePlan(object):
    def __init__(self, n):
        import torch
        self.rowmin = torch.zeros(n)
        self.rowmax = torch.zeros(n)
        self.rank = torch.full((n, 32), -1, dtype=torch.int32)
        self.rank_val = torch.zeros(n, 32)
        self.src = None
        self.src_by_tgt = {}


class OracleBackend(object):
    """Same interface as ssg_b200.dist.CudaBackend, computing with the numpy oracle on CPU tensors."""

    def __init__(self, n_banks, d_img, d):
        rng = np.random.RandomState(5)
        self.W = [rng.randn(d_img, d).astype(np.float32) for _ in range(n_banks)]

    def embed(self, model, images, num_split):
        import torch
        x = images.numpy()
        banks = []
        for W in self.W:
            f = x @ W
            banks.append(f / np.linalg.norm(f, axis=1, keepdims=True))
        return torch.from_numpy(np.stack(banks, 0).astype(np.float32))

    def plan(self, n, ns, d):
        return FakePlan(n)

    def distance_rows(self, plan, src, tgt, k1, row0, rows):
        import torch
        from scipy.spatial.distance import cdist
        plan.src = src
        plan.src_by_tgt[tgt.data_ptr()] = src
        t, s = tgt.numpy(), src.numpy()
        blk = slice(row0, row0 + rows)
        st = np.power(cdist(t[blk], s), 2).astype(np.float32)
        od = np.power(cdist(t[blk], t).astype(np.float32), 2).astype(np.float32)
        plan.rowmin[blk] = torch.from_numpy(st.min(1))
        mx = od.max(1)
        plan.rowmax[blk] = torch.from_numpy(mx)
        odn = od / mx[:, None]
        r = np.argsort(odn, kind="stable")[:, :k1 + 1]
        plan.rank[blk, :k1 + 1] = torch.from_numpy(r.astype(np.int32))
        plan.rank_val[blk, :k1 + 1] = torch.from_numpy(np.take_along_axis(odn, r, 1))

    def tables(self, plan, n):
        return [plan.rowmin, plan.rowmax, plan.rank, plan.rank_val]

    def finish(self, plan, tgt, k1, k2, lambda_value, final):
        import torch
        st = {}
        src = plan.src_by_tgt[tgt.data_ptr()]          # phase B runs after all banks' distance stages
        _, f = O.re_ranking(src.numpy(), tgt.numpy(), k1, k2, lambda_value, mode="f32", stages=st)
        # the gathered tables must be exactly what a single process computes
        assert np.array_equal(plan.rank[:, :k1 + 1].numpy(), st["rank"][:, :k1 + 1])
        assert np.array_equal(plan.rowmax.numpy(), st["od"].max(0))
        final.copy_(torch.from_numpy(f))

    def new_final(self, n):
        import torch
        return torch.empty((n, n), dtype=torch.float64)

    def eps(self, final, rho):
        return float(O.eps_estimate(final.numpy(), rho))

    def dbscan(self, final, eps, min_samples):
        import torch
        return torch.from_numpy(O.dbscan_dfs(final.numpy(), eps, min_samples))


N_T, N_S, D_IMG, D, BANKS, RHO, LAM = 61, 45, 24, 32, 3, 0.05, 0.1


def _images():
    rng = np.random.RandomState(0)
    cent = rng.randn(6, D_IMG)
    t = cent[rng.randint(0, 6, N_T)] + 0.3 * rng.randn(N_T, D_IMG)
    s = cent[rng.randint(0, 6, N_S)] + 0.4 * rng.randn(N_S, D_IMG)
    return t.astype(np.float32), s.astype(np.float32)


def _worker(rank, world, init_file, out_dir):
    import torch
    import torch.distributed as dist
    from ssg_b200 import dist as sd
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        t, s = _images()
        tl, th = sd.shard_bounds(N_T, world, rank)
        sl, sh = sd.shard_bounds(N_S, world, rank)
        labels, eps, keep = sd.sharded_pseudo_label_cycle(
            None, torch.from_numpy(t[tl:th]), torch.from_numpy(s[sl:sh]), N_T, N_S, num_split=BANKS - 1,
            lambda_value=LAM, rho=RHO, backend=OracleBackend(BANKS, D_IMG, D), comm=sd.Comm())
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), labels=np.stack(labels), eps=np.array(eps), keep=keep)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])      # 4 ranks > 3 banks: some ranks own no bank (the 8-GPU regime)
def test_sharded_cycle_matches_single_process(world):
    import torch
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "init")
        mp.spawn(_worker, args=(world, init_file, tmp), nprocs=world, join=True)
        outs = [np.load(os.path.join(tmp, "rank%d.npz" % r)) for r in range(world)]
    # single-process reference
    t, s = _images()
    be = OracleBackend(BANKS, D_IMG, D)
    tf = be.embed(None, torch.from_numpy(t), BANKS - 1).numpy()
    sf = be.embed(None, torch.from_numpy(s), BANKS - 1).numpy()
    want_labels, want_eps = [], []
    for b in range(BANKS):
        _, f = O.re_ranking(sf[b], tf[b], lambda_value=LAM, mode="f32")
        e = O.eps_estimate(f, RHO)
        want_eps.append(e)
        want_labels.append(O.dbscan_dfs(f, e, 4))
    for o in outs:
        assert np.array_equal(o["labels"], np.stack(want_labels))
        np.testing.assert_allclose(o["eps"], want_eps, rtol=0, atol=1e-15)
        assert np.array_equal(o["keep"], O.keep_mask(want_labels))
    assert max(l.max() for l in want_labels) >= 1          # the case is not degenerate
