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
        self.feature_dim = d

    def embed(self, model, images, num_split, out=None):
        import torch
        x = images.numpy()
        banks = []
        for W in self.W:
            f = x @ W
            banks.append(f / np.linalg.norm(f, axis=1, keepdims=True))
        res = torch.from_numpy(np.stack(banks, 0).astype(np.float32))
        if out is not None:                    # this rank's slot of the gather buffer (ssg_b200.dist.embed_and_gather)
            out.copy_(res)
            return out
        return res

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

    # ---- sparse form of final_dist (sparse=True): CSR over the pairs whose expanded rows share a column
    def finish_sparse(self, plan, tgt, k1, k2, lambda_value):
        import torch
        st = {}
        src = plan.src_by_tgt[tgt.data_ptr()]
        _, f = O.re_ranking(src.numpy(), tgt.numpy(), k1, k2, lambda_value, mode="f32", stages=st)
        touched = (st["Vq"] @ st["Vq"].T) != 0
        rowptr = np.concatenate([[0], np.cumsum(touched.sum(1))]).astype(np.int32)
        rows, cols = np.nonzero(touched)                               # row-major: ascending columns inside a row
        return (torch.from_numpy(rowptr), torch.from_numpy(cols.astype(np.int32)), torch.from_numpy(f[rows, cols]),
                float(np.float32(1.0 - lambda_value)))

    sparse_calls = 0

    def eps_sparse(self, n, rowptr, col, val, bound, rho):
        rp, c, v = rowptr.numpy(), col.numpy(), val.numpy()
        row = np.repeat(np.arange(n), np.diff(rp))
        up = v[c > row]
        m_total = n * (n - 1) // 2 - int((up == 0).sum())
        top = int(np.round(rho * m_total))
        low = np.sort(up[(up != 0) & (up < bound)])
        if top > low.size:
            return float("nan"), False
        return (float(low[:top].mean()) if top else float("nan")), True

    def dbscan_sparse(self, n, rowptr, col, val, eps, min_samples):
        import torch
        type(self).sparse_calls += 1
        rp, c, v = rowptr.numpy(), col.numpy(), val.numpy()
        dense = np.full((n, n), np.inf)
        dense[np.repeat(np.arange(n), np.diff(rp)), c] = v
        return torch.from_numpy(O.dbscan_dfs(dense, eps, min_samples))

    # ---- row-sharded finish (shard_finish=True): rows of final_dist + the numpy stand-in of the sharded primitives
    def finish_rows(self, plan, tgt, k1, k2, lambda_value, row0, rows, final_rows):
        import torch
        full = torch.empty((tgt.shape[0], tgt.shape[0]), dtype=torch.float64)
        self.finish(plan, tgt, k1, k2, lambda_value, full)
        final_rows.copy_(full[row0:row0 + rows])

    def new_final_rows(self, rows, n):
        import torch
        return torch.empty((rows, n), dtype=torch.float64)

    def cluster_plan(self, n, max_neighbors=0):
        from shard_fake import FakeClusterPlan
        if getattr(self, "_cplan", None) is None or self._cplan.max_neighbors < max_neighbors:
            self._cplan = FakeClusterPlan(n, max_neighbors or self.nbr_cap)
        return self._cplan

    nbr_cap = 0          # tests set a tiny capacity to drive the collective capacity retry

    def with_capacity_retry(self, n, fn):
        cap = 0
        while True:
            try:
                return fn(self.cluster_plan(n, cap))
            except OverflowError:
                cap = max(self._cplan.max_neighbors * 8, 64)

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


def _worker(rank, world, init_file, out_dir, shard_finish=False, nbr_cap=0, sparse=False, rho=None):
    import sys
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))      # shard_fake (spawned interpreter)
    from ssg_b200 import dist as sd
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        t, s = _images()
        tl, th = sd.shard_bounds(N_T, world, rank)
        sl, sh = sd.shard_bounds(N_S, world, rank)
        be = OracleBackend(BANKS, D_IMG, D)
        be.nbr_cap = nbr_cap
        labels, eps, keep = sd.sharded_pseudo_label_cycle(
            None, torch.from_numpy(t[tl:th]), torch.from_numpy(s[sl:sh]), N_T, N_S, num_split=BANKS - 1,
            lambda_value=LAM, rho=RHO if rho is None else rho, backend=be, comm=sd.Comm(), shard_finish=shard_finish,
            sparse=sparse)
        if sparse:
            np.save(os.path.join(out_dir, "sparse_calls%d.npy" % rank), np.array(OracleBackend.sparse_calls))
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


def _single_process_reference():
    import torch
    t, s = _images()
    be = OracleBackend(BANKS, D_IMG, D)
    tf = be.embed(None, torch.from_numpy(t), BANKS - 1).numpy()
    sf = be.embed(None, torch.from_numpy(s), BANKS - 1).numpy()
    return [O.re_ranking(sf[b], tf[b], lambda_value=LAM, mode="f32")[1] for b in range(BANKS)]


@pytest.mark.parametrize("world,nbr_cap", [(2, 0), (3, 0), (4, 0), (5, 0), (3, 16)])
def test_row_sharded_finish_matches_single_process(world, nbr_cap):
    """shard_finish=True: every rank holds only its rows of final_dist; eps by the distributed radix select
    (histogram all-reduce, list all-gather), DBSCAN through the gathered counts and the all-reduced neighbour CSR.
    nbr_cap=16 starts from a neighbour list that is too small, so all ranks go through the capacity retry together."""
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "init")
        mp.spawn(_worker, args=(world, init_file, tmp, True, nbr_cap), nprocs=world, join=True)
        outs = [np.load(os.path.join(tmp, "rank%d.npz" % r)) for r in range(world)]
    finals = _single_process_reference()
    want_eps = [O.eps_estimate(f, RHO) for f in finals]
    for o in outs:
        # eps: same radix-selected order statistic, float64 sums taken in another order
        np.testing.assert_allclose(o["eps"], want_eps, rtol=1e-13, atol=0)
        assert np.array_equal(o["eps"], outs[0]["eps"])                      # bit-identical across the ranks
        want_labels = [O.dbscan_dfs(f, e, 4) for f, e in zip(finals, o["eps"])]
        assert np.array_equal(o["labels"], np.stack(want_labels))
        assert np.array_equal(o["keep"], O.keep_mask(want_labels))
    assert max(l.max() for l in want_labels) >= 1


def test_row_sharded_eps_fake_matches_oracle_incl_massive_ties():
    """The numpy stand-in itself (single rank and 3 simulated ranks without collectives): 3-pass path, the 6-pass
    fallback, zeros dropped, empty slice -> NaN."""
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_fake import FakeClusterPlan, shard_lo
    rng = np.random.RandomState(3)
    n = 57
    a = rng.rand(n, n)
    a[a < 0.1] = 0.0                          # exact zeros are dropped (np.nonzero, selftraining.py:290)
    d = np.round((a + a.T) / 2, 2)            # symmetric, heavy ties
    for rho in (0.05, 0.5, 1e-6):
        want = O.eps_estimate(d, rho)
        for world in (1, 3):
            for exact in (False, True):
                plans = [FakeClusterPlan(n) for _ in range(world)]
                blocks = [torch.from_numpy(d[shard_lo(n, world, r):shard_lo(n, world, r + 1)].copy())
                          for r in range(world)]
                for p in plans:
                    p.eps_shard_begin()
                for npass in range(6 if exact else 2):
                    for r, p in enumerate(plans):
                        p.eps_shard_hist(blocks[r], n, world, r, npass)
                    tot = sum(p.hist for p in plans)
                    for p in plans:
                        p.hist.copy_(tot)
                        p.eps_shard_pick(npass, rho)
                lists = []
                for r, p in enumerate(plans):
                    c = p.eps_shard_gather(blocks[r], n, world, r, exact)
                    lists.append(p.list[:c].clone() if not exact else None)
                part = sum(p.partial for p in plans)
                got = []
                for p in plans:
                    p.partial.copy_(part)     # rows are disjoint: the sum is the all-gather
                    if not exact:
                        merged = torch.cat(lists)
                        p.list[:len(merged)] = merged
                        p.state[5] = len(merged)
                    got.append(p.eps_shard_finish(n, exact)[0])
                assert all(g == got[0] or (np.isnan(g) and np.isnan(got[0])) for g in got)
                if np.isnan(want):
                    assert np.isnan(got[0])
                else:
                    np.testing.assert_allclose(got[0], want, rtol=1e-13)


@pytest.mark.parametrize("world,rho,expect_sparse", [(2, RHO, True), (4, RHO, True), (3, 0.9, False)])
def test_sparse_owner_finish_matches_single_process(world, rho, expect_sparse):
    """sparse=True: the bank owners finish on the CSR form of final_dist; a rho whose slice cannot be certified
    (0.9: most pairs) must fall back to the dense matrix on the owner and still give the reference's labels."""
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "init")
        mp.spawn(_worker, args=(world, init_file, tmp, False, 0, True, rho), nprocs=world, join=True)
        outs = [np.load(os.path.join(tmp, "rank%d.npz" % r)) for r in range(world)]
        calls = sum(int(np.load(os.path.join(tmp, "sparse_calls%d.npy" % r))) for r in range(world))
    finals = _single_process_reference()
    want_eps = [O.eps_estimate(f, rho) for f in finals]
    assert calls == (BANKS if expect_sparse else 0)
    for o in outs:
        np.testing.assert_allclose(o["eps"], want_eps, rtol=1e-13, atol=0)
        want_labels = [O.dbscan_dfs(f, e, 4) for f, e in zip(finals, o["eps"])]
        assert np.array_equal(o["labels"], np.stack(want_labels))
        assert np.array_equal(o["keep"], O.keep_mask(want_labels))
