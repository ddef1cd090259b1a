"""Multi-GPU pseudo-label cycle: one process per GPU (torch.distributed, NCCL over NVLink), SURVEY.md §8e.

Sharding of ONE cycle over `world` ranks:
  1. embed   : rank r embeds the contiguous image shard [lo_r, hi_r) of the target and of the source set;
  2. exchange: all-gather of the feature banks ([banks, N, 2048] float32 on every rank) — the one real exchange step;
  3. re-rank : phase A, for every bank, rank r computes the distance stages (GEMM, candidate selection, exact
               re-scoring) for its row block [lo_r, hi_r) only (ssg_rerank_distance_rows) and the small per-row tables
               (row min / max, 21 rank columns; ~4.4 MB per bank) are all-gathered; phase B, the owner ranks
               (bank % world) run the cheap remaining stages (k-reciprocal encoding ... final distance), eps and DBSCAN
               of their banks in parallel;
  4. labels  : broadcast from the owners (N int64 per bank).

`sparse=True` (opt-in, SSG_SPARSE_FINISH=1): the owners finish their banks on the sparse form of final_dist (DESIGN.md
3.6) -- no N x N matrix, so phase B costs about a millisecond per bank and the cycle scales with the row-sharded
distance stage.

`shard_finish=True` (opt-in, SSG_SHARD_FINISH=1) replaces phase B and step 4 by a row-sharded finish: every rank runs the
sparse stages of a bank (cheap, they need all rows' tables anyway) and then only ITS rows of final_dist
(ssg_rerank_finish_rows); eps is a distributed radix select (histogram all-reduce, list all-gather), DBSCAN a local row
scan with an all-gather of the neighbour counts and a sum all-reduce that completes the neighbour CSR, after which every
rank labels all rows identically.  Nothing N x N lives on one GPU, and the N^2-sized stages scale with the ranks
instead of with min(banks, world).

All collectives go through a small `Comm` wrapper, and all compute through a `backend` object, so that the sharding
logic itself is exercised on CPU with the gloo backend (tests/test_dist_gloo.py) against a fake backend.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous balanced shard [lo, hi) of n rows for `rank`: the first n % world ranks get one extra row."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_shard(n, world):
    return (int(n) + int(world) - 1) // int(world)


class Comm(object):
    """torch.distributed wrapper with row-block all-gather for uneven shards (pad to the largest shard)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0

    def all_gather_rows(self, local, n_total, dim=0):
        """local: this rank's rows [lo, hi) along `dim` -> tensor with all n_total rows along `dim` on every rank."""
        import torch
        if self.world == 1:
            return local
        m = max_shard(n_total, self.world)
        loc = local.movedim(dim, 0).contiguous()
        pad = torch.zeros((m,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
        pad[: loc.shape[0]] = loc
        out = torch.empty((self.world * m,) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
        self.dist.all_gather_into_tensor(out, pad, group=self.group)
        parts = []
        for r in range(self.world):
            lo, hi = shard_bounds(n_total, self.world, r)
            parts.append(out[r * m: r * m + (hi - lo)])
        return torch.cat(parts, 0).movedim(0, dim).contiguous()

    # ---- feature banks: [banks, rows, d] gathered along the row axis without staging copies.  The gather buffer is
    # [banks, world * m, d] (m = largest shard): rank r owns rows [r*m, r*m + cnt_r) of every bank, the embedding writes
    # its features straight there (gather_slot), one in-place all_gather_into_tensor per bank runs asynchronously
    # (NCCL's own stream: it overlaps whatever the caller launches next, e.g. the embedding of the other image set) and
    # gather_banks_finish returns the [banks, n_total, d] result -- a zero-copy view when the shards are even, one
    # compaction pass otherwise.  (all_gather_rows above pads, gathers, concatenates and transposes: four extra passes
    # over the data; SURVEY.md 2.2 asks for kernels that write straight into the gather buffer.)
    def gather_buffer(self, banks, n_total, d, dtype, device):
        import torch
        m = max_shard(n_total, self.world)
        return torch.empty((banks, self.world * m, d), dtype=dtype, device=device)

    def gather_slot(self, buf, n_total):
        """The rows of `buf` this rank must fill: buf[:, rank*m : rank*m + (hi - lo)]."""
        m = buf.shape[1] // self.world
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        return buf[:, self.rank * m: self.rank * m + (hi - lo)]

    def gather_banks_begin(self, buf):
        """Start the in-place all-gathers of a filled gather buffer; returns the work handles."""
        if self.world == 1:
            return []
        m = buf.shape[1] // self.world
        works = []
        for b in range(buf.shape[0]):
            works.append(self.dist.all_gather_into_tensor(buf[b], buf[b, self.rank * m: (self.rank + 1) * m],
                                                          group=self.group, async_op=True))
        return works

    def gather_banks_finish(self, buf, works, n_total):
        import torch
        for w in works:
            w.wait()
        m = buf.shape[1] // self.world
        if self.world * m == n_total:
            return buf                                       # even shards: the buffer IS the result
        out = torch.empty((buf.shape[0], n_total, buf.shape[2]), dtype=buf.dtype, device=buf.device)
        for r in range(self.world):
            lo, hi = shard_bounds(n_total, self.world, r)
            out[:, lo:hi].copy_(buf[:, r * m: r * m + (hi - lo)])
        return out

    def all_gather_rows_inplace(self, full, n_total):
        """full: [n_total, ...] tensor whose rows [lo_r, hi_r) are valid on rank r -> all rows valid everywhere."""
        if self.world == 1:
            return full
        lo, hi = shard_bounds(n_total, self.world, self.rank)
        full.copy_(self.all_gather_rows(full[lo:hi], n_total, 0))
        return full

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_ints(self, value, device):
        """One Python int per rank -> list of the ints of all ranks (rank order)."""
        import torch
        if self.world == 1:
            return [int(value)]
        mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
        out = torch.empty((self.world,), dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return [int(v) for v in out.tolist()]

    def all_gather_var(self, local, counts):
        """local: 1-D tensor whose first counts[rank] entries are valid -> concatenation over the ranks (rank order)."""
        import torch
        if self.world == 1:
            return local[: counts[0]]
        m = max(max(counts), 1)
        pad = torch.zeros((m,), dtype=local.dtype, device=local.device)
        pad[: counts[self.rank]] = local[: counts[self.rank]]
        out = torch.empty((self.world * m,), dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, pad, group=self.group)
        return torch.cat([out[r * m: r * m + counts[r]] for r in range(self.world)], 0)

    def broadcast(self, t, src):
        if self.world > 1:
            self.dist.broadcast(t, src=src, group=self.group)
        return t

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)


class CudaBackend(object):
    """The real compute: libssg_b200 through the plans of ssg_b200.{embed,rerank,cluster}."""

    def __init__(self, device=None, dist_mode=None, batch=256):
        from . import _lib
        from .cycle import _dist_mode
        self._lib = _lib
        self.dev = _lib.require_cuda(device)
        self.mode = _dist_mode(dist_mode)
        self.batch = batch

    def embed(self, model, images, num_split, out=None):
        from .embed import embed_images
        return embed_images(model, images, num_split, False, self.batch, self.dev.index, out=out)   # [banks, n_local, 2048]

    def plan(self, n, ns, d):
        from .rerank import get_plan
        return get_plan(n, ns, d, self.dev.index)

    def distance_rows(self, plan, src, tgt, k1, row0, rows):
        L = self._lib
        L.check(L.load().ssg_rerank_distance_rows(plan._h, src.data_ptr(), src.shape[0], tgt.data_ptr(), tgt.shape[0],
                                                  tgt.shape[1], int(k1), self.mode, int(row0), int(rows), None,
                                                  L.stream_ptr(self.dev)))

    def tables(self, plan, n):
        """Zero-copy torch views of the plan's per-row tables: rowmin [n], rowmax [n], rank [n,32], rank_val [n,32]."""
        import ctypes
        import torch
        L = self._lib
        ptrs = [ctypes.c_void_p() for _ in range(4)]
        L.check(L.load().ssg_rerank_tables(plan._h, *[ctypes.byref(p) for p in ptrs]))
        specs = [((n,), "<f4"), ((n,), "<f4"), ((n, L.RANK_STRIDE), "<i4"), ((n, L.RANK_STRIDE), "<f4")]
        return [torch.as_tensor(_DevArray(p.value, shape, ts), device=self.dev) for p, (shape, ts) in zip(ptrs, specs)]

    def finish(self, plan, tgt, k1, k2, lambda_value, final):
        L = self._lib
        L.check(L.load().ssg_rerank_finish(plan._h, tgt.data_ptr(), tgt.shape[0], tgt.shape[1], int(k1), int(k2),
                                           float(lambda_value), final.data_ptr(), L.stream_ptr(self.dev)))

    def new_final(self, n):
        import torch
        return torch.empty((n, n), dtype=torch.float64, device=self.dev)

    def eps(self, final, rho):
        from .cluster import get_plan
        return get_plan(final.shape[0], self.dev.index).eps(final, rho)[0]

    def dbscan(self, final, eps, min_samples):
        from .cluster import _with_capacity_retry
        n = final.shape[0]
        return _with_capacity_retry(n, self.dev.index, lambda p: p.dbscan(final, eps, min_samples)[0])


    # ---- sparse form of final_dist (sparse=True): no N x N matrix on the owner
    def finish_sparse(self, plan, tgt, k1, k2, lambda_value):
        return plan.finish_sparse(tgt, k1, k2, lambda_value)              # rowptr, col, val, bound

    def eps_sparse(self, n, rowptr, col, val, bound, rho):
        eps, _, ok = self.cluster_plan(n).eps_sparse(n, rowptr, col, val, bound, rho)
        return eps, ok

    def dbscan_sparse(self, n, rowptr, col, val, eps, min_samples):
        return self.with_capacity_retry(n, lambda p: p.dbscan_sparse(n, rowptr, col, val, eps, min_samples)[0])

    # ---- row-sharded finish (shard_finish=True)
    def finish_rows(self, plan, tgt, k1, k2, lambda_value, row0, rows, final_rows):
        L = self._lib
        L.check(L.load().ssg_rerank_finish_rows(plan._h, tgt.data_ptr(), tgt.shape[0], tgt.shape[1], int(k1), int(k2),
                                                float(lambda_value), int(row0), int(rows),
                                                final_rows.data_ptr() if rows else None, L.stream_ptr(self.dev)))

    def new_final_rows(self, rows, n):
        import torch
        return torch.empty((rows, n), dtype=torch.float64, device=self.dev)

    def cluster_plan(self, n, max_neighbors=0):
        from .cluster import get_plan
        return get_plan(n, self.dev.index, max_neighbors)

    def with_capacity_retry(self, n, fn):
        from .cluster import _with_capacity_retry
        return _with_capacity_retry(n, self.dev.index, fn)


class _DevArray(object):
    """Minimal __cuda_array_interface__ holder for a raw device pointer owned by a plan."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def sharded_eps(cplan, comm, rows, n, rho):
    """selftraining.py:289-293 over a row-sharded symmetric matrix (`rows` = this rank's [hi-lo, n] block): the radix
    select of ssg_eps_estimate with its histograms all-reduced and its candidate list all-gathered (call sequence:
    include/ssg_b200.h, "Row-sharded variants").  Returns (eps, top_num), the same on every rank."""
    from ._lib import EPS_LIST_CAP
    w, r = comm.world, comm.rank
    buf = cplan.buffers(n)
    cplan.eps_shard_begin()
    for npass in (0, 1):
        cplan.eps_shard_hist(rows, n, w, r, npass)
        comm.all_reduce_sum(buf["hist"])
        cplan.eps_shard_pick(npass, rho)
    count = cplan.eps_shard_gather(rows, n, w, r, exact=False)
    counts = comm.all_gather_ints(count, rows.device)
    if min(counts) >= 0 and sum(counts) <= EPS_LIST_CAP:
        comm.all_gather_rows_inplace(buf["partial"], n)
        merged = comm.all_gather_var(buf["list"], counts)
        buf["list"][: merged.shape[0]] = merged
        buf["state"][5] = sum(counts)
        return cplan.eps_shard_finish(n, exact=False)
    # massive ties: more than 2^20 entries share the threshold's leading 24 key bits -> finish the radix select
    for npass in (2, 3, 4, 5):
        cplan.eps_shard_hist(rows, n, w, r, npass)
        comm.all_reduce_sum(buf["hist"])
        cplan.eps_shard_pick(npass, rho)
    cplan.eps_shard_gather(rows, n, w, r, exact=True)
    comm.all_gather_rows_inplace(buf["partial"], n)
    return cplan.eps_shard_finish(n, exact=True)


def sharded_dbscan(backend, comm, rows, n, row0, eps, min_samples=4):
    """sklearn DBSCAN(precomputed) over a row-sharded matrix: local region queries, all-gather of the neighbour
    counts, sum all-reduce completing the neighbour CSR, then every rank labels all n rows (identical results).
    A neighbour list that outgrows the plan raises on every rank alike (the counts are global), so the retry with a
    larger plan stays collective."""
    def attempt(cplan):
        cplan.dbscan_shard_count(rows, n, row0, eps)
        comm.all_gather_rows_inplace(cplan.buffers(n)["cnt"], n)
        total = cplan.dbscan_shard_fill(rows, n, row0, eps)
        if total > 0:
            comm.all_reduce_sum(cplan.buffers(n, total)["nbr"])
        return cplan.dbscan_shard_label(n, min_samples)[0]
    return backend.with_capacity_retry(n, attempt)


def _shard_finish_default():
    import os
    return os.environ.get("SSG_SHARD_FINISH", "0") not in ("", "0")


def embed_and_gather(model, tgt_shard, src_shard, n_tgt, n_src, num_split=2, backend=None, comm=None):
    """Steps 1-2 of the sharded cycle with the exchange overlapped: embed this rank's target shard straight into the
    gather buffer, start its all-gather (asynchronous), embed the source shard while it runs, gather that too.
    Returns (tgt, src): [banks, n, d] float32 feature banks of ALL images on every rank."""
    import torch
    comm = comm or Comm()
    backend = backend or CudaBackend()
    banks = num_split + 1 if num_split > 1 else 1
    dev = getattr(backend, "dev", None) or tgt_shard.device
    d = getattr(backend, "feature_dim", 2048)
    bufs, works = [], []
    for shard, n in ((tgt_shard, n_tgt), (src_shard, n_src)):
        buf = comm.gather_buffer(banks, n, d, torch.float32, dev)
        backend.embed(model, shard, num_split, out=comm.gather_slot(buf, n))
        works.append(comm.gather_banks_begin(buf))
        bufs.append(buf)
    return (comm.gather_banks_finish(bufs[0], works[0], n_tgt), comm.gather_banks_finish(bufs[1], works[1], n_src))


def sharded_pseudo_label_cycle(model, tgt_shard, src_shard, n_tgt, n_src, num_split=2, lambda_value=0.1, rho=1.6e-3,
                               eps_list=None, min_samples=4, k1=20, k2=6, backend=None, comm=None, features=None,
                               shard_finish=None, sparse=None, features_full=None):
    """Run one pseudo-label cycle sharded over the ranks of `comm`.

    tgt_shard / src_shard: this rank's image rows [shard_bounds(n, world, rank)) (host or device tensors).
    `features=(tgt_local, src_local)` ([banks, n_local, d] each) skips the embedding (pre-extracted features);
    `features_full=(tgt, src)` ([banks, n, d], already gathered by embed_and_gather) skips the exchange as well.
    Returns (labels_list [np.int64 arrays], eps_list, keep_mask) — identical on every rank.
    """
    import torch
    comm = comm or Comm()
    backend = backend or CudaBackend()
    banks = num_split + 1 if num_split > 1 else 1
    if features_full is not None:
        tgt, src = features_full                 # already gathered (embed_and_gather): [banks, n, d] on every rank
    elif features is None:
        # the one real exchange step, overlapped with the embedding of the other set
        tgt, src = embed_and_gather(model, tgt_shard, src_shard, n_tgt, n_src, num_split, backend, comm)
    else:
        tloc, sloc = features
        tgt = comm.all_gather_rows(tloc, n_tgt, dim=1)
        src = comm.all_gather_rows(sloc, n_src, dim=1)
    lo, hi = shard_bounds(n_tgt, comm.world, comm.rank)
    plan = backend.plan(n_tgt, n_src, tgt.shape[2])
    k1d = max(k1, k2 - 1)        # the rank table must hold k2 columns too (rerank.py:97), as in ssg_rerank_run
    if _shard_finish_default() if shard_finish is None else shard_finish:
        # row-sharded finish: bank after bank, every rank works on its rows [lo, hi) of every stage
        final_rows = backend.new_final_rows(hi - lo, n_tgt)
        labels, eps_vals = [], []
        for b in range(banks):
            tb, sb = tgt[b].contiguous(), src[b].contiguous()
            backend.distance_rows(plan, sb, tb, k1d, lo, hi - lo)
            for tab in backend.tables(plan, n_tgt):
                comm.all_gather_rows_inplace(tab, n_tgt)
            backend.finish_rows(plan, tb, k1, k2, lambda_value, lo, hi - lo, final_rows)
            cplan = backend.cluster_plan(n_tgt)
            eps = sharded_eps(cplan, comm, final_rows, n_tgt, rho)[0] if eps_list is None else float(eps_list[b])
            lab = sharded_dbscan(backend, comm, final_rows, n_tgt, lo, eps, min_samples)
            labels.append(lab.cpu().numpy())
            eps_vals.append(eps)
        keep = ~(np.stack(labels, 0) == -1).any(0)
        return labels, eps_vals, keep
    if sparse is None:
        from .cycle import _sparse_default
        sparse = _sparse_default()
    use_sparse = bool(sparse) and 0.0 <= lambda_value < 1.0
    # phase A — all ranks, bank after bank: row-block distance stage, then all-gather of the per-row tables; the owner
    # of a bank keeps a copy of its tables (the plan holds one set).  Phase B — the owners finish their banks in
    # parallel (k-reciprocal encoding ... final distance, eps, DBSCAN); nobody waits for an owner between banks.
    saved = []
    bank_t = []
    for b in range(banks):
        tb, sb = tgt[b].contiguous(), src[b].contiguous()
        bank_t.append(tb)
        backend.distance_rows(plan, sb, tb, k1d, lo, hi - lo)
        tabs = backend.tables(plan, n_tgt)
        for tab in tabs:
            comm.all_gather_rows_inplace(tab, n_tgt)
        saved.append([t.clone() for t in tabs] if comm.rank == b % comm.world else None)
    final = None
    labels_dev, eps_out = [], []
    for b in range(banks):
        tb = bank_t[b]
        lab = torch.empty((n_tgt,), dtype=torch.int64, device=tb.device)
        eps_t = torch.zeros((1,), dtype=torch.float64, device=tb.device)
        if comm.rank == b % comm.world:
            for dst, keep in zip(backend.tables(plan, n_tgt), saved[b]):
                dst.copy_(keep)
            done = False
            if use_sparse:
                # the owner never materialises final_dist (cycle.pseudo_label_cycle) unless the
                # shortcut cannot be certified for this bank
                rowptr, col, val, bound = backend.finish_sparse(plan, tb, k1, k2, lambda_value)
                if eps_list is None:
                    eps, ok = backend.eps_sparse(n_tgt, rowptr, col, val, bound, rho)
                else:
                    eps, ok = float(eps_list[b]), True
                if ok and eps < bound:
                    lab.copy_(backend.dbscan_sparse(n_tgt, rowptr, col, val, eps, min_samples))
                    done = True
            if not done:
                if final is None:
                    final = backend.new_final(n_tgt)
                backend.finish(plan, tb, k1, k2, lambda_value, final)
                eps = backend.eps(final, rho) if eps_list is None else float(eps_list[b])
                lab.copy_(backend.dbscan(final, eps, min_samples))
            eps_t[0] = eps
        labels_dev.append(lab)
        eps_out.append(eps_t)
    for b in range(banks):
        comm.broadcast(labels_dev[b], b % comm.world)
        comm.broadcast(eps_out[b], b % comm.world)
    labels = [l.cpu().numpy() for l in labels_dev]
    eps_vals = [float(e.item()) for e in eps_out]
    keep = ~(np.stack(labels, 0) == -1).any(0)
    return labels, eps_vals, keep
