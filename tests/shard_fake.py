"""numpy stand-in for the row-sharded eps / DBSCAN primitives of ssg_b200.cluster.ClusterPlan (csrc/cluster.cu,
"row-sharded eps / DBSCAN"), used by the gloo tests to run the collective choreography of ssg_b200.dist on CPU.

Same buffers (torch CPU tensors instead of device views), same call sequence, same per-pass semantics: 12-bit radix
passes over order-preserving 64-bit keys, the pair-ownership rule of the sharded scan, a global neighbour CSR that is
completed by a sum all-reduce."""
import numpy as np

SHIFT = [52, 40, 28, 16, 4, 0]
WIDTH = [12, 12, 12, 12, 12, 4]
EPS_BINS = 4096


def f64_key(v):
    b = np.ascontiguousarray(v, dtype=np.float64).view(np.uint64)
    return np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))


def f64_from_key(k):
    k = np.uint64(k)
    b = (k & np.uint64(0x7fffffffffffffff)) if (k >> np.uint64(63)) else ~k
    return np.array([b], dtype=np.uint64).view(np.float64)[0]


def shard_lo(n, world, r):
    base, rem = divmod(n, world)
    return r * base + min(r, rem)


def takes(rank, c):
    if c == rank:
        return True
    odd = (rank + c) & 1
    return bool(odd) if rank < c else not odd


def visited_values(rows, n, world, rank):
    """Per local row: the values of the columns this rank visits (cluster.cu ShardGeom / shard_row_scan)."""
    lo = shard_lo(n, world, rank)
    out = []
    for li in range(rows.shape[0]):
        i = lo + li
        parts = []
        for c in range(world):
            if not takes(rank, c):
                continue
            j0 = i + 1 if c == rank else shard_lo(n, world, c)
            parts.append(rows[li, j0:shard_lo(n, world, c + 1)])
        out.append(np.concatenate(parts) if parts else np.zeros(0))
    return out


class FakeClusterPlan(object):
    def __init__(self, n, max_neighbors=0):
        import torch
        self.n_max = n
        self.max_neighbors = max_neighbors or 64 * n + (1 << 20)
        self.hist = torch.zeros(EPS_BINS, dtype=torch.int64)
        self.state = torch.zeros(8, dtype=torch.int64)
        self.partial = torch.zeros(n, dtype=torch.float64)
        self.list = torch.zeros(1 << 20, dtype=torch.float64)
        self.cnt = torch.zeros(n, dtype=torch.int32)
        self.nbr = torch.zeros(self.max_neighbors, dtype=torch.int32)
        self.rowptr = None

    def buffers(self, n, nbr_len=0):
        return {"hist": self.hist, "state": self.state, "partial": self.partial[:n], "list": self.list,
                "cnt": self.cnt[:n], "nbr": self.nbr[:max(nbr_len, 1)]}

    # ---- eps
    def eps_shard_begin(self):
        self.hist.zero_()
        self.state.zero_()

    def _st(self):
        return self.state.numpy().view(np.uint64)

    def eps_shard_hist(self, rows, n, world, rank, npass):
        st = self._st()
        shift, width = SHIFT[npass], WIDTH[npass]
        hs = np.uint64(shift + width)
        h = self.hist.numpy()
        for vals in visited_values(rows.numpy(), n, world, rank):
            vals = vals[vals != 0.0]
            k = f64_key(vals)
            if npass > 0:
                k = k[(k >> hs) == (st[0] >> hs)] if shift + width < 64 else k
            bins = ((k >> np.uint64(shift)) & np.uint64((1 << width) - 1)).astype(np.int64)
            np.add.at(h, bins, 1)

    def eps_shard_pick(self, npass, rho):
        st = self._st()
        h = self.hist.numpy().astype(np.uint64)
        total = int(h.sum())
        if npass == 0:
            top = int(np.rint(rho * float(total)))
            top = min(max(top, 0), total)
            st[3], st[2], st[1], st[0] = total, top, top, 0
        rem = int(st[1])
        if rem > 0:
            cum = np.concatenate([[0], np.cumsum(h.astype(np.int64))])
            b = int(np.searchsorted(cum[1:], rem, side="left"))      # first bin with cum[b+1] >= rem
            st[0] = st[0] | (np.uint64(b) << np.uint64(SHIFT[npass]))
            st[1] = rem - int(cum[b])
        self.hist.zero_()

    def eps_shard_gather(self, rows, n, world, rank, exact):
        st = self._st()
        lo = shard_lo(n, world, rank)
        part = self.partial.numpy()
        lst = self.list.numpy()
        pos = int(st[5])
        if int(st[2]) == 0:
            part[lo:lo + rows.shape[0]] = 0.0
            return None if exact else pos
        for li, vals in enumerate(visited_values(rows.numpy(), n, world, rank)):
            vals = vals[vals != 0.0]
            k = f64_key(vals)
            if exact:
                part[lo + li] = vals[k < st[0]].sum()
                continue
            h = k >> np.uint64(40)
            pre = st[0] >> np.uint64(40)
            part[lo + li] = vals[h < pre].sum()
            inb = vals[h == pre]
            lst[pos:pos + len(inb)] = inb
            pos += len(inb)
        st[5] = pos
        return None if exact else pos

    def eps_shard_finish(self, n, exact):
        st = self._st()
        top = int(st[2])
        if top == 0:
            return float("nan"), 0
        psum = float(self.partial.numpy()[:n].sum())
        if exact:
            thr = f64_from_key(st[0])
            return (psum + float(int(st[1])) * thr) / top, top
        lst = np.sort(self.list.numpy()[:int(st[5])])
        rem = int(st[1])
        thr = lst[rem - 1]
        below = lst[lst < thr]
        return (psum + float(below.sum()) + float(rem - len(below)) * thr) / top, top

    # ---- DBSCAN
    def dbscan_shard_count(self, rows, n, row0, eps):
        self.cnt[row0:row0 + rows.shape[0]] = (rows <= eps).sum(1).to(self.cnt.dtype)

    def dbscan_shard_fill(self, rows, n, row0, eps):
        cnt = self.cnt.numpy()[:n].astype(np.int64)
        self.rowptr = np.concatenate([[0], np.cumsum(cnt)])
        total = int(self.rowptr[n])
        if total > self.max_neighbors:
            raise OverflowError("fake: neighbour list overflow")
        nb = self.nbr.numpy()
        nb[:total] = 0
        r = rows.numpy()
        for li in range(r.shape[0]):
            j = np.nonzero(r[li] <= eps)[0][::-1]      # order inside a row is irrelevant: scramble it
            nb[self.rowptr[row0 + li]:self.rowptr[row0 + li + 1]] = j
        return total

    def dbscan_shard_label(self, n, min_samples=4):
        """Order-free labelling from the neighbour CSR, as db_union / db_label do."""
        import torch
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import connected_components
        rp, nb = self.rowptr, self.nbr.numpy()
        cnt = self.cnt.numpy()[:n]
        core = cnt >= min_samples
        labels = np.full(n, -1, dtype=np.int64)
        ci = np.where(core)[0]
        if len(ci):
            src = np.repeat(np.arange(n), cnt)
            dst = nb[:rp[n]]
            m = core[src] & core[dst]
            A = csr_matrix((np.ones(m.sum(), dtype=np.int8), (src[m], dst[m])), shape=(n, n))
            _, comp = connected_components(A[np.ix_(ci, ci)], directed=False)
            first = {}
            for pos, c in enumerate(comp):
                first.setdefault(c, pos)
            remap = {c: r for r, c in enumerate(sorted(first, key=lambda c: first[c]))}
            labels[ci] = [remap[c] for c in comp]
            for i in np.where(~core)[0]:
                js = nb[rp[i]:rp[i + 1]]
                js = js[core[js]]
                if len(js):
                    labels[i] = labels[js].min()
        ncl = int(labels.max()) + 1 if len(ci) else 0
        return torch.from_numpy(labels), ncl
