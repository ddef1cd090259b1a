"""CPU: the REAL Python host layer -- ssg_b200.rerank / cluster / cycle / dist with their ctypes wrappers, zero-copy
buffer views, capacity retries and collectives -- running end to end against the CPU-emulated library
(tests/emu_device.py + tests/cpu_cuda), single process and over gloo.  tests/test_dist_gloo.py exercises the
choreography with a numpy stand-in for the compute; here the compute is the library's own kernel source.  Exact
distance mode (the tensor-core kernels are not emulated)."""
import os
import shutil
import sys
import tempfile

import numpy as np
import pytest

from oracle import ssg_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
N, NS, D, BANKS, RHO, LAM = 120, 70, 16, 3, 0.03, 0.1


def _features(banks=BANKS):
    import torch
    tgt = np.stack([O.synth_features(N, D, 10 + b, per_cluster=12, noise=0.3)[0] for b in range(banks)])
    src = np.stack([O.synth_features(NS, D, 20 + b, per_cluster=12, noise=0.4)[0] for b in range(banks)])
    return torch.from_numpy(tgt), torch.from_numpy(src)


def _oracle_cycle(banks=BANKS):
    tgt, src = _features(banks)
    labels, eps = [], []
    for b in range(banks):
        _, f = O.re_ranking(src[b].numpy(), tgt[b].numpy(), lambda_value=LAM, mode="f32")
        e = O.eps_estimate(f, RHO)
        eps.append(e)
        labels.append(O.dbscan_dfs(f, e, 4))
    return labels, eps


@pytest.fixture()
def emulated():
    sys.path.insert(0, HERE)
    import emu_device
    undo = emu_device.install()
    yield
    undo()


def test_pseudo_label_cycle_dense_and_sparse_single_process(emulated):
    """ssg_b200.pseudo_label_cycle (what bench.py times) with the dense matrix and with sparse=True, and the row-sharded /
    sparse-owner variants of the sharded cycle at world size 1: identical labels, eps to 1e-12 of the oracle."""
    import ssg_b200
    from ssg_b200 import _lib, dist as sd
    tgt, src = _features()
    want_l, want_e = _oracle_cycle()
    assert max(int(l.max()) for l in want_l) >= 1
    tl, sl = [tgt[b] for b in range(BANKS)], [src[b] for b in range(BANKS)]
    runs = {
        "dense": ssg_b200.pseudo_label_cycle(sl, tl, LAM, RHO, dist_mode=_lib.DIST_EXACT, sparse=False),
        "sparse": ssg_b200.pseudo_label_cycle(sl, tl, LAM, RHO, dist_mode=_lib.DIST_EXACT, sparse=True),
    }
    for name, kw in (("sharded", {}), ("row-sharded finish", {"shard_finish": True}), ("sparse owners", {"sparse": True})):
        runs[name] = sd.sharded_pseudo_label_cycle(None, None, None, N, NS, num_split=BANKS - 1, lambda_value=LAM, rho=RHO,
                                                   backend=sd.CudaBackend(0, _lib.DIST_EXACT), comm=sd.Comm(),
                                                   features=(tgt, src), **kw)
    # against the oracle: final_dist agrees to ~1e-7 (exp() of the source term), hence eps to ~1e-7 as well
    np.testing.assert_allclose(runs["dense"][1], want_e, rtol=1e-6, atol=0)
    for name, (labels, eps, keep) in runs.items():
        # every variant works on the same matrix as the dense run: eps to the order of the float64 additions
        np.testing.assert_allclose(eps, runs["dense"][1], rtol=1e-12, atol=0, err_msg=name)
        for a, b in zip(labels, want_l):
            assert np.array_equal(a, b), name
        assert np.array_equal(keep, O.keep_mask(want_l)), name
    # tensor distance mode (the cycle's default on the GPU; stand-in GEMM, real certification): same results
    for sparse in (False, True):
        labels, eps, _ = ssg_b200.pseudo_label_cycle(sl, tl, LAM, RHO, dist_mode=_lib.DIST_TENSOR, sparse=sparse)
        np.testing.assert_allclose(eps, runs["dense"][1], rtol=1e-12, atol=0)
        for a, b in zip(labels, want_l):
            assert np.array_equal(a, b)
    # frozen eps (iterations > 0) through the sparse form
    labels, _, _ = ssg_b200.pseudo_label_cycle(sl, tl, LAM, RHO, eps_list=want_e, dist_mode=_lib.DIST_EXACT, sparse=True)
    for a, b in zip(labels, want_l):
        assert np.array_equal(a, b)
    # a rho whose slice cannot be certified falls back to the dense matrix, silently and correctly
    l_d, e_d, _ = ssg_b200.pseudo_label_cycle(sl, tl, LAM, 0.9, dist_mode=_lib.DIST_EXACT, sparse=False)
    l_s, e_s, _ = ssg_b200.pseudo_label_cycle(sl, tl, LAM, 0.9, dist_mode=_lib.DIST_EXACT, sparse=True)
    np.testing.assert_allclose(e_s, e_d, rtol=1e-12)
    for a, b in zip(l_s, l_d):
        assert np.array_equal(a, b)


def test_k2_beyond_k1_plus_one_through_every_host_path(emulated):
    """rerank.py:97 reads k2 rank columns; for k2 > k1 + 1 the distance stage must be run wide enough.  ssg_rerank_run
    does that itself; the host paths that call the distance stage on their own (sparse cycle, sharded cycle) must too --
    the finish refuses a table that is too narrow instead of reading stale columns."""
    import ssg_b200
    from ssg_b200 import _lib, dist as sd
    k1, k2 = 3, 7
    tgt, src = _features()
    want = []
    for b in range(BANKS):
        _, f = O.re_ranking(src[b].numpy(), tgt[b].numpy(), k1=k1, k2=k2, lambda_value=LAM, mode="f32")
        want.append(O.dbscan_dfs(f, O.eps_estimate(f, RHO), 4))
    tl, sl = [tgt[b] for b in range(BANKS)], [src[b] for b in range(BANKS)]
    runs = {name: ssg_b200.pseudo_label_cycle(sl, tl, LAM, RHO, k1=k1, k2=k2, dist_mode=_lib.DIST_EXACT, sparse=sp)
            for name, sp in (("dense", False), ("sparse", True))}
    for name, kw in (("sharded", {}), ("row-sharded finish", {"shard_finish": True}), ("sparse owners", {"sparse": True})):
        runs[name] = sd.sharded_pseudo_label_cycle(None, None, None, N, NS, num_split=BANKS - 1, lambda_value=LAM, rho=RHO,
                                                   k1=k1, k2=k2, backend=sd.CudaBackend(0, _lib.DIST_EXACT),
                                                   comm=sd.Comm(), features=(tgt, src), **kw)
    for name, (labels, _, _) in runs.items():
        for a, b in zip(labels, want):
            assert np.array_equal(a, b), name
    # and the refusal itself: a table of k1 + 1 columns is not enough for this k2
    plan = ssg_b200.RerankPlan(N, NS, D, 0)
    plan.distance_rows(src[0], tgt[0], k1, _lib.DIST_EXACT)
    with pytest.raises(Exception, match="rank columns"):
        plan.finish_sparse(tgt[0], k1, k2, LAM)


def test_drop_in_functions_on_the_emulated_library(emulated):
    """reid.rerank.re_ranking / DBSCAN and reid.rerank_plain.re_ranking (numpy in, numpy out) against the oracle."""
    from reid.rerank import re_ranking, DBSCAN
    from reid import rerank_plain
    from oracle import rerank_plain_oracle as P
    tgt, src = _features()
    t, s = tgt[0].numpy(), src[0].numpy()
    e, f = re_ranking(s, t, lambda_value=LAM)
    e_ref, f_ref = O.re_ranking(s, t, lambda_value=LAM, mode="f32")
    assert np.array_equal(e, e_ref)
    np.testing.assert_allclose(f, f_ref, rtol=0, atol=1e-4)
    eps = O.eps_estimate(f_ref, RHO)
    est = DBSCAN(eps=eps, min_samples=4, metric="precomputed", n_jobs=8)
    assert np.array_equal(est.fit_predict(f), O.dbscan_dfs(f_ref, eps, 4))
    assert np.array_equal(est.fit_predict(f), est.labels_)                       # re-usable, as cluster_list needs
    fp, fp2 = rerank_plain.re_ranking(s, t, 20, LAM)
    assert fp is fp2
    np.testing.assert_allclose(fp, P.re_ranking_plain(s, t, k=20, lambda_value=LAM, mode="f32"), rtol=0, atol=2e-6)


def _worker(rank, world, init_file, out_dir, kw):
    import torch.distributed as dist
    sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "self-similarity-grouping_b200")]
    import emu_device
    emu_device.install()
    from ssg_b200 import _lib, dist as sd
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        kw = dict(kw)
        banks = kw.pop("banks", BANKS)
        tgt, src = _features(banks)
        tl, th = sd.shard_bounds(N, world, rank)
        sl, sh = sd.shard_bounds(NS, world, rank)
        labels, eps, keep = sd.sharded_pseudo_label_cycle(
            None, None, None, N, NS, num_split=banks - 1, lambda_value=LAM, rho=RHO, backend=sd.CudaBackend(0, _lib.DIST_EXACT),
            comm=sd.Comm(), features=(tgt[:, tl:th].contiguous(), src[:, sl:sh].contiguous()), **kw)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), labels=np.stack(labels), eps=np.array(eps), keep=keep)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kw", [(2, {}), (3, {"shard_finish": True}), (2, {"sparse": True}), (4, {"shard_finish": True}),
                                      (2, {"banks": 4})])      # BASELINE configs[2]: num_split = 3 -> 4 banks on 2 ranks
def test_sharded_cycle_with_the_real_backend_over_gloo(world, kw):
    """One process per rank, gloo collectives, ssg_b200.dist.CudaBackend driving the emulated library: feature
    all-gather, row-block distance stage, table gather, then the bank-parallel / row-sharded / sparse-owner finish."""
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(HERE, "cpu_cuda"))
    import build_emu
    build_emu.build()                                    # once, before the ranks start
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, os.path.join(tmp, "init"), tmp, kw), nprocs=world, join=True)
        outs = [np.load(os.path.join(tmp, "rank%d.npz" % r)) for r in range(world)]
    want_l, want_e = _oracle_cycle(kw.get("banks", BANKS))
    for o in outs:
        np.testing.assert_allclose(o["eps"], want_e, rtol=1e-6, atol=0)
        assert np.array_equal(o["labels"], np.stack(want_l))
        assert np.array_equal(o["keep"], O.keep_mask(want_l))
        assert np.array_equal(o["eps"], outs[0]["eps"])
