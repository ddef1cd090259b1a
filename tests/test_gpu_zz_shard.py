"""GPU: the row-sharded finish (include/ssg_b200.h "Row-sharded variants").

1. tests/c/shard_check.c -- a plain C program on the C ABI -- simulates `world` ranks one after another on ONE device
   and compares rows of final_dist, eps (3-pass and 6-pass) and DBSCAN labels with the single-GPU entry points.
2. ssg_b200.dist.sharded_pseudo_label_cycle(shard_finish=True) at world size 1 (no process group) against
   ssg_b200.pseudo_label_cycle: the Python choreography and the ctypes wrappers on the device.
The collective choreography itself at world sizes 2-5 runs on CPU (tests/test_dist_gloo.py)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CDIR = os.path.join(ROOT, "tests", "c")


@pytest.mark.gpu
@pytest.mark.parametrize("n,world", [(1500, 3), (2049, 8), (777, 1)])
def test_c_harness_sharded_entry_points_match_single_gpu(n, world):
    subprocess.call(["make", "-C", CDIR], stdout=subprocess.DEVNULL)      # no-op when build() already made it
    r = subprocess.run([os.path.join(CDIR, "_build", "shard_check"), str(n), str(world)], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "SHARD_CHECK PASSED" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_row_sharded_cycle_world1_matches_unsharded():
    import torch
    import ssg_b200
    from ssg_b200 import _lib, dist as sd
    from oracle import ssg_oracle as O
    n, ns, d, banks, rho, lam = 1300, 900, 128, 3, 0.02, 0.1
    tgt = torch.from_numpy(np.stack([O.synth_features(n, d, 10 + b)[0] for b in range(banks)])).cuda()
    src = torch.from_numpy(np.stack([O.synth_features(ns, d, 20 + b, noise=0.6)[0] for b in range(banks)])).cuda()
    want_l, want_e, want_k = ssg_b200.pseudo_label_cycle([src[b] for b in range(banks)], [tgt[b] for b in range(banks)],
                                                         lam, rho, dist_mode=_lib.DIST_EXACT, device=0)
    got_l, got_e, got_k = sd.sharded_pseudo_label_cycle(
        None, None, None, n, ns, num_split=banks - 1, lambda_value=lam, rho=rho,
        backend=sd.CudaBackend(0, _lib.DIST_EXACT), comm=sd.Comm(), features=(tgt, src), shard_finish=True)
    np.testing.assert_allclose(got_e, want_e, rtol=1e-13, atol=0)
    for a, b in zip(got_l, want_l):
        assert np.array_equal(a, b)
    assert np.array_equal(got_k, want_k)
    assert max(int(l.max()) for l in want_l) >= 1
    # frozen eps (iterations > 0, selftraining.py:296-298) takes the same path without the eps stage
    again, _, _ = sd.sharded_pseudo_label_cycle(
        None, None, None, n, ns, num_split=banks - 1, lambda_value=lam, rho=rho, eps_list=want_e,
        backend=sd.CudaBackend(0, _lib.DIST_EXACT), comm=sd.Comm(), features=(tgt, src), shard_finish=True)
    for a, b in zip(again, want_l):
        assert np.array_equal(a, b)
