#!/usr/bin/env python
"""The whole embedding forward (ssg_embed_* through the C ABI) on the CPU: the library built against the functional
tcgen05 / TMA / mbarrier emulation (build_emu.build_tc), the reference's golden images and weights, features compared
with the reference's golden features and written to a file -- kernel variants (environment, read once per process) are
compared by running this script once per variant and diffing the files.  TEST INFRASTRUCTURE.
    python tests/cpu_cuda/run_embed_emu.py N_IMAGES OUT.npy"""
import sys, os, ctypes, time, numpy as np
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [_ROOT, os.path.join(_ROOT, "self-similarity-grouping_b200"), os.path.join(_ROOT, "tests", "cpu_cuda")]
import build_emu
import torch
from ssg_b200 import _lib as L
from oracle import resnet_oracle as R
lib = ctypes.CDLL(build_emu.build_tc())
for name, (res, args) in L.PROTOTYPES.items():
    fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
g = np.load(os.path.join(_ROOT, 'tests', 'golden', 'embed_4img.npz'))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
imgs = np.ascontiguousarray(R.synth_images(int(g["n_img"]), int(g["seed_img"]))[:n].numpy(), np.float32)
model = R.build_model(2, int(g["weight_seed"]))
sd = {k: np.ascontiguousarray(v.detach().numpy(), np.float32) for k, v in model.base.state_dict().items() if v.dtype.is_floating_point}
t0 = time.time()
plan = ctypes.c_void_p(); assert lib.ssg_embed_plan_create(ctypes.byref(plan), 0, max(n, 1), 256, 128) == 0
ck, bk = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
keep = []
for idx in range(lib.ssg_embed_num_layers()):
    ci, co, k, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.ssg_embed_layer_info(idx, ctypes.byref(ci), ctypes.byref(co), ctypes.byref(k), ctypes.byref(s), ck, bk, 64) == 0
    c, b = ck.value.decode(), bk.value.decode()
    arrs = [sd[c + ".weight"], sd[b + ".weight"], sd[b + ".bias"], sd[b + ".running_mean"], sd[b + ".running_var"]]
    keep.append(arrs)
    rc = lib.ssg_embed_load_layer(plan, idx, *[a.ctypes.data for a in arrs], 1e-5, None)
    assert rc == 0, lib.ssg_last_error().decode()
print("weights folded", round(time.time() - t0, 1), "s", flush=True)
feat = np.zeros((3, n, 2048), np.float32)
rc = lib.ssg_embed_forward(plan, imgs.ctypes.data, n, 2, 0, 1, feat.ctypes.data, n * 2048, 0, None)
assert rc == 0, lib.ssg_last_error().decode()
print("forward", round(time.time() - t0, 1), "s")
rel = max(float(np.linalg.norm(feat[b, i] - g["list_S2"][b, i])) for i in range(n) for b in range(3))
print("rel err vs reference golden:", rel, "finite:", bool(np.isfinite(feat).all()))
np.save(sys.argv[2] if len(sys.argv) > 2 else "/tmp/emb_emu.npy", feat)
