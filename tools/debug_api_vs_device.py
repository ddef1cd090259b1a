#!/usr/bin/env python
"""GPU: why does the reference-shaped API leg of bench.py give another eps than the device-resident cycle?"""
import os, sys, json, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-similarity-grouping_b200")]
import numpy as np, torch
import ssg_b200
from ssg_b200 import synth, _lib
import reid.evaluators as E, reid.rerank as RR
from reid.cluster import eps_estimate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16702
dev = torch.device("cuda", 0)
model = synth.build_model(2, 0)
tgt_img, _ = synth.synth_images(n, 1234, dev)
src_img, _ = synth.synth_images(n, 4321, dev)
host_t, host_s = tgt_img.cpu().pin_memory(), src_img.cpu().pin_memory()
tf = ssg_b200.embed_images(model, host_t, 2, False, 512, 0)
sf = ssg_b200.embed_images(model, host_s, 2, False, 512, 0)
tf_dev = ssg_b200.embed_images(model, tgt_img, 2, False, 512, 0)
out = {"embed_host_vs_device_images_equal": bool(torch.equal(tf, tf_dev))}
feats = {}
for tag, host in (("src", host_s), ("tgt", host_t)):
    names = ["%s%06d" % (tag, i) for i in range(n)]
    loader = [(host[i:i + 512], names[i:i + 512], [0] * len(names[i:i + 512]), [0] * len(names[i:i + 512])) for i in range(0, n, 512)]
    f, _ = E.extract_features(model, loader, print_freq=10 ** 9, for_eval=False)
    feats[tag] = [torch.cat([f[nm][i].unsqueeze(0) for nm in names], 0) for i in range(3)]
for b in range(3):
    d = {}
    d["tgt_features_equal"] = bool(torch.equal(feats["tgt"][b], tf[b].cpu()))
    d["src_features_equal"] = bool(torch.equal(feats["src"][b], sf[b].cpu()))
    d["tgt_features_max_abs"] = float((feats["tgt"][b] - tf[b].cpu()).abs().max())
    t, s = tf[b].contiguous(), sf[b].contiguous()
    plan = ssg_b200.rerank.get_plan(n, n, 2048, 0)
    _, f_dev = plan.run(s, t, 20, 6, 0.1, _lib.DIST_TENSOR)
    f_dev = f_dev.clone()
    with contextlib.redirect_stdout(io.StringIO()):
        _, f_host = RR.re_ranking(s.cpu().numpy(), t.cpu().numpy(), lambda_value=0.1)
    d["final_host_equals_device"] = bool(np.array_equal(f_host, f_dev.cpu().numpy()))
    cp = ssg_b200.cluster.get_plan(n, 0)
    d["eps_device"] = cp.eps(f_dev, 1.6e-3)[0]
    d["eps_host_same_matrix"] = eps_estimate(f_dev.cpu().numpy(), 1.6e-3)
    d["eps_host_api_matrix"] = eps_estimate(f_host, 1.6e-3)
    out["bank%d" % b] = d
    del f_host, f_dev
print(json.dumps(out, indent=1))
