#!/usr/bin/env python
"""GPU: do the tensor and the exact distance mode give the same final_dist / eps / labels on REAL trunk features
(nearly identical rows, d^2 ~ 1e-4: the regime where the bench's reference-API leg disagreed with the cycle)?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-similarity-grouping_b200")]
import numpy as np, torch
import ssg_b200
from ssg_b200 import synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
model = synth.build_model(2, 0)
tgt_img, _ = synth.synth_images(n, 1234, dev)
src_img, _ = synth.synth_images(n, 4321, dev)
tf = ssg_b200.embed_images(model, tgt_img, 2, False, 512, 0)
sf = ssg_b200.embed_images(model, src_img, 2, False, 512, 0)
out = {}
for b in range(3):
    t, s = tf[b].contiguous(), sf[b].contiguous()
    plan = ssg_b200.rerank.get_plan(n, n, 2048, 0)
    _, f_ex = plan.run(s, t, 20, 6, 0.1, _lib.DIST_EXACT)
    f_ex = f_ex.clone(); rank_ex = plan.stage(_lib.STAGE_RANK, n)[:, :21].copy(); vec_ex = plan.stage(_lib.STAGE_VEC, n).copy()
    _, f_te = plan.run(s, t, 20, 6, 0.1, _lib.DIST_TENSOR)
    torch.cuda.synchronize()
    rank_te = plan.stage(_lib.STAGE_RANK, n)[:, :21].copy(); vec_te = plan.stage(_lib.STAGE_VEC, n).copy()
    flagged = plan.stage(_lib.STAGE_FLAGGED, n)
    cp = ssg_b200.cluster.get_plan(n, 0)
    e_ex, e_te = cp.eps(f_ex, 1.6e-3)[0], cp.eps(f_te, 1.6e-3)[0]
    l_ex = cp.dbscan(f_ex, e_ex, 4)[0].cpu().numpy(); l_te = cp.dbscan(f_te, e_te, 4)[0].cpu().numpy()
    # the host API path (numpy in / out) in its default mode
    e_host, f_host = ssg_b200.re_ranking(s.cpu().numpy(), t.cpu().numpy(), lambda_value=0.1)
    out["bank%d" % b] = dict(final_equal=bool(torch.equal(f_ex, f_te)), final_max_abs=float((f_ex - f_te).abs().max()),
                             rank_rows_differ=int((rank_ex != rank_te).any(1).sum()), vec_equal=bool(np.array_equal(vec_ex, vec_te)),
                             vec_max_abs=float(np.abs(vec_ex - vec_te).max()), flagged=[int(x) for x in np.ravel(flagged)[:4]],
                             eps=[e_ex, e_te], labels_differ=int((l_ex != l_te).sum()),
                             host_equal_exact=bool(np.array_equal(f_host, f_ex.cpu().numpy())))
print(json.dumps(out, indent=1))
