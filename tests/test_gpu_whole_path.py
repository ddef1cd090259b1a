"""Whole-path parity gate (SURVEY.md §4 "End-to-end parity"): IMAGES -> labels.

The reference side was produced by the UNMODIFIED reference in the build container (oracle/make_goldens.py
cycle_case -> tests/golden/cycle_n512_S2.npz): reid.evaluators.extract_features (fp32 torch CPU, list mode) on 512
target + 384 source seeded identity images, the driver's own bank re-stacking, compute_dist and generate_selflabel
(selftraining.py:196-218, 255-313) in the O-f32 arithmetic (the bit-/1e-4 parity target) and as-is (fp16).

The CUDA side goes through the same driver-shaped calls: reid.evaluators.extract_features (dict of CPU tensors) ->
re-stacking -> ssg_b200.cycle.compute_dist -> ssg_b200.cycle.generate_selflabel.  What bf16 convolutions (relative
feature error ~4e-3) do to rank tables, eps and labels is MEASURED here (round 2, B200: gpurun_out/ ->
profiles/r02_whole_path_parity.json) and bounded.

Conditioning of this input (stored in the golden): a random-init ResNet-50 maps all images to nearly the same
direction -- the median squared distance between two unit-norm bank rows is 2.4e-4 .. 4.9e-4 and the nearest
neighbour sits at 1.3e-4 .. 2.6e-4 -- so a feature error of relative size 4e-3 is ~40 % of a neighbour distance.  The
golden therefore also holds what that error size does to the REFERENCE ITSELF: the oracle (O-f32) re-run on its own
features perturbed by Gaussian noise of relative L2 size 4e-3 per row changes 52-60 % of its own rank-table entries,
the top-21 set of 85-94 % of its rows, and its labels to ARI 0.57-0.77 / 0.84-0.89 / 0.99-1.0 at rho = 1.6e-3 /
1.6e-2 / 5e-2.  The CUDA path measures 57-60 %, 94-95 %, ARI 0.57-0.84 / 0.86-0.89 / 0.992-1.0: it behaves like the
reference under a perturbation of its feature-error size, which is the most an implementation with that feature
error can do.  Bounds asserted:

  * features: relative L2 error of every bank row <= 8e-3 (measured 4.1e-3; the round-1 bound was 3e-2);
  * rank tables: entry / set mismatch <= the reference's own sensitivity to 4e-3 feature noise + 0.10 / + 0.15;
  * eps: within 3 % of the reference's at every rho, within 1 % at the well-conditioned rho = 5e-2;
  * labels: ARI(GPU, reference O-f32) >= the reference's own noise-ARI - 0.10 per bank and rho, and >= 0.99 at
    rho = 5e-2, where the reference recovers the 64 identities (ARI 0.99 against the truth);
  * against the unmodified fp16 reference the ARI is reported next to ARI(O-f32, fp16 reference) -- the two CPU
    arithmetics differ from each other by ARI 0.83-0.98 at the driver's rho values.
Rank tables and labels ARE bit-exact against the oracle when the comparison starts from the same features
(tests/test_gpu_rerank.py, test_gpu_cluster.py, test_gpu_api_rows.py); distances are within 1e-4 there.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _restack(features, names, banks):
    """selftraining.py:197-209."""
    import torch
    return [torch.cat([features[f][i].unsqueeze(0) for f in names], 0) for i in range(banks)]


def whole_path_metrics(golden_path, batch=64):
    """Runs the CUDA path on the golden's inputs and returns the parity metrics (also used by bench.py)."""
    import types
    import contextlib
    import io
    import torch
    from sklearn.metrics import adjusted_rand_score
    import ssg_b200
    from ssg_b200 import cycle, synth, _lib
    import reid.evaluators as E
    from oracle import resnet_oracle as R
    g = np.load(golden_path)
    n, ns, S, lam = int(g["n"]), int(g["ns"]), int(g["num_split"]), float(g["lam"])
    banks = S + 1
    model = synth.build_model(S, int(g["weight_seed"]))
    feats = {}
    for tag, cnt, seed in (("tgt", n, int(g["seed_tgt"])), ("src", ns, int(g["seed_src"]))):
        imgs, _ = R.synth_identity_images(cnt, seed, int(g["per_identity"]), float(g["noise"]))
        names = ["%s%05d" % (tag, i) for i in range(cnt)]
        loader = [(imgs[i:i + batch], names[i:i + batch], [0] * len(names[i:i + batch]), [0] * len(names[i:i + batch]))
                  for i in range(0, cnt, batch)]
        f, _ = E.extract_features(model, loader, print_freq=10 ** 9, for_eval=False)      # a1: dict of CPU tensors
        assert isinstance(f[names[0]], list) and len(f[names[0]]) == banks and not f[names[0]][0].is_cuda
        feats[tag] = _restack(f, names, banks)                                             # a5
    out = {"n": n, "ns": ns, "banks": banks,
           "reference_under_4e-3_feature_noise": {
               "rank_entry_mismatch": [float(x) for x in g["noise_rank_entry_mismatch"]],
               "rank_set_mismatch_rows": [float(x) for x in g["noise_rank_set_mismatch_rows"]],
               "ari": [[float(x) for x in g["noise_ari_r%d" % ri]] for ri in range(len(g["rhos"]))]},
           "d2_nearest_neighbour_median": [float(x) for x in g["d2_nn"]],
           "d2_median": [float(x) for x in g["d2_median"]]}
    head = g["feat_tgt_head"]
    k = head.shape[1]
    out["feature_rel_err_max"] = max(
        float(np.linalg.norm(feats["tgt"][b][i].numpy() - head[b, i]) / np.linalg.norm(head[b, i]))
        for b in range(banks) for i in range(k))
    with contextlib.redirect_stdout(io.StringIO()):
        _, r_dist = cycle.compute_dist(feats["src"], feats["tgt"], lambda_value=lam, no_rerank=False, num_split=S)  # a6
    iu = np.triu_indices(n)
    out["rank_entry_mismatch"], out["rank_set_mismatch_rows"] = [], []
    out["final_abs_err_max"], out["final_abs_err_p99"], out["final_within_1e-4"] = [], [], []
    out["final_within_1_fp16_ulp_of_fp16_ref"] = []
    for b in range(banks):
        # rank table of this bank (the plan holds the tables of the last run: re-run the bank alone)
        plan = ssg_b200.rerank.get_plan(n, ns, 2048)
        plan.run(feats["src"][b].cuda(), feats["tgt"][b].cuda(), 20, 6, lam, _lib.DIST_EXACT)
        torch.cuda.synchronize()
        rank = plan.stage(_lib.STAGE_RANK, n)[:, :21]
        ref_rank = g["rank21_b%d" % b].astype(np.int32)
        out["rank_entry_mismatch"].append(float((rank != ref_rank).mean()))
        out["rank_set_mismatch_rows"].append(float(np.mean([set(a) != set(c) for a, c in zip(rank, ref_rank)])))
        f_gpu = r_dist[b].cpu().numpy()[iu]
        err = np.abs(f_gpu - g["final_f32_b%d" % b].astype(np.float64))
        out["final_abs_err_max"].append(float(err.max()))
        out["final_abs_err_p99"].append(float(np.percentile(err, 99)))
        out["final_within_1e-4"].append(float((err <= 1e-4).mean()))
        ref16 = g["final_ref_b%d" % b].astype(np.float64)
        ulp = np.spacing(np.maximum(ref16, 6e-5).astype(np.float16)).astype(np.float64)
        out["final_within_1_fp16_ulp_of_fp16_ref"].append(float((np.abs(f_gpu - ref16) <= ulp).mean()))
    rhos = [float(r) for r in g["rhos"]]
    out["rho"] = rhos
    for key in ("ari_vs_f32", "exact_label_fraction_vs_f32", "ari_vs_fp16_ref", "ari_f32_vs_fp16_ref", "eps_rel_err",
                "clusters", "clusters_ref"):
        out[key] = []
    for ri, rho in enumerate(rhos):
        args = types.SimpleNamespace(no_rerank=False, rho=rho)
        with contextlib.redirect_stdout(io.StringIO()):
            labels, clusters = cycle.generate_selflabel([[]] * banks, r_dist, 0, args, [])                     # a12
        row = {k2: [] for k2 in ("a", "x", "h", "c", "e", "n", "m")}
        for b in range(banks):
            want, want16 = g["labels_f32_r%d_b%d" % (ri, b)], g["labels_ref_r%d_b%d" % (ri, b)]
            got = np.asarray(labels[b])
            assert got.dtype == np.int64 and got.shape == (n,)
            row["a"].append(float(adjusted_rand_score(want, got)))
            row["x"].append(float((want == got).mean()))
            row["h"].append(float(adjusted_rand_score(want16, got)))
            row["c"].append(float(adjusted_rand_score(want16, want)))
            eps_ref = float(g["eps_f32_r%d_b%d" % (ri, b)])
            row["e"].append(abs(float(clusters[b].eps) - eps_ref) / eps_ref)
            row["n"].append(int(got.max()) + 1)
            row["m"].append(int(want.max()) + 1)
        for key, kk in (("ari_vs_f32", "a"), ("exact_label_fraction_vs_f32", "x"), ("ari_vs_fp16_ref", "h"),
                        ("ari_f32_vs_fp16_ref", "c"), ("eps_rel_err", "e"), ("clusters", "n"), ("clusters_ref", "m")):
            out[key].append(row[kk])
    return out


def test_images_to_labels_against_the_unmodified_reference(golden_dir):
    m = whole_path_metrics(os.path.join(golden_dir, "cycle_n512_S2.npz"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "whole_path_parity.json"), "w") as f:
        json.dump(m, f, indent=1)
    print(json.dumps(m))
    g = np.load(os.path.join(golden_dir, "cycle_n512_S2.npz"))
    assert m["feature_rel_err_max"] <= 8e-3, m["feature_rel_err_max"]
    for b in range(m["banks"]):
        assert m["rank_entry_mismatch"][b] <= float(g["noise_rank_entry_mismatch"][b]) + 0.10, m["rank_entry_mismatch"]
        # (a statistic that saturates near 1: measured 0.94-0.95 against 0.85-0.94 for the Gaussian perturbation)
        assert m["rank_set_mismatch_rows"][b] <= float(g["noise_rank_set_mismatch_rows"][b]) + 0.15
    for ri in range(len(m["rho"])):
        assert max(m["eps_rel_err"][ri]) <= 3e-2, m["eps_rel_err"]
        for b in range(m["banks"]):
            assert m["ari_vs_f32"][ri][b] >= float(g["noise_ari_r%d" % ri][b]) - 0.10, (ri, b, m["ari_vs_f32"])
    wc = m["rho"].index(5e-2)
    assert max(m["eps_rel_err"][wc]) <= 1e-2, m["eps_rel_err"]
    assert min(m["ari_vs_f32"][wc]) >= 0.99, m["ari_vs_f32"]
