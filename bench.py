#!/usr/bin/env python
"""bench.py — pseudo-label-cycle throughput on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      (the reference's CPU path on this box's host cores)

A step is one pass of the hot path over one batch of synthetic input: the pseudo-label cycle of
selftraining.py:189-222 at Market-1501 shape (BASELINE.json configs[1]): N = 16 702 target images and as
many source images, num_split = 2 -> 3 feature banks of 2048-d; per bank re-ranking (k1=20, k2=6,
lambda=0.1), eps at rho=1.6e-3 and DBSCAN(min_samples=4).  `value` is the throughput of the WHOLE cycle in Mpairs/s =
banks*N^2 ordered pairs taken from images to labels per second (value * ms_per_step = the pairs of one step), with the
images resident in HBM; `e2e` is the same metric through the host-buffer API (pinned host images in, host labels
out, copies inside the timed region).  The two halves of the BASELINE metric are reported beside it: "embed"
(images/s of the embedding stage) and "rerank" (Mpairs/s of the re-rank + eps + DBSCAN stage alone).

Further keys (round 2): `parity_gate` (a small oracle comparison that must pass before anything is printed;
BASELINE.md 3.6) and `result.whole_path_parity` (images -> labels against the unmodified reference, last B200
measurement); `e2e_u8` (the e2e leg with raw uint8 pixels, normalised on the device: a quarter of the H2D bytes);
`e2e_reference_api` (one cycle through the reference-SHAPED API as the unmodified driver calls it: dict of per-image
CPU tensors -> re-stack -> numpy N x N matrices -> labels; N=1 only); `result.labels_sha1` (labels of the last timed
step, for byte-comparison across GPU counts); `finetune_step` with --finetune-step.  --impl reference times the
reference's CPU path at three sizes and extrapolates with the fit of BASELINE.md 3.3.  Other BASELINE configs:
--banks 4 (configs[2]), --n 36411 --features-only --rho R (configs[3]), --n 126441 --shard-finish (configs[4]).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "self-similarity-grouping_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "pseudo-label cycle: images/s embed + Mpairs/s re-rank+DBSCAN @ N=16702"
UNIT = "Mpairs/s (banks*N^2 ordered pairs per second through the whole cycle: embed both sets + re-rank + eps + DBSCAN)"
UNIT_STAGE = "Mpairs/s (banks*N^2 ordered pairs re-ranked + eps + DBSCAN-labelled per second, that stage alone)"


def cycle_mpairs(n, banks, img_per_s, stage_mpairs):
    """Whole-cycle throughput implied by the two stage rates: 2N images embedded, then banks*N^2 pairs."""
    pairs = float(banks) * n * n
    return pairs / (2.0 * n / img_per_s + pairs / (stage_mpairs * 1e6)) / 1e6
D = 2048
LAMBDA, RHO, K1, K2, MIN_SAMPLES = 0.1, 1.6e-3, 20, 6, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--num-images", dest="n", type=int, default=16702,
                    help="target (= source) set size; use --num-images under torchrun (its parser rejects --n as ambiguous)")
    ap.add_argument("--banks", type=int, default=3)
    ap.add_argument("--dist-mode", default="tensor", choices=["tensor", "exact"])
    ap.add_argument("--cpu-sample", type=int, default=2560, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-embed-sample", type=int, default=512, help="images of the bounded CPU embedding sample")
    ap.add_argument("--batch", type=int, default=512, help="images per embedding batch")
    ap.add_argument("--features-only", action="store_true", help="skip the embedding stage (synthetic features)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", action="store_true",
                    help="N>1: independent replicas (weak scaling) instead of sharding one cycle over the GPUs")
    ap.add_argument("--shard-finish", action="store_true",
                    help="N>1 sharded mode: row-sharded finish (every rank holds only its rows of final_dist; "
                         "distributed eps and DBSCAN) instead of the bank-parallel finish")
    ap.add_argument("--sparse-finish", action="store_true",
                    help="never materialise final_dist (CSR over the touched pairs, certified "
                         "eps + DBSCAN on it; DESIGN.md 3.6) instead of the dense N x N float64 matrix")
    ap.add_argument("--rho", type=float, default=RHO, help="fraction of the smallest pair distances averaged into eps")
    ap.add_argument("--finetune-step", action="store_true",
                    help="after the cycle, time one FinedTrainer2 step on the pseudo-labels (rank 0; BASELINE configs[4])")
    ap.add_argument("--no-reference-api", action="store_true",
                    help="skip the e2e leg through the reference-shaped API (dict of CPU tensors -> numpy N x N -> labels)")
    ap.add_argument("--no-u8", action="store_true", help="skip the uint8-pixel e2e leg (ssg_embed_forward_u8)")
    ap.add_argument("--quick", action="store_true",
                    help="profiling runs (ncu): exactly --warmup warm-up steps, no e2e leg, no CPU baseline")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- inputs
def synth_bank(n, d, seed, noise, device, per_cluster=20):
    """Feature-level generator of SURVEY.md §8d on the device: n/20 Gaussian centres + noise, L2-normalised."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    c = max(n // per_cluster, 1)
    centres = torch.randn(c, d, generator=g, device=device)
    lab = torch.randint(0, c, (n,), generator=g, device=device)
    f = centres[lab] + noise * torch.randn(n, d, generator=g, device=device)
    return (f / f.norm(dim=1, keepdim=True)).contiguous()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_cycle_sample(n_sample, d=D, seed=0, mode="ref"):
    """The reference's CPU path (oracle port; the unmodified reference when /root/reference exists) on a
    bounded sample: one bank, n_sample target and source rows.  Returns (seconds, kind)."""
    import numpy as np
    from oracle import ssg_oracle as O, refshim
    from sklearn.cluster import DBSCAN
    tgt, _ = O.synth_features(n_sample, d, seed)
    src, _ = O.synth_features(n_sample, d, seed + 1, noise=0.6)
    kind = "reference" if refshim.available() else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        _, final = refshim.ref_re_ranking(src, tgt, mode=mode, lambda_value=LAMBDA)
    else:
        _, final = O.re_ranking(src, tgt, k1=K1, k2=K2, lambda_value=LAMBDA, mode=mode)
    eps = O.eps_estimate(final, RHO * 10)     # a small sample needs a larger rho for a non-empty slice
    DBSCAN(eps=eps, min_samples=MIN_SAMPLES, metric="precomputed", n_jobs=8).fit_predict(final)
    return time.perf_counter() - t0, kind


def cpu_embed_sample(n_img, seed=1234, budget_s=10.0):
    """reid/evaluators.py:18-60 on the host cores (torch CPU fp32, all threads), num_split=2: batches of 32 images
    until `n_img` images are done or `budget_s` seconds have passed (at least one batch).  -> (seconds, images)."""
    import torch
    from oracle import resnet_oracle as R
    torch.set_num_threads(os.cpu_count() or 1)
    model = R.build_model(2, 0)
    imgs = R.synth_images(min(n_img, 64), seed)
    done, t0 = 0, time.perf_counter()
    while done < n_img:
        n = min(32, n_img - done)
        part = imgs[(done % 64):(done % 64) + n]
        names = ["i%d" % (done + i) for i in range(part.shape[0])]
        R.extract_features(model, [(part, names, [0] * len(names), [0] * len(names))], for_eval=False)
        done += part.shape[0]
        if time.perf_counter() - t0 >= budget_s:
            break
    return time.perf_counter() - t0, done


def fit_stage_time(points):
    """Least-squares fit t(N) = a*N^2 + c*N through the measured (N, seconds) points of the re-rank + eps + DBSCAN
    stage (t = a*N^2*d + b*N^2 + c*N; d is fixed at 2048 here, so the two N^2 terms are one
    coefficient).  Returns (a, c); c is clamped at 0 (a negative linear term would be noise)."""
    import numpy as np
    n = np.array([p[0] for p in points], dtype=np.float64)
    t = np.array([p[1] for p in points], dtype=np.float64)
    if len(set(n.tolist())) < 2:
        return float((t / (n * n)).mean()), 0.0
    A = np.stack([n * n, n], 1)
    (a, c), *_ = np.linalg.lstsq(A, t, rcond=None)
    if c < 0 or a <= 0:
        a, c = float((t * n * n).sum() / (n ** 4).sum()), 0.0
    return float(a), float(c)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # Every step is a bounded sample.  The O(N^2) stage is timed at three sizes n0, 2*n0, 4*n0 (one size per step, in
    # turn) and extrapolated to the full N with the fit of BASELINE.md 3.3; n0 is sized from a 512-row probe (also the
    # warm-up) so that the whole run stays within ~4 minutes whatever --steps is (n0 = 1024 -> {1024, 2048, 4096}, the
    # sizes BASELINE.md names, when the budget allows: e.g. --steps 3).
    budget = 240.0
    embed_budget = min(10.0, 0.25 * budget / max(args.steps, 1))
    t_probe, _ = cpu_cycle_sample(512)
    c_probe = max(t_probe, 1e-3) / 512.0 ** 2
    cycles = max(args.steps, 1) / 3.0
    n0 = int(((budget - embed_budget * args.steps) / (21.0 * cycles * c_probe)) ** 0.5) // 64 * 64
    n0 = max(256, min(1024, n0))
    sizes = [n0, 2 * n0, 4 * n0]
    points, etimes, eimgs_all = [(512, t_probe)], [], []
    step_s = []
    kind = "port"
    for i in range(args.steps):
        n_s = sizes[i % 3]
        t, kind = cpu_cycle_sample(n_s)
        esec, eimgs = cpu_embed_sample(args.cpu_embed_sample, budget_s=embed_budget)
        points.append((n_s, t))
        etimes.append(esec)
        eimgs_all.append(eimgs)
        step_s.append(t + esec)
    a, c = fit_stage_time(points)
    t_full = a * args.n ** 2 + c * args.n                      # one bank at the full size, extrapolated
    stage_val = args.n ** 2 / t_full / 1e6
    measured = {}
    for n_s, t in points:
        measured.setdefault(n_s, []).append(t)
    measured_rows = ", ".join("N=%d: %.2f s (%.3f Mpairs/s)" % (k, sum(v) / len(v), k * k / (sum(v) / len(v)) / 1e6)
                              for k, v in sorted(measured.items()))
    img_rate = sum(eimgs_all) / sum(etimes)
    eimgs = eimgs_all[-1]
    # the same metric as the GPU arm: whole-cycle Mpairs/s at the full workload, from the two measured stage rates
    val = cycle_mpairs(args.n, args.banks, img_rate, stage_val)
    sample = ("re-rank + eps + DBSCAN of ONE bank (fp16 reference arithmetic, Ns=N, d=%d) measured at %s; fit "
              "t(N) = %.3e*N^2 + %.3e*N -> EXTRAPOLATED to N=%d: %.0f s per bank = %.3f Mpairs/s; embedding: %d images "
              "per step through the torch CPU ResNet-50 x2 passes (%.1f images/s); value = banks*N^2 / (2N / "
              "images_per_s + banks*N^2 / stage_pairs_per_s) at N=%d, banks=%d"
              % (D, measured_rows, a, c, args.n, t_full, stage_val, eimgs, img_rate, args.n, args.banks))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sum(step_s) / len(step_s) * 1e3, "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None, "dtype": "f16/f64 (numpy/scipy/sklearn)", "data": "synthetic",
        "config": {"workload": "configs[1]: N=Ns=%d synthetic, pseudo-label cycle (re-rank k1=20 k2=6 lambda=0.1 -> eps "
                               "rho=1.6e-3 -> DBSCAN min_samples=4); each step is a bounded sample of it: %s"
                               % (args.n, sample),
                   "value_is": "whole-cycle Mpairs/s on the host cores, EXTRAPOLATED from the measured small-N stage "
                               "times (three-size fit, BASELINE.md 3.3) and the measured embedding rate; the measured "
                               "rows are in cpu_baseline.sample; stage rates under 'rerank' and 'embed'"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "fit": {"a_s_per_pair": a, "c_s_per_row": c, "points": [[int(n_), float(t_)] for n_, t_ in points]},
                         "note": "cdist/numpy loops are single-threaded; sklearn DBSCAN uses n_jobs=8; torch CPU conv uses "
                                 "all %d host threads" % (os.cpu_count() or 1)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rerank": {"value": stage_val, "unit": UNIT_STAGE, "ms_per_step": t_full * args.banks * 1e3,
                   "extrapolated": True},
        "embed": {"value": img_rate, "unit": "images/s (two forward passes per image)",
                  "cores": os.cpu_count(), "sample": "%d images, torch CPU fp32, all host threads" % eimgs},
        "gpu_launches": 0,
    }
    print(json.dumps(line))



# ------------------------------------------------------------------------------------------- parity gate
def parity_gate(dev_index):
    """no number is reported unless the CUDA path reproduces the oracle on a small case.
    (1) re-rank -> eps -> DBSCAN on 384 synthetic feature rows against the oracle restatement (final_dist <= 1e-4,
    eps to 1e-12, labels bit-exact); (2) the embedding trunk on the reference's 4 golden images (relative error of
    every bank <= 8e-3).  The images -> labels comparison against the unmodified reference (512 + 384 images) is
    tests/test_gpu_whole_path.py; its last measured numbers are copied from profiles/ into `result`."""
    import numpy as np
    import torch
    import ssg_b200
    from oracle import ssg_oracle as O, resnet_oracle as R
    n, ns, d, lam, rho = 384, 256, 256, 0.1, 1.6e-2
    tgt, _ = O.synth_features(n, d, 0)
    src, _ = O.synth_features(ns, d, 1, noise=0.6)
    dev = torch.device("cuda", dev_index)
    _, f = ssg_b200.re_ranking_device(torch.from_numpy(src).to(dev), torch.from_numpy(tgt).to(dev), lambda_value=lam,
                                      dist_mode=1)
    plan = ssg_b200.ClusterPlan(n, 0, dev_index)
    eps, _ = plan.eps(f, rho)
    labels, _ = plan.dbscan(f, eps, 4)
    fh = f.cpu().numpy()
    _, f_ref = O.re_ranking(src, tgt, lambda_value=lam, mode="f32")
    err = float(np.abs(fh - f_ref).max())
    eps_ref = O.eps_estimate(fh, rho)
    labels_ok = bool(np.array_equal(labels.cpu().numpy(), O.dbscan_dfs(fh, eps, 4)))
    g = np.load(os.path.join(ROOT, "tests", "golden", "embed_4img.npz"))
    imgs = R.synth_images(int(g["n_img"]), int(g["seed_img"]))
    model = R.build_model(2, int(g["weight_seed"]))
    names = ["im%03d" % i for i in range(imgs.shape[0])]
    feats, _ = ssg_b200.extract_features(model, [(imgs, names, [0] * len(names), [0] * len(names))], for_eval=False)
    rel = max(float((feats[k][b] - torch.from_numpy(g["list_S2"][b, i])).norm()) for i, k in enumerate(names)
              for b in range(3))
    gate = {"final_dist_max_abs_err": err, "final_dist_tol": 1e-4, "eps_rel_err": abs(eps - eps_ref) / eps_ref,
            "labels_bit_exact": labels_ok, "embed_rel_err": rel, "embed_tol": 8e-3}
    gate["ok"] = bool(err <= 1e-4 and gate["eps_rel_err"] <= 1e-12 and labels_ok and rel <= 8e-3)
    return gate


def reference_api_cycle(model, host_tgt, host_src, num_split, batch):
    """One cycle through the REFERENCE-SHAPED API, as the unmodified driver calls it (selftraining.py:189-222):
    reid.evaluators.extract_features over a loader of host batches -> OrderedDict of per-image CPU tensors -> the
    driver's bank re-stacking -> reid.rerank.re_ranking (numpy in, N x N float64 numpy out) -> eps
    (selftraining.py:289-293 through reid.cluster.eps_estimate) -> reid.rerank.DBSCAN.fit_predict(host matrix).
    Returns (labels per bank, seconds per part)."""
    import contextlib
    import io
    import time
    import torch
    import reid.evaluators as E
    import reid.rerank as RR
    from reid.cluster import eps_estimate
    parts = {}
    banks = num_split + 1 if num_split > 1 else 1
    feats = {}
    t0 = time.perf_counter()
    for tag, host in (("src", host_src), ("tgt", host_tgt)):
        cnt = host.shape[0]
        names = ["%s%06d" % (tag, i) for i in range(cnt)]
        loader = [(host[i:i + batch], names[i:i + batch], [0] * len(names[i:i + batch]), [0] * len(names[i:i + batch]))
                  for i in range(0, cnt, batch)]
        f, _ = E.extract_features(model, loader, print_freq=10 ** 9, for_eval=False)
        feats[tag] = [torch.cat([f[nm][i].unsqueeze(0) for nm in names], 0) for i in range(banks)]   # selftraining.py:197-209
    parts["extract_features_and_restack_s"] = time.perf_counter() - t0
    labels = []
    t_rr = t_cl = 0.0
    for b in range(banks):
        t1 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            _, final = RR.re_ranking(feats["src"][b].numpy(), feats["tgt"][b].numpy(), lambda_value=LAMBDA)
        t2 = time.perf_counter()
        eps = eps_estimate(final, RHO)
        labels.append(RR.DBSCAN(eps=eps, min_samples=MIN_SAMPLES, metric="precomputed", n_jobs=8).fit_predict(final))
        parts.setdefault("eps", []).append(float(eps))
        t3 = time.perf_counter()
        t_rr += t2 - t1
        t_cl += t3 - t2
        del final
    parts["re_ranking_s"], parts["eps_dbscan_s"] = t_rr, t_cl
    parts["total_s"] = time.perf_counter() - t0
    return labels, parts


# ------------------------------------------------------------------------------------------- GPU arm
ALGO = {   # per launch: (bound, algorithmic work as a function of (rows, cols, d))
    "gemm_dist_tc": ("tensor", lambda r, c, d: 2.0 * d * r * c),        # flops: 2*d per ordered pair
    "sqdist_exact": ("fp64", lambda r, c, d: 3.0 * d * r * c),
    "jaccard_final": ("hbm", lambda r, c, d: 8.0 * r * c),              # float64 final_dist written once
    "eps_hist": ("hbm", lambda r, c, d: 8.0 * r * c / 2),               # upper triangle read once per pass
    "eps_sum": ("hbm", lambda r, c, d: 8.0 * r * c / 2),
    "dbscan_count": ("hbm", lambda r, c, d: 8.0 * r * c),
    "dbscan_fill": ("hbm", lambda r, c, d: 8.0 * r * c),
    "row_select": ("hbm", lambda r, c, d: 4.0 * r * c),                 # fp32 distance block read once
    "row_minmax": ("hbm", lambda r, c, d: 4.0 * r * c),
}


def workload_name(n, banks, world):
    """Which BASELINE.json config a run corresponds to (by shape)."""
    if n == 126441:
        return "configs[4] (MSMT17 shape)"
    if n == 36411:
        return "configs[3] (DukeMTMC shape)"
    if n == 16702 and banks == 4:
        return "configs[2] (num_split=3)"
    if n == 16702 and banks == 3:
        return "configs[1]"
    return "custom"


def finetune_step(model, images, labels, keep, dev, P=16, K=4, row0=0):
    """One FinedTrainer2 step on pseudo-labels (selftraining.py:149-161, 239-253; trainers.py:204-271): P identities x
    K images drawn from the kept images, global + per-bank triplet losses (own CUDA kernels, csrc/triplet.cu), model
    forward / backward through torch autograd (cuDNN convolutions: library code), SGD.  -> dict with the device time of
    that step ("ms": the reference's stock path) and of the same step with every convolution on this repo's tcgen05 kernels
    (ssg_b200.train: "own_convs", "own_convs_bf16_activations", and "own_convs_bf16_cuda_graph" = the whole step replayed
    as one CUDA graph; "cudnn_cuda_graph" for comparison)."""
    import numpy as np
    import torch
    from reid.loss import TripletLoss
    from reid.trainers import FinedTrainer2
    # `images` holds the rows [row0, row0 + len(images)) of the set (this rank's shard in a sharded run): the batch is
    # drawn from those rows only
    lab0 = np.asarray(labels[0])
    mine = np.zeros(lab0.shape[0], dtype=bool)
    mine[row0:row0 + images.shape[0]] = True
    ok = keep & mine
    ids = [c for c in np.unique(lab0[ok]) if c >= 0 and (lab0[ok] == c).sum() >= K][:P]
    if len(ids) < 2:
        return {"skipped": "fewer than two pseudo-identities with %d kept images on this rank" % K}
    idx = np.concatenate([np.flatnonzero((lab0 == c) & ok)[:K] for c in ids])
    imgs = images[torch.from_numpy(idx - row0).to(images.device)].to(dev).float()
    pids = [torch.from_numpy(np.asarray(l)[idx].astype(np.int64)) for l in labels]
    model = model.to(dev)
    crit = [TripletLoss(0.5, K, True).to(dev), TripletLoss(0.5, K, True).to(dev)]
    trainer = FinedTrainer2(model, crit)
    opt = torch.optim.SGD(model.parameters(), lr=6e-5, momentum=0.9, weight_decay=5e-4, nesterov=True)
    model.train()
    times = []
    for it in range(3):                                          # 2 warm-up steps, the third is reported
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        inputs, p, _ = trainer._parse_data((imgs, None, pids, [0] * len(idx)))
        loss, prec = trainer._forward(inputs, p, 0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    out = {"ms": times[-1], "batch": int(len(idx)), "identities": int(len(ids)), "instances": K, "loss": float(loss.item()),
           "note": "FinedTrainer2 step: model forward/backward = torch autograd over cuDNN (library code); the global + "
                   "per-bank triplet losses and their gradients are this repo's kernels (csrc/triplet.cu)"}
    # the same step with every convolution (forward, data gradient, weight gradient) on this repo's tcgen05 kernels
    # (ssg_b200.train.own_convs -> csrc/train.cu); BatchNorm / ReLU / pooling / SGD stay with torch
    try:
        from ssg_b200 import train as own
        times = []
        with own.own_convs(model) as swapped:
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                inputs, p, _ = trainer._parse_data((imgs, None, pids, [0] * len(idx)))
                loss, prec = trainer._forward(inputs, p, 0)
                opt.zero_grad()
                loss.backward()
                opt.step()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
        out["own_convs"] = {"ms": times[-1], "convolutions": int(swapped), "loss": float(loss.item()),
                            "note": "forward + dgrad + wgrad of all convolutions on the repo's tcgen05 GEMM kernels behind "
                                    "torch.autograd.Function (NCHW fp32 <-> NHWC bf16 conversions per layer included)"}
        # ... and with the activations between the convolutions kept in bf16 channels-last (no conversions; BatchNorm /
        # ReLU / adds on bf16 as under autocast; parameters, BatchNorm statistics and weight gradients fp32)
        times = []
        with own.own_convs(model, activations="bf16", cast_back=model.base.layer4) as swapped:
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                inputs, p, _ = trainer._parse_data((imgs, None, pids, [0] * len(idx)))
                loss, prec = trainer._forward(inputs, p, 0)
                opt.zero_grad()
                loss.backward()
                opt.step()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
        out["own_convs_bf16_activations"] = {
            "ms": times[-1], "convolutions": int(swapped), "loss": float(loss.item()),
            "note": "as own_convs, the tensors between the convolutions stay bf16 channels-last (zero-copy in and out of "
                    "the kernels); trunk output cast back to fp32 in front of the heads"}
        # ... and the whole step (forward, losses, backward, SGD) captured once in a CUDA graph and replayed: with the own
        # convolutions + bf16 activations, and -- for reference -- with the cuDNN convolutions
        def graphed(tag, ctx):
            try:
                for c in crit:
                    c.check = False                                # the host-side no-negative check would break the capture
                pids_dev = [q.to(dev) for q in pids]

                def step_fn(im, *pp):
                    return trainer._forward([im], list(pp), 0)[0]
                with ctx:
                    gs = own.GraphedStep(step_fn, opt, [imgs] + pids_dev)
                ts = []
                for it in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    lg = gs(imgs, *pids_dev)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                out[tag] = {"ms": ts[-1], "loss": float(lg.item())}
            except Exception as exc:
                out[tag] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
            finally:
                for c in crit:
                    c.check = True
        import contextlib
        import gc
        # the autograd graphs of the eager legs must be gone: their AccumulateGrad nodes are bound to the default stream,
        # and a capture that reaches one of them is invalidated (torch re-creates the nodes on the warm-up stream)
        loss = prec = inputs = p = None
        gc.collect()
        graphed("own_convs_bf16_cuda_graph", own.own_convs(model, activations="bf16", cast_back=model.base.layer4))
        graphed("cudnn_cuda_graph", contextlib.nullcontext())
    except Exception as exc:                                       # reported, never hidden
        out["own_convs"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    model.eval()
    return out


def load_whole_path_parity():
    """Last measured images -> labels parity against the unmodified reference (tests/test_gpu_whole_path.py on a B200)."""
    path = os.path.join(ROOT, "profiles", "r02_whole_path_parity.json")
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        m = json.load(f)
    keep = ("feature_rel_err_max", "rank_entry_mismatch", "rank_set_mismatch_rows", "final_within_1e-4", "rho", "ari_vs_f32",
            "exact_label_fraction_vs_f32", "ari_vs_fp16_ref", "ari_f32_vs_fp16_ref", "eps_rel_err",
            "reference_under_4e-3_feature_noise")
    out = {kk: m.get(kk) for kk in keep}
    out["source"] = "profiles/r02_whole_path_parity.json (N=512 + 384 synthetic identity images, see the test's docstring)"
    return out


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": p.get("hbm_gbs", 6650.0), "tensor": p.get("bf16_tflops_sustained", 1400.0),
                "source": "measured (MEASURED_PEAKS.json; bf16 sustained)"}
    return {"hbm": 6650.0, "tensor": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import ssg_b200
    from ssg_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = _lib.DIST_TENSOR if args.dist_mode == "tensor" else _lib.DIST_EXACT
    n, banks = args.n, args.banks
    num_split = banks - 1 if banks > 1 else 1
    with_embed = not args.features_only
    pairs_per_step = float(banks) * n * n

    # N>1 default: ONE cycle sharded over the ranks (strong scaling): image shards -> all-gather of the feature banks
    # (NCCL) -> row-block distance stage + table gather -> bank-parallel finish (ssg_b200.dist).  --replicas: every rank
    # owns an independent set of the same shape (weak scaling, no data-path collective).
    sharded = world > 1 and not args.replicas
    if with_embed:
        model = synth.build_model(num_split, 0)
        if sharded:
            from ssg_b200 import dist as sdist
            comm = sdist.Comm()
            backend = sdist.CudaBackend(local, mode, args.batch)
            lo, hi = sdist.shard_bounds(n, world, rank)
            tgt_full, _ = synth.synth_images(n, 1234, dev)          # same seeds on every rank: one global data set
            tgt_img = tgt_full[lo:hi].clone()
            del tgt_full
            src_full, _ = synth.synth_images(n, 4321, dev)
            src_img = src_full[lo:hi].clone()
            del src_full
            torch.cuda.empty_cache()
        else:
            tgt_img, _ = synth.synth_images(n, 1234 + 100 * rank, dev)
            src_img, _ = synth.synth_images(n, 4321 + 100 * rank, dev)
        plan_e = ssg_b200.embed.get_plan(args.batch, local)
        plan_e.load_model(model)
    elif sharded:
        # one global synthetic feature set (same seeds on every rank); this rank keeps its row shard [banks, n_local, d]
        from ssg_b200 import dist as sdist
        comm = sdist.Comm()
        backend = sdist.CudaBackend(local, mode, args.batch)
        lo, hi = sdist.shard_bounds(n, world, rank)
        tgt_f = torch.stack([synth_bank(n, D, 10 + b, 0.5, dev)[lo:hi] for b in range(banks)]).contiguous()
        src_f = torch.stack([synth_bank(n, D, 20 + b, 0.6, dev)[lo:hi] for b in range(banks)]).contiguous()
    else:
        tgt_f = [synth_bank(n, D, 1000 * rank + 10 + b, 0.5, dev) for b in range(banks)]
        src_f = [synth_bank(n, D, 1000 * rank + 20 + b, 0.6, dev) for b in range(banks)]

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    stage_ms = {"embed": 0.0, "rerank": 0.0}

    def cycle(tgt_in, src_in, record):
        """selftraining.py:196-218: embed source + target sets, then re-rank / eps / DBSCAN per bank."""
        if record:
            ev[0].record()
        if with_embed and sharded:
            # embed this rank's shards straight into the gather buffers; the all-gather of the target banks overlaps the
            # embedding of the source shard (ssg_b200.dist.embed_and_gather)
            tf, sf = sdist.embed_and_gather(model, tgt_in, src_in, n, n, num_split, backend, comm)
        elif with_embed:
            tf = ssg_b200.embed_images(model, tgt_in, num_split, False, args.batch, local)
            sf = ssg_b200.embed_images(model, src_in, num_split, False, args.batch, local)
            tfl, sfl = [tf[b] for b in range(banks)], [sf[b] for b in range(banks)]
        else:
            tfl, sfl = tgt_in, src_in
            if sharded and not tfl.is_cuda:                      # e2e leg: this rank's feature shards from pinned host memory
                tfl, sfl = tfl.to(dev, non_blocking=True), sfl.to(dev, non_blocking=True)
        if record:
            ev[1].record()
        if sharded and with_embed:
            out = sdist.sharded_pseudo_label_cycle(model, None, None, n, n, num_split, LAMBDA, args.rho, backend=backend,
                                                   comm=comm, features_full=(tf, sf),
                                                   shard_finish=True if args.shard_finish else None,
                                                   sparse=True if args.sparse_finish else None)
        elif sharded:
            out = sdist.sharded_pseudo_label_cycle(None, None, None, n, n, num_split, LAMBDA, args.rho, backend=backend,
                                                   comm=comm, features=(tfl, sfl),
                                                   shard_finish=True if args.shard_finish else None,
                                                   sparse=True if args.sparse_finish else None)
        else:
            out = ssg_b200.pseudo_label_cycle(sfl, tfl, LAMBDA, args.rho, dist_mode=mode, device=local,
                                              sparse=True if args.sparse_finish else None)
        if record:
            ev[2].record()
            torch.cuda.synchronize()
            stage_ms["embed"] += ev[0].elapsed_time(ev[1])
            stage_ms["rerank"] += ev[1].elapsed_time(ev[2])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(tgt_in, src_in, steps, profile=False):
        stage_ms["embed"] = stage_ms["rerank"] = 0.0
        barrier()
        if profile:
            _lib.profile(reset=True)
            _lib.profile(on=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = cycle(tgt_in, src_in, True)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), stage_ms["embed"], stage_ms["rerank"]], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = {}
        if profile:
            prof = _lib.profile()
            _lib.profile(on=False)
        return [float(v) for v in t.tolist()], out, prof

    gate = None
    if not args.quick:
        gate = parity_gate(local)
        if not gate["ok"]:
            raise SystemExit("bench.py: parity gate FAILED, no number is reported: %s" % json.dumps(gate))
    dev_in = (tgt_img, src_img) if with_embed else (tgt_f, src_f)
    n_warm = args.warmup if args.quick else max(args.warmup, 3)
    for _ in range(n_warm):
        cycle(dev_in[0], dev_in[1], False)
    clocks = ClockSampler(local)
    (ms_dev, ms_embed, ms_rerank), out, prof = timed(dev_in[0], dev_in[1], args.steps, profile=True)
    clk = clocks.stop()

    # end to end through the host-buffer API: pinned host inputs, copies inside the timed region
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms_dev / args.steps, "embed_ms": ms_embed / args.steps,
                              "rerank_ms": ms_rerank / args.steps,
                              "kernels_ms_per_step": {kk: round(v[0] / args.steps, 4) for kk, v in prof.items()}}))
        if world > 1:
            dist.destroy_process_group()
        return
    if with_embed:
        host_in = (tgt_img.cpu().pin_memory(), src_img.cpu().pin_memory())
        h2d = (tgt_img.shape[0] + src_img.shape[0]) * 3 * 256 * 128 * 4      # per rank
    elif sharded:
        host_in = (tgt_f.cpu().pin_memory(), src_f.cpu().pin_memory())
        h2d = (tgt_f.numel() + src_f.numel()) * 4
    else:
        host_in = ([t.cpu().pin_memory() for t in tgt_f], [t.cpu().pin_memory() for t in src_f])
        h2d = 2 * banks * n * D * 4
    cycle(host_in[0], host_in[1], False)
    (ms_e2e, ms_e2e_embed, ms_e2e_rerank), out_e2e, _ = timed(host_in[0], host_in[1], args.steps)

    # e2e through raw uint8 pixels (ToTensor + Normalize on the device, a quarter of the H2D bytes)
    e2e_u8 = None
    if with_embed and not args.no_u8 and not sharded:
        def to_u8(img):
            return (img.permute(0, 2, 3, 1) * 50.0 + 128.0).clamp_(0, 255).to(torch.uint8).contiguous()
        u8_host = (to_u8(tgt_img).cpu().pin_memory(), to_u8(src_img).cpu().pin_memory())
        ku = max(1, min(args.steps, 5))
        cycle(u8_host[0], u8_host[1], False)
        (ms_u8, ms_u8_embed, _), out_u8, _ = timed(u8_host[0], u8_host[1], ku)
        e2e_u8 = {"value": pairs_per_step * world * ku / (ms_u8 / 1e3) / 1e6, "unit": UNIT, "ms_per_step": ms_u8 / ku,
                  "embed_ms_per_step": ms_u8_embed / ku, "steps": ku,
                  "h2d_bytes_per_step": 2 * n * 256 * 128 * 3 * world,
                  "api": "ssg_b200.embed_images(pinned host uint8 HWC pixels; normalised on the device) + "
                         "pseudo_label_cycle -> host labels", "clusters": [int(l.max()) + 1 for l in out_u8[0]]}
        del u8_host
    # e2e through the reference-shaped API (what the unmodified driver pays): N=1 only, one timed pass after one warm-up
    ref_api = None
    if with_embed and world == 1 and not args.no_reference_api:
        reference_api_cycle(model, host_in[0][:2048], host_in[1][:2048], num_split, args.batch)     # warm-up (plans)
        torch.cuda.synchronize()
        lab_api, parts = reference_api_cycle(model, host_in[0], host_in[1], num_split, args.batch)
        ref_api = {"value": pairs_per_step / parts["total_s"] / 1e6, "unit": UNIT, "ms_per_step": parts["total_s"] * 1e3,
                   "parts_s": {kk: round(v, 3) for kk, v in parts.items() if kk != "eps"}, "steps": 1,
                   "eps": parts["eps"], "eps_device_resident_path": [float(e) for e in out_e2e[1]],
                   "labels_differ_per_bank": [int((a != b).sum()) for a, b in zip(lab_api, out_e2e[0])],
                   "h2d_bytes_per_step": h2d + banks * (2 * n * D * 4 + n * n * 8),
                   "d2h_bytes_per_step": 2 * banks * n * D * 4 + banks * n * n * (8 + 4) + banks * n * 8,
                   "api": "reid.evaluators.extract_features(loader of host batches) -> OrderedDict of per-image CPU "
                          "tensors -> re-stack -> reid.rerank.re_ranking (numpy, N x N float64 + float32 to the host) -> "
                          "eps -> reid.rerank.DBSCAN.fit_predict(host matrix)",
                   "labels_equal_device_resident_path": bool(all((a == b).all() for a, b in zip(lab_api, out_e2e[0])))}

    labels, eps_list, keep = out
    k = args.steps
    # whole job: a sharded cycle copies every image once (the ranks' shards add up to the set), replicas copy one set each
    h2d_job = (2 * n * 3 * 256 * 128 * 4 if sharded else h2d * world) if with_embed else \
        (2 * banks * n * D * 4 if sharded else h2d * world)
    units = 1 if sharded else world          # sharded: all ranks together process ONE data set
    value = pairs_per_step * units * k / (ms_dev / 1e3) / 1e6             # whole cycle (== stage alone with --features-only)
    e2e_value = pairs_per_step * units * k / (ms_e2e / 1e3) / 1e6
    stage_value = pairs_per_step * units * k / (ms_rerank / 1e3) / 1e6
    e2e_stage_value = pairs_per_step * units * k / (ms_e2e_rerank / 1e3) / 1e6

    finetune = None
    if args.finetune_step and with_embed and rank == 0:
        finetune = finetune_step(model, tgt_img, labels, keep, dev, row0=lo if sharded else 0)
    line = None
    if rank == 0:
        peaks = load_peaks()
        launches = sum(v[1] for v in prof.values())
        kern_ms = {kk: v[0] for kk, v in prof.items()}
        top = max(kern_ms, key=kern_ms.get) if kern_ms else None
        roof = None
        conv_names = ("conv1x1_tc", "conv3x3_tc", "conv_stem_tc")
        if top in conv_names:
            # the three conv groups are one kernel template (gemm_kernel<BN, EpiConv>): rate them together
            conv_ms = sum(prof[c][0] for c in conv_names if c in prof)
            flops = 2.0 * 2669150208 * 2 * (2 * n) * k * (1.0 / world if sharded else 1.0)   # rank 0's share
            ach = flops / (conv_ms / 1e3) / 1e12
            # DRAM traffic of the same launches, from the committed ncu capture of one embedding batch (not re-measured
            # here: a number printed under a profiler is never a bench value, but the byte counts are clock independent)
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "r02_final_conv_traffic.json")
            if os.path.isfile(tpath):
                with open(tpath) as f:
                    tj = json.load(f)
                batches = 2.0 * n * (1.0 / world if sharded else 1.0) / tj["batch_images"]
                traffic = {"bytes_per_step": tj["dram_bytes_per_batch"] * batches, "source": "profiles/r02_final_conv_traffic.json",
                           "algorithmic_note": "see DESIGN.md 3.1b: what bounds the convolution tiles (shared-memory operand bandwidth "
                                               "of the UMMA, load latency; HBM in the layer-1/2 1x1 convolutions)"}
            conv_launches = sum(prof[c][1] for c in conv_names if c in prof)
            roof = {"kernel": "gemm_kernel<StagedEpi> (conv1x1_tc+conv3x3_tc+conv_stem_tc, %d launches)" % conv_launches,
                    "bound": "tensor", "achieved": ach, "peak": peaks["tensor"], "unit": "TFLOP/s",
                    "frac": ach / peaks["tensor"],
                    # DRAM bytes per launch, averaged over the template's launches of one step (ncu capture of one batch)
                    "traffic": (traffic["bytes_per_step"] * k / conv_launches) if traffic and conv_launches else None,
                    "traffic_detail": traffic, "flops_per_launch": flops / max(conv_launches, 1),
                    "ms_per_launch": conv_ms / max(conv_launches, 1), "ms_per_step": conv_ms / k,
                    "peak_source": peaks["source"],
                    "note": "algorithmic flops = 2 x 2 669 150 208 conv MACs per image-pass (SURVEY.md 8d), summed over "
                            "all conv launches of the step"}
        elif top in ALGO:
            bound, fn = ALGO[top]
            per_launch_ms = prof[top][0] / prof[top][1]
            work = fn(n, n, D)
            if bound == "tensor":
                ach = work / (per_launch_ms / 1e3) / 1e12
                roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peaks["tensor"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tensor"], "traffic": None, "ms_per_launch": per_launch_ms,
                        "peak_source": peaks["source"],
                        "note": "algorithmic flops 2*d*N^2 per launch; the bf16x3 split issues 3x that on the tensor pipe"}
            elif bound == "hbm":
                ach = work / (per_launch_ms / 1e3) / 1e9
                roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": ach / peaks["hbm"], "traffic": None, "ms_per_launch": per_launch_ms,
                        "peak_source": peaks["source"]}
            else:
                roof = {"kernel": top, "bound": bound, "achieved": work / (per_launch_ms / 1e3) / 1e12, "peak": None,
                        "unit": "Tflop64/s", "frac": None, "traffic": None, "ms_per_launch": per_launch_ms}
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only (the other ranks would idle)
            sec, kind = cpu_cycle_sample(args.cpu_sample)
            cpu = {"value": args.cpu_sample ** 2 / sec / 1e6, "unit": UNIT_STAGE, "cores": 1, "kind": kind,
                   "seconds": sec, "rerank_stage_value": args.cpu_sample ** 2 / sec / 1e6,
                   "sample": "1 bank, N=Ns=%d rows of the %d-row workload (re_ranking fp16 reference arithmetic + eps + "
                             "sklearn DBSCAN n_jobs=8); numpy/scipy parts are single-threaded" % (args.cpu_sample, n)}
            if with_embed:
                esec, eimgs = cpu_embed_sample(args.cpu_embed_sample)
                cpu["embed"] = {"value": eimgs / esec, "unit": "images/s", "cores": os.cpu_count(),
                                "sample": "%d images, torch CPU fp32 ResNet-50 x2 passes, all host threads" % eimgs}
                # the same metric as `value`: whole-cycle Mpairs/s implied by the two measured CPU stage rates
                cpu["value"] = cycle_mpairs(n, banks, eimgs / esec, cpu["rerank_stage_value"])
                cpu["unit"] = UNIT
                cpu["sample"] += "; value = banks*N^2 / (2N / images_per_s + banks*N^2 / stage_pairs_per_s) at the full size"
        embed = None
        if with_embed:
            img_per_s = 2.0 * n * units * k / (ms_embed / 1e3)
            flops = 2.0 * 2669150208 * 2 * (2 * n) * k * units / world   # per GPU
            embed = {"value": img_per_s, "unit": "images/s (source + target sets; two forward passes per image)",
                     "ms_per_step": ms_embed / k, "images_per_step": 2 * n,
                     "tensor_tflops_algorithmic": flops / (ms_embed / 1e3) / 1e12,
                     "frac_of_measured_bf16_peak": flops / (ms_embed / 1e3) / 1e12 / peaks["tensor"],
                     "e2e_value": 2.0 * n * units * k / (ms_e2e_embed / 1e3), "batch": args.batch}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": k,
            "warmup": n_warm, "ms_per_step": ms_dev / k, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": ("bf16 conv (fp32 accumulate) + " if with_embed else "") +
                     ("f32/f64 re-rank (bf16x3 tensor-core candidates, exact f64 re-score)"
                      if mode == _lib.DIST_TENSOR else "f32/f64 re-rank (exact f64 distances)"),
            "data": "synthetic",
            "config": {"workload": workload_name(n, banks, world) + ": N=Ns=%d synthetic 256x128 images, random-init ResNet-50 (num_split=%d -> %d "
                                   "banks x 2048-d): embed source+target sets (flip TTA) -> per bank re-rank k1=20 k2=6 "
                                   "lambda=0.1 -> eps rho=1.6e-3 -> DBSCAN min_samples=4%s"
                                   % (n, num_split, banks, "" if with_embed else "; EMBED SKIPPED (--features-only)"),
                       "l2": "inputs exceed the 126 MB L2 (%.1f GB of images, %.1f GB distance block per bank)"
                             % (2 * n * 3 * 256 * 128 * 4 / 1e9, n * n * 4 / 1e9),
                       "value_is": "whole-cycle Mpairs/s: banks*N^2 pairs / ms_per_step (images in HBM -> labels); the two "
                                   "stages of the BASELINE metric are under 'embed' (images/s) and 'rerank' (Mpairs/s of "
                                   "that stage alone)",
                       "parallelism": ("one cycle sharded over %d GPUs: image shards embedded straight into the gather buffer, "
                                       "in-place NCCL all-gather of the feature banks (the target set's overlaps the "
                                       "source set's embedding; counted in the embed stage), row-block distance stage, "
                                       "%s finish"
                                       % (world, "row-sharded" if (args.shard_finish or sdist._shard_finish_default())
                                          else "bank-parallel")) if sharded else
                                      "%d independent replicas (one target set per GPU)" % world,
                       "dist_mode": args.dist_mode,
                       "final_dist": "sparse (CSR over the touched pairs)" if args.sparse_finish else "dense float64 N x N"},
            "embed": embed,
            "rerank": {"value": stage_value, "unit": UNIT_STAGE, "ms_per_step": ms_rerank / k,
                       "e2e_value": e2e_stage_value},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / k, "rerank_ms_per_step": ms_e2e_rerank / k,
                    "embed_ms_per_step": ms_e2e_embed / k, "h2d_bytes_per_step": h2d_job, "d2h_bytes_per_step": banks * n * 8 * world,
                    "h2d_bytes_per_step_rank0": h2d,
                    "api": "ssg_b200.embed_images(pinned host images) + ssg_b200.pseudo_label_cycle -> host labels"},
            "e2e_u8": e2e_u8, "e2e_reference_api": ref_api,
            "gpu_launches": int(launches),
            "kernels_ms_per_step": {kk: round(v[0] / k, 4) for kk, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
            "parity_gate": gate, "finetune_step": finetune,
            "result": {"clusters": [int(l.max()) + 1 for l in labels], "eps": [round(e, 6) for e in eps_list],
                       "eps_exact": [float(e).hex() for e in eps_list],
                       "labels_sha1": hashlib.sha1(b"".join(np.ascontiguousarray(l, dtype=np.int64).tobytes()
                                                            for l in labels)).hexdigest(),
                       "rho": args.rho,
                       "whole_path_parity": load_whole_path_parity(),
                       "kept_images": int(keep.sum()),
                       "rows_recomputed_exactly_last_bank": int(ssg_b200.rerank.get_plan(n, n, D, local).stage(
                           _lib.STAGE_FLAGGED, n)[0]) if mode == _lib.DIST_TENSOR else None},
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
