"""Generate tests/golden/*.npz by running the UNMODIFIED reference (needs /root/reference; build
container only).  Usage: python -m oracle.make_goldens

The reference has no tests or golden vectors of its own; these files pin the oracle
restatement and the CUDA path to the reference's behaviour with numpy 2.3.5 / scipy 1.18.1 /
scikit-learn 1.9.0 / torchvision 0.26 (versions recorded inside each file).
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refshim, ssg_oracle as O, resnet_oracle as R, triplet_oracle as TO  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def versions():
    import scipy, sklearn, torch, torchvision
    return np.array("numpy %s scipy %s sklearn %s torch %s torchvision %s" % (
        np.__version__, scipy.__version__, sklearn.__version__, torch.__version__,
        torchvision.__version__))


def rerank_case(name, n, ns, d, seed, lam, rhos, per_cluster=20, noise=0.5):
    from sklearn.cluster import DBSCAN
    tgt, _ = O.synth_features(n, d, seed, per_cluster, noise)
    src, _ = O.synth_features(ns, d, seed + 100, per_cluster, noise * 1.2)
    e32, f32 = refshim.ref_re_ranking(src, tgt, mode="f32", lambda_value=lam)
    e16, f16 = refshim.ref_re_ranking(src, tgt, mode="ref", lambda_value=lam)
    st = {}
    O.re_ranking(src, tgt, lambda_value=lam, mode="f32", stages=st)
    out = dict(tgt=tgt, src=src, lam=lam, euclid_f32=e32, final_f32=f32, euclid_ref=e16, final_ref=f16,
               vec_f32=st["vec"], rank21_f32=st["rank"][:, :21], rhos=np.array(rhos), versions=versions())
    for bi, rho in enumerate(rhos):
        # selftraining.py:289-296,306 verbatim on the reference's own output
        tri = np.triu(f32, 1)
        tri = tri[np.nonzero(tri)]
        tri = np.sort(tri, axis=None)
        top = np.round(rho * tri.size).astype(int)
        eps = tri[:top].mean()
        lab = DBSCAN(eps=eps, min_samples=4, metric="precomputed", n_jobs=8).fit_predict(f32)
        out["eps_%d" % bi] = eps
        out["labels_%d" % bi] = lab
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "clusters", [int(out["labels_%d" % b].max()) + 1 for b in range(len(rhos))])


def rerank_init_case(name, q, g, d, seed):
    f, _ = O.synth_features(q + g, d, seed, per_cluster=10)
    qf, gf = f[:q], f[q:]
    q_g, q_q, g_g = qf @ gf.T, qf @ qf.T, gf @ gf.T
    out = refshim.ref_re_ranking_init(q_g, q_q, g_g, stable=True)
    np.savez_compressed(os.path.join(OUT, name), qf=qf, gf=gf, final=out, versions=versions())
    print(name, out.shape, out.dtype)


def embed_case(name, n_img, seed_img):
    import torch
    ref = refshim.load_reference()
    imgs = R.synth_images(n_img, seed_img)
    names = ["im%03d" % i for i in range(n_img)]
    batches = [(imgs, names, list(range(n_img)), [0] * n_img)]
    out = dict(n_img=n_img, seed_img=seed_img, weight_seed=0, versions=versions())
    for S in (1, 2, 3):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = ref.models.create("resnet50", num_classes=0, num_split=S, pretrained=False)
        m.base.load_state_dict(R.make_state_dict(0), strict=False)
        m.eval()
        fl, _ = ref.evaluators.extract_features(m, batches, for_eval=False)
        fe, _ = ref.evaluators.extract_features(m, batches, for_eval=True)
        if S == 1:
            out["list_S1"] = torch.stack([fl[k] for k in names]).numpy()[None]
        else:
            out["list_S%d" % S] = torch.stack(
                [torch.stack([fl[k][b] for k in names]) for b in range(S + 1)]).numpy()
        out["eval_S%d" % S] = torch.stack([fe[k] for k in names]).numpy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def cycle_case(name, n, ns, num_split=2, lam=0.1, rhos=(1.6e-3, 1.6e-2, 5e-2), per_identity=8, noise=0.5, keep_rows=16):
    """The WHOLE path of selftraining.py:196-218, 255-313 through the unmodified reference, from IMAGES to labels:
    extract_features (fp32 torch CPU, list mode) on seeded identity images -> bank re-stacking -> compute_dist
    (re_ranking per bank: O-f32 and the as-is fp16 arithmetic) -> generate_selflabel (eps at rho, sklearn DBSCAN).
    compute_dist / generate_selflabel are the driver's OWN functions, imported from the reference's selftraining.py.
    Stored: labels, eps, rank tables, final_dist (float32-rounded upper triangle incl. diagonal; the rounding error
    6e-8 is far below the 1e-4 tolerance) for both arithmetic variants, and the features of the first rows."""
    import contextlib
    import io
    import types
    import torch
    from sklearn.metrics import adjusted_rand_score
    ref = refshim.load_reference()
    with refshim._reference_on_path(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib
        drv = importlib.import_module("selftraining")          # /root/reference/selftraining.py, unmodified
    torch.set_num_threads(os.cpu_count() or 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ref.models.create("resnet50", num_classes=0, num_split=num_split, pretrained=False)
    m.base.load_state_dict(R.make_state_dict(0), strict=False)
    m.eval()
    banks = num_split + 1 if num_split > 1 else 1
    feats = {}
    ident = {}
    for tag, cnt, seed in (("tgt", n, 1234), ("src", ns, 4321)):
        imgs, lab = R.synth_identity_images(cnt, seed, per_identity, noise)
        names = ["%s%05d" % (tag, i) for i in range(cnt)]
        batches = [(imgs[i:i + 64], names[i:i + 64], [0] * len(names[i:i + 64]), [0] * len(names[i:i + 64]))
                   for i in range(0, cnt, 64)]
        f, _ = ref.evaluators.extract_features(m, batches, print_freq=10 ** 9, for_eval=False)
        # selftraining.py:197-209 bank re-stacking, in data-set order
        feats[tag] = [torch.cat([f[k][i].unsqueeze(0) for k in names], 0) for i in range(banks)]
        ident[tag] = lab.numpy()
    out = dict(n=n, ns=ns, num_split=num_split, lam=lam, rhos=np.array(rhos), per_identity=per_identity, noise=noise,
               seed_tgt=1234, seed_src=4321, weight_seed=0, identity=ident["tgt"], versions=versions(),
               feat_tgt_head=np.stack([b[:keep_rows].numpy() for b in feats["tgt"]]),
               feat_src_head=np.stack([b[:keep_rows].numpy() for b in feats["src"]]))
    # how well-conditioned is the problem?  (1) spread of the features: nearest / 21st-nearest / median squared distance;
    # (2) the REFERENCE arithmetic (oracle O-f32) fed its own features perturbed by Gaussian noise of relative L2 size
    # 4e-3 per row -- the size of the bf16 trunk's feature error -- : rank-table and label changes caused by the
    # perturbation alone.  The CUDA path cannot be closer to the reference than the reference is to itself under that
    # perturbation; tests/test_gpu_whole_path.py compares against these numbers.
    from scipy.spatial.distance import cdist
    rng = np.random.RandomState(2024)
    cond = {k: [] for k in ("d2_nn", "d2_k21", "d2_median", "noise_rank_entry_mismatch", "noise_rank_set_mismatch_rows")}
    noisy_final = []
    for b in range(banks):
        t, s_ = feats["tgt"][b].numpy(), feats["src"][b].numpy()
        d2 = np.sort(cdist(t, t) ** 2, axis=1)
        cond["d2_nn"].append(float(np.median(d2[:, 1])))
        cond["d2_k21"].append(float(np.median(d2[:, 20])))
        cond["d2_median"].append(float(np.median(d2[:, n // 2])))

        def perturb(x):
            e = rng.randn(*x.shape).astype(np.float32)
            e *= (4e-3 * np.linalg.norm(x, axis=1, keepdims=True) / np.linalg.norm(e, axis=1, keepdims=True))
            return (x + e).astype(np.float32)
        st0, st1 = {}, {}
        O.re_ranking(s_, t, lambda_value=lam, mode="f32", stages=st0)
        _, f1 = O.re_ranking(perturb(s_), perturb(t), lambda_value=lam, mode="f32", stages=st1)
        r0, r1 = st0["rank"][:, :21], st1["rank"][:, :21]
        cond["noise_rank_entry_mismatch"].append(float((r0 != r1).mean()))
        cond["noise_rank_set_mismatch_rows"].append(float(np.mean([set(a) != set(c) for a, c in zip(r0, r1)])))
        noisy_final.append(f1)
    for k, v in cond.items():
        out[k] = np.array(v)
    print(name, "conditioning", {k: [round(x, 5) for x in v] for k, v in cond.items()})
    iu = np.triu_indices(n)
    for mode in ("f32", "ref"):
        # the driver's `from reid.rerank import *` bound re_ranking when selftraining.py was imported: patch the `np` of
        # THAT function's module (its globals), which is a fresh import of the same unmodified file
        ctx = refshim.f32_stable_globals(drv.re_ranking) if mode == "f32" else contextlib.nullcontext()
        with ctx, contextlib.redirect_stdout(io.StringIO()):
            _, r_dist = drv.compute_dist(feats["src"], feats["tgt"], lambda_value=lam, no_rerank=False,
                                         num_split=num_split)
        for b in range(banks):
            out["final_%s_b%d" % (mode, b)] = r_dist[b][iu].astype(np.float32)
            if mode == "f32":
                st = {}
                O.re_ranking(feats["src"][b].numpy(), feats["tgt"][b].numpy(), lambda_value=lam, mode="f32", stages=st)
                _, f_o = O.re_ranking(feats["src"][b].numpy(), feats["tgt"][b].numpy(), lambda_value=lam, mode="f32")
                assert np.array_equal(f_o, r_dist[b]), "oracle restatement != reference (O-f32) on bank %d" % b
                out["rank21_b%d" % b] = st["rank"][:, :21].astype(np.int16)
        for ri, rho in enumerate(rhos):
            args = types.SimpleNamespace(no_rerank=False, rho=rho)
            with contextlib.redirect_stdout(io.StringIO()):
                labels, clusters = drv.generate_selflabel([[]] * banks, r_dist, 0, args, [])
            for b in range(banks):
                out["labels_%s_r%d_b%d" % (mode, ri, b)] = labels[b].astype(np.int32)
                out["eps_%s_r%d_b%d" % (mode, ri, b)] = np.float64(clusters[b].eps)
            if mode == "f32":
                with contextlib.redirect_stdout(io.StringIO()):
                    labels_n, _ = drv.generate_selflabel([[]] * banks, noisy_final, 0, args, [])
                out["noise_ari_r%d" % ri] = np.array([adjusted_rand_score(labels[b], labels_n[b]) for b in range(banks)])
                print(name, "rho", rho, "ARI(reference, reference on 4e-3-perturbed features)",
                      [round(float(x), 3) for x in out["noise_ari_r%d" % ri]])
            print(name, mode, "rho", rho, "clusters", [int(l.max()) + 1 for l in labels],
                  "noise", [int((l < 0).sum()) for l in labels],
                  "ARI vs identity", [round(adjusted_rand_score(ident["tgt"], l), 3) for l in labels])
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "bytes", os.path.getsize(os.path.join(OUT, name)))


def triplet_cases(name):
    """reid/loss/triplet.py TripletLoss of the unmodified reference on CPU (forward + autograd backward)."""
    import torch
    refshim.load_reference()
    with refshim._reference_on_path():
        import reid.loss.triplet as T
    out = dict(versions=versions())
    cases = [  # P, K, d, seed, margin, use_semi, extra rows, sep
        (4, 4, 32, 0, 0.5, 1, 0, 1.0), (16, 4, 2048, 1, 0.5, 1, 0, 0.2), (16, 4, 512, 2, 0.0, 1, 3, 0.3),
        (8, 8, 256, 3, 0.3, 0, 0, 0.5), (32, 4, 512, 5, 0.5, 1, 0, 0.15), (5, 3, 77, 6, 0.3, 1, 2, 0.4)]
    out["cases"] = np.array(cases, dtype=np.float64)
    for ci, (P, K, d, seed, margin, semi, extra, sep) in enumerate(cases):
        x, t = TO.synth_batch(P, K, d, seed, sep, extra)
        xt = torch.from_numpy(x).requires_grad_(True)
        crit = T.TripletLoss(margin=margin, num_instances=K, use_semi=bool(semi))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loss, prec = crit(xt, torch.from_numpy(t), 0)
        loss.backward()
        out["x_%d" % ci], out["t_%d" % ci] = x, t
        out["loss_%d" % ci], out["prec_%d" % ci] = np.float32(loss.item()), np.float32(float(prec))
        out["grad_%d" % ci] = xt.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, [float(out["loss_%d" % c]) for c in range(len(cases))])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "cycle":        # only the whole-path golden (minutes of CPU time)
        cycle_case("cycle_n512_S2.npz", 512, 384)
        sys.exit(0)
    rerank_case("rerank_n160_d256.npz", 160, 150, 256, seed=3, lam=0.1, rhos=[1.6e-3, 1.6e-2, 5e-2])
    rerank_case("rerank_n257_d2048.npz", 257, 200, 2048, seed=7, lam=0.1, rhos=[1.6e-2, 4e-2])
    rerank_case("rerank_n96_d64_ties.npz", 96, 64, 64, seed=11, lam=0.3, rhos=[2e-2], per_cluster=8,
                noise=0.0)   # noise 0 => duplicate features: exact ties everywhere
    rerank_init_case("rerank_init_q40_g90.npz", 40, 90, 512, seed=5)
    embed_case("embed_4img.npz", 4, 1234)
    triplet_cases("triplet_cases.npz")
    cycle_case("cycle_n512_S2.npz", 512, 384)
