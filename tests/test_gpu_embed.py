"""GPU parity of the embedding path (ResNet-50 trunk on tcgen05 + flip TTA + pooled, normalised banks).

Floating-point kernel, so the oracle here is a plain PyTorch fp32 reference of the same op (and, end to
end, the reference model restated in oracle/resnet_oracle.py + the reference's own golden features).
The kernels compute in bf16 with fp32 accumulation; tolerances, written per test:
  * single convolution, inputs/weights already rounded to bf16: |err| <= 1e-2 * max|out| (one bf16 output
    rounding, 2^-9 relative, + accumulation order);
  * whole trunk (53 convolutions, activations re-rounded to bf16 after each): relative L2 error of a
    2048-d bank <= 8e-3 (measured 4.2e-3 on the B200; the round-1 bound was 3e-2) and cosine similarity >= 0.9995 against the fp32 reference.
"""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from ssg_b200 import _lib
    return _lib


def _conv(lib, x_nchw, w, bias, stride, res=None, relu=True):
    """Run ssg_op_conv on NCHW fp32 torch tensors (rounded to bf16), return NCHW fp32."""
    import torch
    B, C, H, W = x_nchw.shape
    cout, cin, k, _ = w.shape
    x = x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    wk = w.permute(0, 2, 3, 1).contiguous().view(cout, -1).to(torch.bfloat16)
    OH, OW = H // stride, W // stride
    y = torch.empty((B, OH, OW, cout), dtype=torch.bfloat16, device="cuda")
    scratch = torch.empty_like(x)
    r = res.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16) if res is not None else None
    lib.check(lib.load().ssg_op_conv(x.data_ptr(), B, H, W, cin, k, stride, wk.data_ptr(), bias.data_ptr(), cout,
                                     r.data_ptr() if r is not None else None, int(relu), y.data_ptr(),
                                     scratch.data_ptr(), lib.stream_ptr()))
    torch.cuda.synchronize()
    return y.float().permute(0, 3, 1, 2)


def _ref_conv(x, w, bias, stride, res=None, relu=True):
    import torch
    import torch.nn.functional as F
    xb, wb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        y = F.conv2d(xb, wb, bias, stride=stride, padding=w.shape[-1] // 2)
    if res is not None:
        y = y + res.to(torch.bfloat16).float()
    return F.relu(y) if relu else y


CONV_CASES = [
    # B, H, W, cin, cout, k, stride, residual
    (2, 64, 32, 64, 64, 1, 1, False), (2, 64, 32, 64, 256, 1, 1, True), (3, 64, 32, 256, 128, 1, 1, False),
    (2, 64, 32, 64, 64, 3, 1, False), (2, 32, 16, 128, 128, 3, 1, False), (3, 16, 8, 256, 256, 3, 1, False),
    (8, 8, 4, 512, 512, 3, 1, False), (5, 8, 4, 512, 512, 3, 1, False),
    (2, 64, 32, 128, 128, 3, 2, False), (2, 32, 16, 256, 256, 3, 2, False), (4, 16, 8, 512, 512, 3, 2, False),
    (2, 64, 32, 256, 512, 1, 2, False), (4, 16, 8, 1024, 2048, 1, 2, False), (4, 8, 4, 2048, 512, 1, 1, False),
    # conv3 of layers 3 / 4: 128x256 tiles with the residual sub-tile ring (one tile, ragged M, > 148 tiles)
    (3, 16, 8, 256, 1024, 1, 1, True), (5, 8, 4, 512, 2048, 1, 1, True), (64, 16, 8, 256, 1024, 1, 1, True),
]


@pytest.mark.parametrize("B,H,W,cin,cout,k,stride,use_res", CONV_CASES)
def test_conv_blocks_match_torch(lib, B, H, W, cin, cout, k, stride, use_res):
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + cin + cout + k + stride)
    x = torch.randn(B, cin, H, W, generator=g, device="cuda")
    w = torch.randn(cout, cin, k, k, generator=g, device="cuda") / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g, device="cuda") * 0.1
    res = torch.randn(B, cout, H // stride, W // stride, generator=g, device="cuda") if use_res else None
    got = _conv(lib, x, w, bias, stride, res)
    want = _ref_conv(x, w, bias, stride, res)
    tol = 1e-2 * float(want.abs().max())
    assert float((got - want).abs().max()) <= tol


def test_fold_bn_and_stem_match_torch(lib):
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(7)
    n = 3
    img = torch.randn(n, 3, 256, 128, generator=g, device="cuda")
    w = torch.randn(64, 3, 7, 7, generator=g, device="cuda") * 0.1
    gamma = torch.rand(64, generator=g, device="cuda") + 0.5
    beta = torch.randn(64, generator=g, device="cuda") * 0.1
    mean = torch.randn(64, generator=g, device="cuda") * 0.1
    var = torch.rand(64, generator=g, device="cuda") + 0.5
    wf = torch.empty((64, 192), dtype=torch.bfloat16, device="cuda")
    bf = torch.empty(64, dtype=torch.float32, device="cuda")
    L = lib.load()
    lib.check(L.ssg_op_fold_bn(w.data_ptr(), 64, 3, 7, gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                               var.data_ptr(), 1e-5, 192, wf.data_ptr(), bf.data_ptr(), lib.stream_ptr()))
    scale = gamma / torch.sqrt(var + 1e-5)
    w_ref = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).reshape(64, 147)
    assert torch.equal(wf[:, :147], w_ref.to(torch.bfloat16)) and float(wf[:, 147:].float().abs().max()) == 0.0
    torch.testing.assert_close(bf, beta - mean * scale, rtol=1e-6, atol=1e-7)
    col = torch.empty((2 * n * 8192, 192), dtype=torch.bfloat16, device="cuda")
    conv = torch.empty((2 * n, 128, 64, 64), dtype=torch.bfloat16, device="cuda")
    pool = torch.empty((2 * n, 64, 32, 64), dtype=torch.bfloat16, device="cuda")
    lib.check(L.ssg_op_stem(img.data_ptr(), n, 1, wf.data_ptr(), bf.data_ptr(), col.data_ptr(), conv.data_ptr(),
                            pool.data_ptr(), lib.stream_ptr()))
    torch.cuda.synchronize()
    both = torch.cat([img, img.flip(3)], 0).to(torch.bfloat16).float()
    w_eff = wf[:, :147].float().view(64, 7, 7, 3).permute(0, 3, 1, 2)
    ref = F.relu(F.conv2d(both, w_eff, bf, stride=2, padding=3))
    got = conv.float().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    ref_pool = F.max_pool2d(got, 3, 2, 1)                       # pooling itself is exact on the bf16 values
    assert torch.equal(pool.float().permute(0, 3, 1, 2), ref_pool)


@pytest.mark.parametrize("S,eval_mode", [(1, 0), (2, 0), (3, 0), (3, 1), (2, 1)])
def test_pooled_tail_matches_torch(lib, S, eval_mode):
    import torch
    import torch.nn.functional as F
    n = 5
    g = torch.Generator(device="cuda").manual_seed(S)
    x = torch.rand(2 * n, 8, 4, 2048, generator=g, device="cuda").to(torch.bfloat16)
    banks = S + 1 if S > 1 else 1
    feat = torch.zeros((banks, n + 2, 2048) if not eval_mode else (n + 2, banks * 2048), device="cuda")
    lib.check(lib.load().ssg_op_pooled_tail(x.data_ptr(), n, S, eval_mode, 1, feat.data_ptr(),
                                            0 if eval_mode else feat.stride(0), 2, lib.stream_ptr()))
    torch.cuda.synchronize()
    xn = x.float().permute(0, 3, 1, 2)

    def pools(t):
        out = [F.avg_pool2d(t, t.shape[2:]).flatten(1)]
        if S > 1:
            step = 8 // S
            out += [F.avg_pool2d(t[:, :, step * s: step * (s + 1)], (step, 4)).flatten(1) for s in range(S)]
        return out
    a, b = pools(xn[:n]), pools(xn[n:])
    if eval_mode:
        o = torch.cat(a, 1) + torch.cat(b, 1)
        want = o / o.norm(dim=1, keepdim=True)
        torch.testing.assert_close(feat[2:], want, rtol=1e-5, atol=1e-6)
    else:
        for k in range(banks):
            o = a[k] + b[k]
            torch.testing.assert_close(feat[k, 2:], o / o.norm(dim=1, keepdim=True), rtol=1e-5, atol=1e-6)
    assert float(feat[..., :2, :].abs().max() if not eval_mode else feat[:2].abs().max()) == 0.0


def _rel_err(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("S", [1, 2, 3])
def test_trunk_end_to_end_against_fp32_oracle_and_reference_golden(S, golden_dir):
    import torch
    import ssg_b200
    from oracle import resnet_oracle as R
    g = np.load(os.path.join(golden_dir, "embed_4img.npz"))
    n = int(g["n_img"])
    imgs = R.synth_images(n, int(g["seed_img"]))
    model = R.build_model(S, int(g["weight_seed"]))
    names = ["im%03d" % i for i in range(n)]
    batches = [(imgs, names, list(range(n)), [0] * n)]
    # list mode (for_eval=False) and eval mode through the drop-in extract_features
    fl, lab = ssg_b200.extract_features(model, batches, for_eval=False)
    fe, _ = ssg_b200.extract_features(model, batches, for_eval=True)
    assert list(fl.keys()) == names and lab["im002"] == 2
    gl, ge = g["list_S%d" % S], g["eval_S%d" % S]
    for i, k in enumerate(names):
        if S > 1:
            assert isinstance(fl[k], list) and len(fl[k]) == S + 1 and not fl[k][0].is_cuda
            for b in range(S + 1):
                want = torch.from_numpy(gl[b, i])
                assert _rel_err(fl[k][b], want) <= 8e-3
                assert float(torch.dot(fl[k][b], want)) >= 0.9995
                assert abs(float(fl[k][b].norm()) - 1.0) < 1e-5
        else:
            assert _rel_err(fl[k], torch.from_numpy(gl[0, i])) <= 8e-3
        want = torch.from_numpy(ge[i])
        assert fe[k].shape == want.shape
        assert _rel_err(fe[k], want) <= 8e-3 and float(torch.dot(fe[k], want)) >= 0.9995


def test_trunk_batch_invariance_and_reset_params_regime():
    """(1) the same image gives the same features whatever batch it rides in (tiles span 4 images in layer4);
    (2) the reference's own random init (std 1e-3, resnet.py:136-148) drives activations to ~1e-10: bf16 keeps
    the exponent range, features stay finite and close to the fp32 oracle (SURVEY.md hard part 4)."""
    import torch
    import ssg_b200
    from oracle import resnet_oracle as R
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = R.build_model(2, 0)
    imgs = R.synth_images(7, 5).cuda()
    plan = ssg_b200.EmbedPlan(16)
    plan.load_model(model)
    a = plan.forward(imgs, 2).clone()
    b = plan.forward(imgs[2:5].contiguous(), 2).clone()
    torch.cuda.synchronize()
    assert torch.equal(a[:, 2:5], b)
    for m in model.base.modules():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.normal_(m.weight, std=0.001)
        elif isinstance(m, torch.nn.BatchNorm2d):
            torch.nn.init.constant_(m.weight, 1); torch.nn.init.constant_(m.bias, 0)
            m.running_mean.zero_(); m.running_var.fill_(1)
    plan.load_model(model)
    got = plan.forward(imgs, 2)
    with torch.no_grad():
        mo = model.cuda()
        x1 = mo(imgs, False)[0]
        x1f = mo(imgs.flip(3), False)[0]
    assert bool(torch.isfinite(got).all())
    for k in range(3):
        o = x1[k] + x1f[k]
        want = o / o.norm(dim=1, keepdim=True)
        assert _rel_err(got[k], want) <= 5e-2


def test_eug_label_estimation_matches_numpy_restatement():
    """reid/eug.py:193-253 (both branches) on features extracted by the CUDA trunk, against the reference's numpy
    arithmetic applied to the same features."""
    import torch
    import reid.eug
    from oracle import resnet_oracle as R, ssg_oracle as O
    model = R.build_model(1, 0)
    rng = np.random.RandomState(0)
    pats = R.synth_images(6, 11)
    def make(n, seed):
        g = torch.Generator().manual_seed(seed)
        lab = torch.randint(0, 6, (n,), generator=g)
        return pats[lab] + 0.4 * torch.randn(n, 3, 256, 128, generator=g), lab.numpy()
    l_img, l_lab = make(12, 1)
    u_img, u_lab = make(20, 2)
    l_data = [("l%d" % i, int(l_lab[i]), 0) for i in range(12)]
    u_data = [("u%d" % i, int(u_lab[i]), 0) for i in range(20)]
    store = {"l": l_img, "u": u_img}

    def loader_factory(dataset, training):
        imgs = store[dataset[0][0][0]]
        names = [d[0] for d in dataset]
        return [(imgs, names, [d[1] for d in dataset], [0] * len(names))]

    for rerank in (False, True):
        eug = reid.eug.EUG("resnet50", 32, "Weight" if rerank else "Dissimilarity", 0, None, l_data, u_data, None, 20,
                           pretrained_model=model, rerank=rerank, loader_factory=loader_factory)
        out = eug.estimate_label()
        u_f, l_f = eug.get_feature(u_data), eug.get_feature(l_data)
        if not rerank:
            labels, scores = out
            d = np.stack([np.linalg.norm(l_f - uf, axis=1) for uf in u_f])
            assert np.array_equal(labels, l_lab[d.argmin(1)].astype(np.float64))
            np.testing.assert_allclose(scores, -d.min(1), atol=1e-5)
        else:
            labels, scores, conf = out
            rr = O.re_ranking_init(u_f @ l_f.T, u_f @ u_f.T, l_f @ l_f.T)
            idx = rr.argmin(1)
            assert np.mean(labels == l_lab[idx]) >= 0.9           # near-ties may resolve differently (GPU GEMM vs np.dot)
            # the similarity blocks come from a GPU GEMM here and from np.dot in the restatement: near-tied ranks may
            # resolve differently, which moves individual re-ranked distances; agreement is statistical
            assert np.median(np.abs(scores + rr.min(1))) < 1e-3
            assert conf.shape == (20,) and np.all(conf <= 1.0 + 1e-5) and np.all(np.isfinite(conf))
        sel = eug.select_top_data(scores, 5)
        assert sel.sum() == 5 and len(eug.generate_new_train_data(sel, labels)) == 12 + 5


def test_uint8_pixels_give_bit_identical_features():
    """Row f5: raw uint8 HWC pixels normalised on the device (ssg_embed_forward_u8) == the loader's ToTensor +
    Normalize on the CPU (selftraining.py:36-45) followed by the float32 path, bit for bit; also through
    embed_images (host uint8 tensor, a quarter of the H2D bytes) and the drop-in extract_features."""
    import torch
    import ssg_b200
    from ssg_b200.embed import IMAGENET_MEAN, IMAGENET_STD
    from oracle import resnet_oracle as R
    model = R.build_model(2, 0)
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (5, 256, 128, 3), dtype=torch.uint8, generator=g)
    u8[0, :, 0, :] = 0
    u8[0, :, 127, :] = 255                              # image borders + extreme values
    mean, std = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1), torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    f32 = u8.permute(0, 3, 1, 2).to(torch.float32).div(255).sub_(mean).div_(std).contiguous()   # T.ToTensor, T.Normalize
    plan = ssg_b200.EmbedPlan(16)
    plan.load_model(model)
    a = plan.forward(f32.cuda(), 2).clone()
    b = plan.forward(u8.cuda(), 2).clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    c = ssg_b200.embed_images(model, u8.pin_memory(), batch=4)       # 2 batches through the copy/compute overlap
    assert torch.equal(c, a)
    names = ["u%d" % i for i in range(5)]
    fa, _ = ssg_b200.extract_features(model, [(f32, names, [0] * 5, [0] * 5)], for_eval=True)
    fb, _ = ssg_b200.extract_features(model, [(u8, names, [0] * 5, [0] * 5)], for_eval=True)
    assert all(torch.equal(fa[k], fb[k]) for k in names)
    with pytest.raises(ValueError):
        plan.forward(u8.permute(0, 3, 1, 2).contiguous().cuda(), 2)  # CHW uint8 is not a supported layout


def test_conv_variants_are_bit_identical(tmp_path):
    """The default trunk (stem with resident weights, parity-plane operand staging and fused max-pool; 128x256
    residual-ring tiles) against the plain variants (SSG_STEM_BRES=0 SSG_STEM_POOL=0 SSG_CONV_BN256_RES=0, read once
    per process -> a subprocess): the same products accumulate in the same order, so the features must agree bit for
    bit."""
    import subprocess
    import sys
    import torch
    import ssg_b200
    from oracle import resnet_oracle as R
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, numpy as np, torch\n"
        "sys.path[:0] = %r\n"
        "import ssg_b200\n"
        "from oracle import resnet_oracle as R\n"
        "plan = ssg_b200.EmbedPlan(16); plan.load_model(R.build_model(2, 0))\n"
        "out = plan.forward(R.synth_images(9, 11).cuda(), 2); torch.cuda.synchronize()\n"
        "np.save(sys.argv[1], out.cpu().numpy())\n" % ([os.path.join(root, "self-similarity-grouping_b200"), root],))
    plan = ssg_b200.EmbedPlan(16)
    plan.load_model(R.build_model(2, 0))
    got = plan.forward(R.synth_images(9, 11).cuda(), 2)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    # round 2: the layer-1 chained kernel (conv3 + next conv1 in one launch, SSG_CONV_CHAIN), the kernel-row-sharing
    # 3x3 with resident weights (SSG_KHS_BRES), the deeper pipelines (SSG_CONV_NORES) and the two-CTA cta_group::2
    # tiles (SSG_CONV_PAIR) are defaults; their plain predecessors must give the same bits
    for name, switches in (("plain", dict(SSG_STEM_BRES="0", SSG_STEM_POOL="0", SSG_CONV_BN256_RES="0")),
                           ("no_chain", dict(SSG_CONV_CHAIN="0")),
                           ("khs_streamed", dict(SSG_KHS_BRES="0")),
                           ("no_pairs", dict(SSG_CONV_PAIR="0")),            # default: two-CTA tiles where they win
                           ("all_pairs", dict(SSG_CONV_PAIR="7")),
                           ("pdl", dict(SSG_PDL="1")),                       # programmatic dependent launch (opt-in)
                           ("khs_pair", dict(SSG_KHS_PAIR="1")),             # two-CTA form of the layer-1 3x3 kernel
                           ("wide_k128", dict(SSG_WIDE_K="128")),            # 256-wide tiles also for K = 128
                           ("round1_default", dict(SSG_CONV_CHAIN="0", SSG_KHS_BRES="0", SSG_CONV_PAIR="0",
                                                   SSG_CONV_NORES="0"))):
        out_file = str(tmp_path / (name + ".npy"))
        subprocess.run([sys.executable, "-c", script, out_file], check=True, env=dict(os.environ, **switches), timeout=300)
        assert np.array_equal(got, np.load(out_file)), name


def test_host_image_path_equals_device_image_path_across_consecutive_calls():
    """embed_images on pinned HOST images (double-buffered staging on a copy stream) must give the bits of the
    device-image path, also when two sets are embedded back to back: round 1 re-allocated the staging buffers per call
    and the second call's copies overwrote the buffers the first call's last forwards were still reading (found in round
    2: the last batches of the first set came out wrong by up to 4e-3 absolute)."""
    import torch
    import ssg_b200
    from ssg_b200 import synth
    model = synth.build_model(2, 0)
    dev = torch.device("cuda", 0)
    a, _ = synth.synth_images(700, 11, dev)
    b, _ = synth.synth_images(700, 12, dev)
    fa_dev = ssg_b200.embed_images(model, a, 2, False, 256, 0).clone()
    fb_dev = ssg_b200.embed_images(model, b, 2, False, 256, 0).clone()
    ha, hb = a.cpu().pin_memory(), b.cpu().pin_memory()
    for _ in range(3):                                # back to back, no synchronisation in between
        fa = ssg_b200.embed_images(model, ha, 2, False, 256, 0)
        fb = ssg_b200.embed_images(model, hb, 2, False, 256, 0)
        torch.cuda.synchronize()
        assert torch.equal(fa, fa_dev) and torch.equal(fb, fb_dev)
    # raw uint8 pixels through the same staging path
    u8 = (a.permute(0, 2, 3, 1) * 50 + 128).clamp(0, 255).to(torch.uint8).contiguous()
    f_u8_dev = ssg_b200.embed_images(model, u8, 2, False, 256, 0).clone()
    h_u8 = u8.cpu().pin_memory()
    f1 = ssg_b200.embed_images(model, h_u8, 2, False, 256, 0)
    f2 = ssg_b200.embed_images(model, ha, 2, False, 256, 0)
    torch.cuda.synchronize()
    assert torch.equal(f1, f_u8_dev) and torch.equal(f2, fa_dev)
