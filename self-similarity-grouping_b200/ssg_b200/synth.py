"""Synthetic inputs of the benchmark: seeded images with identity structure and a random-init
ResNet-50 state_dict with randomised BatchNorm statistics (so that BN folding is exercised)."""
from collections import OrderedDict


def make_state_dict(seed=0, randomise_bn=True):
    """torchvision ResNet-50 default init (seeded) + randomised BN affine/statistics; keys as in torchvision."""
    import torch
    import torchvision
    torch.manual_seed(seed)
    net = torchvision.models.resnet50(weights=None)
    sd = net.state_dict()
    if randomise_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for k in list(sd.keys()):
            if k.endswith("running_mean"):
                sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
            elif k.endswith("running_var"):
                sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
            elif ("bn" in k or "downsample.1" in k) and k.endswith("weight"):
                sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
            elif ("bn" in k or "downsample.1" in k) and k.endswith("bias"):
                sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    return OrderedDict((k, v) for k, v in sd.items() if not k.startswith("fc."))


def build_model(num_split=2, seed=0):
    """A reference-style model (reid.models.create) carrying the synthetic weights, in eval mode."""
    import warnings
    import reid.models
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = reid.models.create("resnet50", num_classes=0, num_split=num_split, pretrained=False)
    m.base.load_state_dict(make_state_dict(seed), strict=False)
    return m.eval()


def synth_images(n, seed, device, per_identity=20, noise=0.5, h=256, w=128, chunk=1024):
    """[n,3,h,w] float32 on `device`: one random pattern per identity + per-image Gaussian noise, so that the
    embedded features carry cluster structure for re-ranking/DBSCAN (random images alone embed to nearly
    identical features)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    ids = max(n // per_identity, 1)
    pat = torch.randn(ids, 3, h, w, generator=g, device=device)
    lab = torch.randint(0, ids, (n,), generator=g, device=device)
    out = torch.empty(n, 3, h, w, device=device)
    # torch's vectorised gather kernel asserts on sources beyond 2^31 bytes (seen at n = 126 441: 6 322 patterns = 2.5 GB):
    # gather from slices of at most `part` identities then (same values, same random stream)
    part = max(1, (1 << 30) // (3 * h * w * 4))
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        idx = lab[r0:r1]
        if ids <= part:
            base = pat[idx]
        else:
            base = torch.empty(r1 - r0, 3, h, w, device=device)
            for p0 in range(0, ids, part):
                sel = (idx >= p0) & (idx < p0 + part)
                if bool(sel.any()):
                    base[sel] = pat[p0:p0 + part][idx[sel] - p0]
        out[r0:r1] = base + noise * torch.randn(r1 - r0, 3, h, w, generator=g, device=device)
    return out, lab
