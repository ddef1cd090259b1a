"""CPU restatement (torch fp32) of the embedding half of the hot path.  TEST INFRASTRUCTURE ONLY.

  * :class:`OracleResNet`  — reid/models/resnet.py:31-134 for depth 50, num_classes=0, cluster=False:
    torchvision ResNet-50 trunk up to layer4, global + `num_split` horizontal-stripe average pools.
    (torchvision is the reference's own third-party dependency, resnet.py:19-23; it is in the image.)
  * :func:`extract_features` — reid/evaluators.py:18-60 + reid/feature_extraction/cnn.py:10-23.
  * :func:`make_state_dict`  — the shared synthetic weights of SURVEY.md §8d (torchvision default init,
    seed 0, randomised BatchNorm statistics so that BN folding is exercised).

Pinned against the unmodified reference model in tests/test_oracle_vs_reference.py (build container)
and through tests/golden/embed_*.npz elsewhere.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F
import torchvision


class OracleResNet(torch.nn.Module):
    def __init__(self, num_split=1, num_features=2048):
        super().__init__()
        self.num_split = num_split
        self.base = torchvision.models.resnet50(weights=None)
        # resnet.py:64-70 — x2 head (computed, unused by feature extraction)
        self.feat = torch.nn.Linear(2048, num_features, bias=False)
        self.feat_bn = torch.nn.BatchNorm1d(num_features)

    def trunk(self, x):
        # resnet.py:87-92 — every child of torchvision's resnet up to (excluding) avgpool
        b = self.base
        x = b.maxpool(b.relu(b.bn1(b.conv1(x))))
        return b.layer4(b.layer3(b.layer2(b.layer1(x))))

    def forward(self, x, for_eval=False):
        x = self.trunk(x)
        if self.num_split > 1:
            # resnet.py:93-108 — note h // S rows per stripe: S=3 on an 8-row map drops rows 6-7
            h = x.size(2)
            x1 = [F.avg_pool2d(x, x.size()[2:]).view(x.size(0), -1)]
            step = h // self.num_split
            for s in range(self.num_split):
                xx = x[:, :, step * s: step * (s + 1), :]
                x1.append(F.avg_pool2d(xx, xx.size()[2:]).view(xx.size(0), -1))
        else:
            x1 = F.avg_pool2d(x, x.size()[2:]).view(x.size(0), -1)
        x2 = F.relu(self.feat_bn(self.feat(F.avg_pool2d(x, x.size()[2:]).view(x.size(0), -1))))
        if for_eval and isinstance(x1, list):
            x1 = torch.cat(x1, dim=1)      # resnet.py:122-124
        return x1, x2


def make_state_dict(seed=0, randomise_bn=True):
    """Synthetic trunk weights shared by oracle and kernels (keys = torchvision resnet50 names)."""
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


def build_model(num_split=1, seed=0):
    m = OracleResNet(num_split)
    m.base.load_state_dict(make_state_dict(seed), strict=False)
    return m.eval()


def fliplr(img):
    """reid/evaluators.py:12-16."""
    return img.flip(3)


def extract_features(model, batches, for_eval=True):
    """reid/evaluators.py:18-60: batches yield (imgs, fnames, pids, cams)."""
    model.eval()
    features, labels = OrderedDict(), OrderedDict()
    with torch.no_grad():
        for imgs, fnames, pids, cams in batches:
            out = model(imgs, for_eval)[0]
            out_f = model(fliplr(imgs), for_eval)[0]
            if (not for_eval) and isinstance(out, list):
                banks = []
                for a, b in zip(out, out_f):
                    o = a + b
                    banks.append(o / torch.norm(o, p=2, dim=1, keepdim=True))
                for idx, (fname, pid) in enumerate(zip(fnames, pids)):
                    features[fname] = [x[idx] for x in banks]
                    labels[fname] = pid
            else:
                o = out + out_f
                o = o / torch.norm(o, p=2, dim=1, keepdim=True)
                for fname, f, pid in zip(fnames, o, pids):
                    features[fname] = f
                    labels[fname] = pid
    return features, labels


def synth_images(n, seed=1234, h=256, w=128):
    """torch.randn(N,3,256,128) with a seeded CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, h, w, generator=g)


def synth_identity_images(n, seed=1234, per_identity=8, noise=0.5, h=256, w=128):
    """[n,3,h,w] float32 on the CPU with identity structure (one random pattern per identity + per-image Gaussian
    noise) so that the embedded features cluster: random images alone embed to nearly identical features
    (SURVEY.md 8d).  Seeded CPU generator: the same bits in the build container and on the GPU box.
    -> (images, identity of every image)."""
    g = torch.Generator().manual_seed(seed)
    ids = max(n // per_identity, 1)
    pat = torch.randn(ids, 3, h, w, generator=g)
    lab = torch.randint(0, ids, (n,), generator=g)
    out = torch.empty(n, 3, h, w)
    for r0 in range(0, n, 64):
        r1 = min(n, r0 + 64)
        out[r0:r1] = pat[lab[r0:r1]] + noise * torch.randn(r1 - r0, 3, h, w, generator=g)
    return out, lab
