"""The torch module the drivers build, train and hand to extract_features (reid/models/resnet.py:31-134):
torchvision ResNet trunk up to layer4 in ``self.base``, global + ``num_split`` stripe average pools, the
2048->num_features ``feat``/``feat_bn`` head.  Same attribute names and state_dict keys as the reference, so its
checkpoints load unchanged.  The CUDA embedding path ingests ``self.base`` (depth 50); this forward (plain
PyTorch) is what the fine-tuning step of the driver differentiates through."""
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init
import torchvision

__all__ = ['ResNet', 'resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152']

_DEPTHS = {18: 'resnet18', 34: 'resnet34', 50: 'resnet50', 101: 'resnet101', 152: 'resnet152'}


class ResNet(nn.Module):
    def __init__(self, depth, checkpoint=None, pretrained=True, num_features=2048, dropout=0.1, num_classes=0,
                 num_split=1, mode='Dissimilarity', cluster=False):
        super(ResNet, self).__init__()
        if depth not in _DEPTHS:
            raise KeyError("Unsupported depth:", depth)
        if cluster:
            raise NotImplementedError("the DEC head (--dce-loss) is outside the pseudo-label hot path")
        self.depth, self.checkpoint, self.pretrained = depth, checkpoint, pretrained
        self.num_features, self.dropout, self.num_classes = num_features, dropout, num_classes
        self.num_split, self.cluster = num_split, cluster
        if self.dropout > 0:
            self.drop = nn.Dropout(self.dropout)
        ctor = getattr(torchvision.models, _DEPTHS[depth])
        self.base = ctor(weights="IMAGENET1K_V1" if pretrained else None)
        out_planes = self.base.fc.in_features
        if self.checkpoint:
            state = torch.load(checkpoint, map_location='cpu')
            self.load_state_dict(state['state_dict'], strict=False)
        if self.num_features > 0:
            self.feat = nn.Linear(out_planes, self.num_features, bias=False)
            self.feat_bn = nn.BatchNorm1d(self.num_features)
            self.relu = nn.ReLU(inplace=True)
            init.normal_(self.feat.weight, std=0.001)
            init.constant_(self.feat_bn.weight, 1)
            init.constant_(self.feat_bn.bias, 0)
        if self.num_classes > 0:
            self.classifier_x2 = nn.Linear(self.num_features, self.num_classes)
            init.normal_(self.classifier_x2.weight, std=0.001)
            init.constant_(self.classifier_x2.bias, 0)
        if not self.pretrained:
            self.reset_params()

    def trunk(self, x):
        for name, module in self.base._modules.items():
            if name == 'avgpool':
                break
            x = module(x)
        return x

    def forward(self, x, for_eval=False):
        x = self.trunk(x)
        pooled = F.avg_pool2d(x, x.size()[2:]).view(x.size(0), -1)
        if self.num_split > 1:
            rows = x.size(2) // self.num_split
            x1 = [pooled]
            for s in range(self.num_split):
                stripe = x[:, :, rows * s: rows * (s + 1), :]
                x1.append(F.avg_pool2d(stripe, stripe.size()[2:]).view(stripe.size(0), -1))
        else:
            x1 = pooled
        x2 = None
        if self.num_features > 0:
            x2 = self.relu(self.feat_bn(self.feat(pooled)))
        if self.num_classes > 0:
            x2 = self.classifier_x2(self.drop(x2))
        if for_eval and isinstance(x1, list):
            x1 = torch.cat(x1, dim=1)
        return x1, x2

    def reset_params(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    init.constant_(m.bias, 0)


def resnet18(**kwargs):
    return ResNet(18, **kwargs)


def resnet34(**kwargs):
    return ResNet(34, **kwargs)


def resnet50(**kwargs):
    return ResNet(50, **kwargs)


def resnet101(**kwargs):
    return ResNet(101, **kwargs)


def resnet152(**kwargs):
    return ResNet(152, **kwargs)
