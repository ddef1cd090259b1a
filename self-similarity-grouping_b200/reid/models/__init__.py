"""Model factory with the reference's surface (reid/models/__init__.py:6-53): names(), create(name, ...).
Only the ResNet family used by the drivers is provided (run.sh uses resnet50)."""
from .resnet import ResNet, resnet18, resnet34, resnet50, resnet101, resnet152  # noqa: F401

__factory = {
    'resnet18': resnet18,
    'resnet34': resnet34,
    'resnet50': resnet50,
    'resnet101': resnet101,
    'resnet152': resnet152,
}


def names():
    return sorted(__factory.keys())


def create(name, *args, **kwargs):
    if name not in __factory:
        raise KeyError("Unknown model:", name)
    return __factory[name](*args, **kwargs)
