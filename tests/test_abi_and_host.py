"""CPU: the C-ABI library loads without a GPU and exports exactly what include/ssg_b200.h declares;
host-side logic that needs no device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "ssg_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ssg_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    # the ctypes prototype table mirrors the header one to one
    assert sorted(_lib.PROTOTYPES.keys()) == declared
    assert lib.ssg_version() >= 100


def test_constants_match_header():
    from ssg_b200 import _lib
    with open(os.path.join(ROOT, "include", "ssg_b200.h")) as f:
        src = f.read()
    defs = dict(re.findall(r"#define\s+(SSG_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", src))
    assert int(defs["SSG_RANK_STRIDE"]) == _lib.RANK_STRIDE
    assert int(defs["SSG_V_STRIDE"]) == _lib.V_STRIDE and int(defs["SSG_VQ_STRIDE"]) == _lib.VQ_STRIDE
    assert int(defs["SSG_ERR_CAPACITY"]) == _lib.ERR_CAPACITY and int(defs["SSG_DIST_TENSOR"]) == _lib.DIST_TENSOR
    assert int(defs["SSG_STAGE_FLAGGED"]) == _lib.STAGE_FLAGGED and int(defs["SSG_F64"]) == _lib.F64


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ssg_b200
    from ssg_b200 import _lib
    x = np.zeros((8, 16), np.float32)
    with pytest.raises(_lib.SsgError):
        ssg_b200.re_ranking(x, x)
    with pytest.raises(_lib.SsgError):
        ssg_b200.DBSCAN(eps=0.5, min_samples=4, metric="precomputed").fit_predict(np.zeros((4, 4)))
    with pytest.raises(_lib.SsgError):
        ssg_b200.eps_estimate(np.zeros((4, 4)), 0.1)


def test_embed_layer_table_is_torchvision_resnet50():
    import torchvision
    from ssg_b200 import embed
    table = embed.layer_table()
    assert len(table) == 53
    sd = torchvision.models.resnet50(weights=None).state_dict()
    convs = [k[:-len(".weight")] for k, v in sd.items() if v.dim() == 4]
    assert sorted(t[5] for t in table) == sorted(convs)
    macs = 0
    for _, cin, cout, k, stride, ck, bk in table:
        assert tuple(sd[ck + ".weight"].shape) == (cout, cin, k, k)
        assert sd[bk + ".running_var"].shape[0] == cout
    # 2 669 150 208 conv MACs per 256x128 image
    H, W = 256, 128
    sizes = {}
    x = (H // 2, W // 2)
    for _, cin, cout, k, stride, ck, bk in table:
        if ck == "conv1":
            macs += x[0] * x[1] * cout * cin * k * k
            cur = (x[0] // 2, x[1] // 2)
            sizes["in"] = cur
            continue
        name = ck.split(".")
        if name[2] == "conv1":
            block_in = sizes["in"]
            o = block_in
        elif name[2] == "conv2":
            o = (sizes["in"][0] // stride, sizes["in"][1] // stride)
            sizes["mid"] = o
        elif name[2] == "conv3":
            o = sizes["mid"]
        else:
            o = sizes["mid"]
        macs += o[0] * o[1] * cout * cin * k * k
        if name[2] == "conv3" and not any(t[5] == ".".join(name[:2]) + ".downsample.0" for t in table):
            sizes["in"] = sizes["mid"]
        if name[2] == "downsample":
            sizes["in"] = sizes["mid"]
    assert macs == 2669150208


def test_keep_mask_and_dbscan_params():
    import ssg_b200
    m = ssg_b200.generate_keep_mask([np.array([0, -1, 2, 3]), np.array([1, 1, -1, 0])])
    assert m.tolist() == [True, False, False, True]
    est = ssg_b200.DBSCAN(eps=0.3, min_samples=4, metric="precomputed", n_jobs=8)
    assert est.get_params()["eps"] == 0.3 and est.set_params(eps=0.5).eps == 0.5
    with pytest.raises(ValueError):
        ssg_b200.DBSCAN(metric="euclidean").fit(np.zeros((3, 3)))


def test_reid_drop_in_surface():
    import inspect
    import reid
    import reid.evaluators, reid.rerank, reid.rerank_initial, reid.feature_extraction, reid.models  # noqa: E401
    sig = inspect.signature(reid.rerank.re_ranking)
    assert list(sig.parameters)[:8] == ["input_feature_source", "input_feature", "k1", "k2", "lambda_value",
                                        "MemorySave", "Minibatch", "no_rerank"]
    assert sig.parameters["k1"].default == 20 and sig.parameters["lambda_value"].default == 0.2
    assert list(inspect.signature(reid.evaluators.extract_features).parameters) == \
        ["model", "data_loader", "print_freq", "for_eval", "metric"]
    assert list(inspect.signature(reid.rerank_initial.re_ranking_init).parameters) == \
        ["q_g_dist", "q_q_dist", "g_g_dist", "k1", "k2", "lambda_value"]
    assert list(inspect.signature(reid.feature_extraction.extract_cnn_feature).parameters) == \
        ["model", "inputs", "for_eval", "modules"]
    # `from reid.rerank import *` must shadow sklearn's DBSCAN in the driver namespace (selftraining.py:27-28)
    ns = {}
    exec("from sklearn.cluster import DBSCAN\nfrom reid.rerank import *", ns)
    assert ns["DBSCAN"].__module__.startswith("ssg_b200")
    m = reid.models.create("resnet50", num_classes=0, num_split=2, pretrained=False)
    keys = set(m.state_dict().keys())
    assert "base.layer4.2.conv3.weight" in keys and "feat.weight" in keys and "feat_bn.running_mean" in keys


def test_fine_tune_surface_and_host_errors():
    """Row f1: TripletLoss / FinedTrainer2 keep the reference's constructor and call signatures
    (reid/loss/triplet.py:12,19; reid/trainers.py:205,211,257) and, like everything else, refuse to run without a GPU."""
    import inspect
    import torch
    from reid.loss import TripletLoss
    from reid.trainers import FinedTrainer2
    assert list(inspect.signature(TripletLoss.__init__).parameters) == ["self", "margin", "num_instances", "use_semi"]
    assert list(inspect.signature(TripletLoss.forward).parameters) == ["self", "inputs", "targets", "epoch", "w"]
    assert list(inspect.signature(FinedTrainer2.__init__).parameters) == ["self", "model", "criterions", "beta"]
    assert list(inspect.signature(FinedTrainer2.train).parameters) == \
        ["self", "epoch", "train_loader", "optimizer", "print_freq"]
    crit = TripletLoss(margin=0.5, num_instances=4)
    assert crit.margin == 0.5 and crit.K == 4 and crit.use_semi is True
    if not torch.cuda.is_available():
        from ssg_b200 import _lib
        with pytest.raises(_lib.SsgError):
            crit(torch.zeros(8, 16), torch.arange(2).repeat_interleave(4), 0)
    with pytest.raises(NotImplementedError):
        crit(torch.zeros(8, 16), torch.arange(2).repeat_interleave(4), 0, w=torch.ones(8))


def test_uint8_input_is_validated_on_the_host():
    """Row f5: layout errors surface as ValueError before anything is launched (checked without a GPU via the
    shape logic of EmbedPlan.forward's callers)."""
    from ssg_b200 import embed
    assert embed.IMAGENET_MEAN == (0.485, 0.456, 0.406) and embed.IMAGENET_STD == (0.229, 0.224, 0.225)
    src = open(embed.__file__).read()
    assert "ssg_embed_forward_u8" in src and "(256, 128, 3)" in src


@pytest.mark.skipif(not os.path.isdir("/root/reference/reid"), reason="/root/reference not present")
def test_unmodified_driver_imports_resolve_through_the_drop_in():
    """Route A of INTEGRATION.md: with this package AHEAD of the reference on sys.path, every import statement of the
    unmodified selftraining.py:15-28 resolves -- hot-path names to this repository, everything else to the reference's
    own files through the extended package paths (reid/_reference.py).  Runs in a subprocess (fresh sys.modules); the
    two third-party modules missing from this image (h5py, metric_learn) are stubbed as for the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, warnings
warnings.simplefilter("ignore")
sys.path[:0] = [%r, %r]
sys.path.append("/root/reference")
from oracle import refshim
refshim._install_stubs()
from reid import datasets
from reid import models
from reid.dist_metric import DistanceMetric
from reid.loss import TripletLoss, FocalLoss
from reid.trainers import Trainer, FinedTrainer, FinedTrainer2, JointTrainer2, DistillTrainer   # + semitraining.py:19, eug.py:4
from reid.evaluators import Evaluator, extract_features
from reid.utils.data import transforms as T
from reid.utils.data.preprocessor import Preprocessor
from reid.utils.data.sampler import RandomIdentitySampler
from reid.utils.logging import Logger
from reid.utils.serialization import load_checkpoint, save_checkpoint
from sklearn.cluster import DBSCAN, AffinityPropagation
from reid.rerank import *
from reid.eug import *
import reid
assert reid.__file__.startswith(%r), reid.__file__
ours = lambda o: sys.modules[o.__module__].__file__.startswith(%r)
assert ours(TripletLoss) and ours(FinedTrainer2) and ours(extract_features) and ours(re_ranking) and ours(Evaluator)
assert DBSCAN.__module__.startswith("ssg_b200")                      # shadows sklearn's (selftraining.py:27-28)
assert not ours(Trainer) and not ours(Preprocessor) and not ours(DistanceMetric)   # the reference's own
assert "market1501" in datasets.names()
import torch
fl = FocalLoss(gamma=2.0, alpha=0.25)(torch.randn(6, 2), torch.randint(0, 2, (6,)), 0)
assert fl.dim() == 0 and bool(torch.isfinite(fl))
from reid.utils import to_numpy, to_torch
assert to_torch(to_numpy(torch.ones(3))).sum() == 3
print("driver imports ok")
''' % (os.path.join(root, "self-similarity-grouping_b200"), root, os.path.join(root, "self-similarity-grouping_b200"),
       os.path.join(root, "self-similarity-grouping_b200"))
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "driver imports ok" in out.stdout, out.stderr[-2000:]


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/ssg_b200.h is the C ABI: it must compile as C99 (no torch / C++ types), and a C program that only
    includes it must link against libssg_b200.so and run its no-GPU entry points."""
    import shutil
    import subprocess
    from ssg_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "ssg_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                   check=True)
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include "ssg_b200.h"\n'
                   'int main(void) { int cin, cout, k, s; char a[64], b[64];\n'
                   '  if (ssg_version() < 100) return 1;\n'
                   '  if (ssg_embed_layer_info(0, &cin, &cout, &k, &s, a, b, 64) != SSG_OK) return 2;\n'
                   '  if (ssg_embed_layer_info(-1, 0, 0, 0, 0, 0, 0, 0) != SSG_ERR_INVALID) return 3;\n'
                   '  printf("%d %d %d %d %s %d %s\\n", cin, cout, k, s, a, ssg_embed_num_layers(), ssg_last_error());\n'
                   '  return 0; }\n')
    exe = tmp_path / "probe"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lssg_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("3 64 7 2 conv1 53 ")


def test_no_unbound_names_in_python_sources():
    """No pyflakes in this image: tools/lint_names.py flags names that are read but never bound anywhere in a file --
    the only static net under the code paths that need a GPU to execute (bench.py's GPU arm, the ctypes wrappers)."""
    import glob
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = [os.path.join(root, "bench.py"), os.path.join(root, "__graft_entry__.py")]
    for pat in ("self-similarity-grouping_b200/ssg_b200/*.py", "self-similarity-grouping_b200/reid/*.py",
                "self-similarity-grouping_b200/reid/*/*.py", "tests/*.py", "oracle/*.py", "tools/*.py"):
        files += sorted(glob.glob(os.path.join(root, pat)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "lint_names.py")] + files, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_rerank_plain_drop_in_surface():
    """Row f4: reid.rerank_plain keeps the reference's names and signatures (rerank_plain.py:27,125); without a GPU the
    CUDA-backed function raises loudly instead of falling back."""
    import inspect
    import numpy as np
    import pytest as _pt
    import torch
    from reid import rerank_plain
    sig = inspect.signature(rerank_plain.re_ranking)
    assert list(sig.parameters)[:6] == ["input_feature_source", "input_feature", "k", "lambda_value", "MemorySave", "Minibatch"]
    assert sig.parameters["k"].default == 20 and sig.parameters["lambda_value"].default == 0.1
    assert list(inspect.signature(rerank_plain.re_ranking_lh).parameters)[:5] == [
        "input_feature_source", "input_feature", "k1", "k2", "lambda_value"]
    if not torch.cuda.is_available():
        with _pt.raises(RuntimeError):
            rerank_plain.re_ranking(np.zeros((4, 8), np.float32), np.zeros((30, 8), np.float32))
