"""Import the UNMODIFIED reference (``/root/reference``) as a CPU oracle.  Test infrastructure only.

The reference's ``import reid`` fails in this image because ``reid/feature_extraction/database.py:3``
imports h5py and ``reid/metric_learning/__init__.py:3-4`` imports metric_learn (both absent and both
unused on the pseudo-label path).  We register empty stand-ins for those two third-party modules and
then import the reference package untouched.  Nothing here exists on the GPU box (no /root/reference
there): callers must check :func:`available` first.

Variants:
  * O-ref : reference functions as they are (fp16 storage, unstable argsort) — statistical checks only.
  * O-f32 : the same function objects executed with the module-global ``np`` of ``reid.rerank``
            replaced by a proxy mapping ``float16 -> float32`` and ``argsort -> kind='stable'``;
            the reference source is untouched.  This is the bit-/1e-4 parity target.
"""
import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get("SSG_REFERENCE_ROOT", "/root/reference")
_ref_pkg = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "reid", "rerank.py"))


def _install_stubs():
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if "metric_learn" not in sys.modules:
        ml = types.ModuleType("metric_learn")
        for name in ("ITML_Supervised", "LMNN", "LSML_Supervised", "SDML_Supervised", "NCA", "LFDA",
                     "RCA_Supervised"):
            setattr(ml, name, type(name, (), {}))
        bm = types.ModuleType("metric_learn.base_metric")
        bm.BaseMetricLearner = type("BaseMetricLearner", (), {})
        ml.base_metric = bm
        sys.modules["metric_learn"] = ml
        sys.modules["metric_learn.base_metric"] = bm


@contextlib.contextmanager
def _reference_on_path():
    """Temporarily make ``reid`` resolve to the reference (our own drop-in package has the same name)."""
    saved = {k: v for k, v in sys.modules.items() if k == "reid" or k.startswith("reid.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "reid" or k.startswith("reid.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def load_reference():
    """Return a namespace with the reference modules (imported once, detached from sys.modules)."""
    global _ref_pkg
    if _ref_pkg is not None:
        return _ref_pkg
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    _install_stubs()
    import warnings
    with _reference_on_path(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import reid  # noqa: F401  (the reference's package)
        import reid.rerank as rerank
        import reid.rerank_initial as rerank_initial
        import reid.evaluators as evaluators
        import reid.models as models
        import reid.feature_extraction as feature_extraction
        ns = types.SimpleNamespace(reid=reid, rerank=rerank, rerank_initial=rerank_initial,
                                   evaluators=evaluators, models=models,
                                   feature_extraction=feature_extraction)
    _ref_pkg = ns
    return ns


class _NpF32Stable(object):
    """numpy proxy: float16 -> float32, argsort/argpartition -> stable full argsort."""

    def __init__(self, np):
        self._np = np
        self.float16 = np.float32

    def __getattr__(self, name):
        return getattr(self._np, name)

    def argsort(self, a, *args, **kw):
        kw.setdefault("kind", "stable")
        return self._np.argsort(a, *args, **kw)

    def argpartition(self, a, kth, *args, **kw):
        # rerank_initial.py:52 only relies on positions 0..k1 being sorted; a stable full argsort is a
        # valid (and deterministic) instance of that contract.
        return self._np.argsort(a, kind="stable")


@contextlib.contextmanager
def f32_stable(module):
    import numpy
    saved = module.np
    module.np = _NpF32Stable(numpy)
    try:
        yield
    finally:
        module.np = saved


@contextlib.contextmanager
def f32_stable_globals(func):
    """The same substitution for whichever module object `func` was defined in (patches func.__globals__['np'])."""
    import numpy
    g = func.__globals__
    saved = g["np"]
    g["np"] = _NpF32Stable(numpy)
    try:
        yield
    finally:
        g["np"] = saved


def ref_re_ranking(src, tgt, mode="f32", quiet=True, **kw):
    """reid/rerank.py:27 re_ranking of the unmodified reference.  mode: 'ref' (fp16) | 'f32' (O-f32)."""
    ref = load_reference()
    ctx = f32_stable(ref.rerank) if mode == "f32" else contextlib.nullcontext()
    out = open(os.devnull, "w") if quiet else sys.stdout
    with ctx, contextlib.redirect_stdout(out):
        return ref.rerank.re_ranking(src, tgt, **kw)


def ref_re_ranking_init(q_g, q_q, g_g, stable=True, **kw):
    """reid/rerank_initial.py:40 re_ranking_init of the unmodified reference."""
    ref = load_reference()
    ctx = f32_stable(ref.rerank_initial) if stable else contextlib.nullcontext()
    with ctx:
        return ref.rerank_initial.re_ranking_init(q_g, q_q, g_g, **kw)
