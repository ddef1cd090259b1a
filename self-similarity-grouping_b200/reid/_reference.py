"""Locating the reference's own ``reid`` package, for the parts this drop-in does not provide.

Route A of INTEGRATION.md puts this package ahead of the reference on ``sys.path``; a regular package shadows a
same-named one completely, so everything the hot path does not cover (``reid.datasets``, ``reid.utils.data``,
``reid.dist_metric``, the other trainers and losses ...) would become unimportable.  Each package here therefore
appends the matching directory of the reference to its ``__path__``: sub-modules that exist here win, everything else
resolves to the reference's files, unmodified.  The reference is found through ``SSG_REFERENCE_ROOT`` or as the next
``reid`` package on ``sys.path``; without one the drop-in still works for the names it defines.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_cached = False
_root = None


def reference_reid_dir():
    """Directory of the reference's ``reid`` package, or None."""
    global _cached, _root
    if _cached:
        return _root
    cands = []
    env = os.environ.get("SSG_REFERENCE_ROOT")
    if env:
        cands.append(os.path.join(env, "reid"))
    for p in sys.path:
        cands.append(os.path.join(p or ".", "reid"))
    for c in cands:
        c = os.path.abspath(c)
        if c != _HERE and os.path.isfile(os.path.join(c, "__init__.py")) and os.path.isfile(os.path.join(c, "rerank.py")):
            _root = c
            break
    _cached = True
    return _root


def extend_path(package_path, *sub):
    """Append <reference>/reid/<sub...> to a package's ``__path__`` (no-op when there is no reference)."""
    root = reference_reid_dir()
    if root is None:
        return False
    d = os.path.join(root, *sub)
    if os.path.isdir(d) and d not in package_path:
        package_path.append(d)
        return True
    return False


def load_shadowed(relpath, alias):
    """Import the reference's copy of a module this package shadows (e.g. ``trainers.py``) under ``reid.<alias>``;
    its relative imports resolve inside the merged package.  Returns the module or None."""
    root = reference_reid_dir()
    if root is None:
        return None
    path = os.path.join(root, relpath)
    if not os.path.isfile(path):
        return None
    name = "reid." + alias
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules[name]
        raise
    return mod
