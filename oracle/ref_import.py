"""TEST INFRASTRUCTURE ONLY — imports the real pyinterpolate from /root/reference/src (build container) or from
oracle/_ref (the git-ignored copy `oracle/make_ref.py` makes so that the reference travels to the GPU box).

The reference's top-level ``__init__`` pulls in geopandas / shapely / prettytable /
matplotlib / dask, none of which exist in this image.  The hot-path functions only use
them for ``isinstance`` checks, plotting and printing, so placeholder modules are enough
(SURVEY.md §8c).  This module exists so ``tests/golden/make_golden.py`` can run the
unmodified reference in the build container and ``bench.py --impl reference`` can time it on the GPU box's host
cores; it is never imported by the product package.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import os

REFERENCE_SRC = "/root/reference/src"
if not os.path.isdir(REFERENCE_SRC):
    REFERENCE_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_STUBBED = ("geopandas", "shapely", "prettytable", "matplotlib", "dask")


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_reference():
    """Return the real ``pyinterpolate`` package (raises if /root/reference is absent)."""
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "pyinterpolate")):
        raise RuntimeError("reference sources are not available on this machine")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import pyinterpolate
    return pyinterpolate
