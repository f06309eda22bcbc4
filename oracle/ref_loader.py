"""Import the REAL reference `bayesml.gaussianmixture`: from /root/reference in the build container, else from the
copy `oracle/build_ref.py` made under `oracle/_ref/` (git-ignored; it travels to the GPU box with the snapshot).

TEST INFRASTRUCTURE: used by `tests/golden/make_golden.py` (fixture generation), by the CPU tests that compare the
oracle with the reference, and by `bench.py`'s CPU arms (`--impl reference`, `cpu_baseline`). matplotlib is not installed, and `bayesml/__init__.py` imports every model
(metatree needs pyplot at import), so a bare package object is registered and only the
gaussianmixture sub-package (+ base, _check, _exceptions) is executed.
"""
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))


_ARCHIVE = os.path.join(_HERE, "_ref", "bayesml_ref.zip")


def _find_root():
    """A directory holding the reference's `bayesml/` tree, or the archive oracle/build_ref.py packed from it."""
    for cand in (os.environ.get("BAYESML_REFERENCE_ROOT"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "bayesml", "gaussianmixture")):
            return cand
    if os.path.exists(_ARCHIVE):
        return _ARCHIVE
    return os.environ.get("BAYESML_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def reference_is_copy() -> bool:
    """True when the reference is the oracle/_ref archive (GPU box) rather than the /root/reference tree."""
    return REFERENCE_ROOT == _ARCHIVE


def reference_available() -> bool:
    return reference_is_copy() or os.path.isdir(os.path.join(REFERENCE_ROOT, "bayesml", "gaussianmixture"))


def load_reference_gaussianmixture():
    """Returns the reference's `bayesml.gaussianmixture` module (GenModel, LearnModel)."""
    return load_reference_module("gaussianmixture")


def load_reference_module(name):
    """Returns the reference's `bayesml.<name>` sub-package (gaussianmixture, hiddenmarkovnormal, multivariate_normal)."""
    if not reference_available():
        raise ImportError(f"reference tree not found under {REFERENCE_ROOT}")
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        if mod not in sys.modules:
            sys.modules[mod] = mock.MagicMock(name=mod)
    sys.dont_write_bytecode = True  # the reference mount is read-only
    if "bayesml" not in sys.modules:
        pkg = types.ModuleType("bayesml")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "bayesml")]     # a directory, or a path inside the zip archive (zipimport)
        sys.modules["bayesml"] = pkg
    import importlib
    return importlib.import_module("bayesml." + name)
