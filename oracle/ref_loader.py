"""Import the REAL reference `bayesml.gaussianmixture` from /root/reference (build container only).

TEST INFRASTRUCTURE. /root/reference does not exist on the GPU box; this loader is used only by
`tests/golden/make_golden.py` (fixture generation) and by the CPU tests that are skipped when the
reference tree is absent. matplotlib is not installed, and `bayesml/__init__.py` imports every model
(metatree needs pyplot at import), so a bare package object is registered and only the
gaussianmixture sub-package (+ base, _check, _exceptions) is executed.
"""
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("BAYESML_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bayesml", "gaussianmixture"))


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
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "bayesml")]
        sys.modules["bayesml"] = pkg
    import importlib
    return importlib.import_module("bayesml." + name)
