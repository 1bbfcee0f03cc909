"""Import the UNMODIFIED reference package (/root/reference/AI4XDE) on top of ddelite.

Container-only helper (the GPU box has no /root/reference): used by
tests/golden/make_golden.py to generate fixtures and by the `not gpu` tests that
cross-check oracle/cases.py against the reference's own closures when the tree is present.
matplotlib and scikit-optimize are absent from the image, so inert stand-ins are
registered for the names the reference imports at module load
(/root/reference/AI4XDE/utils/Visualization.py:2, algorithm/HPO.py:5-8, Uniform.py:1).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "AI4XDE"))


class _Inert(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        mod = _Inert(self.__name__ + "." + name)
        setattr(self, name, mod)
        return mod

    def __call__(self, *a, **k):
        return self


def _stub(name):
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
    except Exception:
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            sub = ".".join(parts[:i])
            if sub not in sys.modules:
                sys.modules[sub] = _Inert(sub)
                sys.modules[sub].__path__ = []


def load():
    """Return the reference `AI4XDE` package running on oracle.ddelite (float64 default,
    because /root/reference/AI4XDE/solver/PDESolver.py:6 switches it at import)."""
    if not available():
        raise RuntimeError("reference tree not present")
    from . import ddelite

    ddelite.install_as_deepxde()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.tri", "matplotlib.colors",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1", "skopt", "skopt.space",
                 "skopt.utils", "skopt.plots", "skopt.sampler"):
        _stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("AI4XDE")
