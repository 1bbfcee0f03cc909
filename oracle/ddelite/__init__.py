"""ddelite: CPU restatement of the DeepXDE (PyTorch backend) names AI4XDE imports
(SURVEY.md Appendix E).  Test infrastructure -- see oracle/__init__.py."""
from . import backend, callbacks, config, data, geometry, gradients, icbc, losses, metrics
from . import nn, optimizers, utils
from . import gradients as grad
from . import nn as maps
from .backend import Variable
from .icbc import (
    DirichletBC,
    IC,
    NeumannBC,
    OperatorBC,
    PeriodicBC,
    PointSetBC,
    PointSetOperatorBC,
    RobinBC,
)
from .model import LossHistory, Model, TrainState, saveplot

__all__ = [
    "backend", "callbacks", "config", "data", "geometry", "grad", "gradients", "icbc", "losses",
    "maps", "metrics", "nn", "optimizers", "utils", "Model", "Variable", "saveplot",
]


def install_as_deepxde():
    """Register this package under the name ``deepxde`` so UNMODIFIED reference sources
    (``import deepxde as dde``; ``import deepxde.backend as bkd``) run on the oracle.
    Used only by tests/golden/make_golden.py and the reference-import tests."""
    import sys

    me = sys.modules[__name__]
    sys.modules["deepxde"] = me
    for name in ("backend", "callbacks", "config", "data", "geometry", "gradients", "icbc",
                 "losses", "metrics", "nn", "optimizers", "utils"):
        sys.modules["deepxde." + name] = getattr(me, name)
    sys.modules["deepxde.optimizers.config"] = optimizers.config
    return me
