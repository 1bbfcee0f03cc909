"""`deepxde`-compatible namespace of the fused B200 path: the names AI4XDE imports
(SURVEY.md Appendix E), same call signatures, fused kernels underneath.

    import ai4xde_b200.dde as dde        # what ai4xde_b200's own solver/algorithm/cases use
    ai4xde_b200.install_as_deepxde()     # or: make `import deepxde` resolve here, so an
                                         # UNMODIFIED AI4XDE checkout runs on the fused path
"""
from . import backend, callbacks, config, data, geometry, gradients, icbc, losses, metrics
from . import nn, optimizers, utils
from . import gradients as grad
from . import nn as maps
from .backend import Variable
from .icbc import (DirichletBC, IC, NeumannBC, OperatorBC, PeriodicBC, PointSetBC,
                   PointSetOperatorBC, RobinBC)
from .model import LossHistory, Model, TrainState, saveplot

__all__ = ["backend", "callbacks", "config", "data", "geometry", "grad", "gradients", "icbc",
           "losses", "maps", "metrics", "nn", "optimizers", "utils", "Model", "Variable", "saveplot"]
