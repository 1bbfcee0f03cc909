"""ai4xde_b200 -- B200-native PINN residual hot path behind AI4XDE's Solver / algorithm / case API.

Layout: `csrc/` CUDA kernels + C ABI (include/ai4xde_b200.h), `_lib` ctypes binding, `jet` autograd
bridge, `dde` DeepXDE-compatible namespace, `solver` / `algorithm` / `cases` mirrors of the
reference packages.  There is no CPU fallback: importing works anywhere, computing needs the
built library and a CUDA device.
"""
import sys

from . import _lib, jet  # noqa: F401
from . import dde  # noqa: F401

__version__ = "0.1.0"


def install_as_deepxde():
    """Register ai4xde_b200.dde under the module name `deepxde` (and its submodules) so that
    unmodified reference sources (`import deepxde as dde`, `import deepxde.backend as bkd`,
    `from deepxde import gradients`) run on the fused path."""
    sys.modules["deepxde"] = dde
    for name in ("backend", "callbacks", "config", "data", "geometry", "gradients", "icbc", "losses",
                 "metrics", "nn", "optimizers", "utils"):
        sys.modules["deepxde." + name] = getattr(dde, name)
    sys.modules["deepxde.optimizers.config"] = dde.optimizers.config
    return dde
