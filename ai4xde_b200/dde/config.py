"""Numeric policy of the fused path.

The reference flips DeepXDE to float64 at import (/root/reference/AI4XDE/solver/PDESolver.py:6).
The fused kernels compute in float32 (TF32 on the tensor-core engine) by contract, so
`set_default_float("float64")` is accepted for source compatibility, keeps host-side NumPy point
sets in float64 (the resamplers' RNG stream and index selection depend on that), and leaves
device tensors in float32 -- said once, loudly."""
import warnings

import numpy as np
import torch

import os

_state = {"host": np.float32, "warned": False, "engine": "auto", "graphs": "auto",
          "graph_max_rows": int(os.environ.get("AI4XDE_B200_GRAPH_MAX_ROWS", "300000"))}


def real(package):
    if package is np:
        return _state["host"]
    if package is torch:
        return torch.float32
    raise ValueError(package)


def set_default_float(value):
    if value == "float64":
        _state["host"] = np.float64
        if not _state["warned"]:
            warnings.warn(
                "ai4xde_b200: float64 requested; host point sets stay float64, the fused B200 path "
                "computes in float32/TF32 (master weights float32).",
                stacklevel=2,
            )
            _state["warned"] = True
    elif value == "float32":
        _state["host"] = np.float32
    else:
        raise ValueError(f"{value} not supported")


def default_float():
    return "float64" if _state["host"] is np.float64 else "float32"


def set_engine(name):
    """'auto' | 'fp32' | 'tf32' -- arithmetic engine of the jet kernels (include/ai4xde_b200.h)."""
    if name not in ("auto", "fp32", "tf32"):
        raise ValueError(name)
    _state["engine"] = name


def engine():
    return _state.get("engine_scope") or _state["engine"]


def set_lbfgs_engine(name, max_rows=None):
    """Engine of the L-BFGS phase when the engine is 'auto'.  'fp32' (default): point sets of up to `max_rows` rows
    (default 2^17) run that phase on the FP32 kernels -- L-BFGS builds its curvature pairs from DIFFERENCES of
    successive gradients, and the 1e-3 rounding of TF32 operands stalls it (measured on the reference-default
    Helmholtz net 2-[150]x3-1 sin, Adam 5000 + L-BFGS 3000: final L2 error 4.0 % with TF32 gradients, 1.5 % with
    FP32 gradients in the L-BFGS phase); larger sets stay on the tensor cores.  'auto': no special case."""
    if name not in ("fp32", "auto"):
        raise ValueError(name)
    _state["lbfgs_engine"] = name
    if max_rows is not None:
        _state["lbfgs_fp32_max_rows"] = int(max_rows)


def lbfgs_engine():
    return _state.get("lbfgs_engine", "fp32"), _state.get("lbfgs_fp32_max_rows", 1 << 17)


class engine_scope:
    """with engine_scope('fp32'): ... -- overrides the engine for the duration (None: no override)"""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.prev = _state.get("engine_scope")
        if self.name is not None:
            _state["engine_scope"] = self.name

    def __exit__(self, *exc):
        _state["engine_scope"] = self.prev


def set_cuda_graphs(mode, max_rows=None):
    """'auto' (capture the Adam step of small problems, which are launch-bound), True, or False."""
    if mode not in ("auto", True, False):
        raise ValueError(mode)
    _state["graphs"] = mode
    if max_rows is not None:
        _state["graph_max_rows"] = int(max_rows)


def cuda_graphs():
    return _state["graphs"], _state["graph_max_rows"]


def cuda_graph_max_bytes():
    """'auto' also captures the Adam step of larger point sets when the layers the backward keeps (the stash) stay below
    this many bytes -- narrow nets on millions of points are still launch-bound on the host side (a 2-[20]x3-1 step at
    2^20 points is 1.3 ms of kernels and ~20 launches), wide ones are not and would only pin tens of GB in the graph's
    private pool."""
    return int(_state.get("graph_max_bytes", 2 << 30))


def set_cuda_graph_max_bytes(nbytes):
    _state["graph_max_bytes"] = int(nbytes)


def set_residual_programs(on):
    """True (default): the loss side of a train step -- pde / BC algebra, per-term reductions, their reverse pass --
    runs as one traced program per point (ai4xde_b200/residual.py) whenever the case's code can be traced and the
    program reproduces the autograd path on the live batch; False: always torch.autograd over the jet leaves."""
    _state["residual_programs"] = bool(on)


def residual_programs():
    return _state.get("residual_programs", True)


def device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def set_random_seed(seed):
    import random

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
