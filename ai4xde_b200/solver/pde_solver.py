"""PINNSolver -- the solver front end the algorithms subclass.

API contract kept from /root/reference/AI4XDE/solver/PDESolver.py:9-97 (what user scripts and the
algorithm classes call): constructor `(name, PDECase)`; attributes `name, PDECase, model, error,
losshistory, train_state, train_cost`; methods `train(**kw)`, `closure()`, `train_step(lr, iterations,
callbacks, eval)`, `eval()`, `save(path, add_time)`, `plot_loss_history(axes, train, use_time)`.
Behaviour kept: one `train_step` = an Adam phase of `iterations` steps followed by an L-BFGS phase that
runs until it stalls or reaches its `maxiter`; `eval` appends the L2 relative error on the case's test
data; `train` wall-clocks `closure`.

The reference flips DeepXDE to float64 when this module is imported (PDESolver.py:6).  The fused path
computes in float32 / TF32 by contract, so only the HOST point sets follow that switch
(ai4xde_b200/dde/config.py); the one-time notice is silenced here because it is the default."""
import os
import time
import warnings
from dataclasses import dataclass, field
from typing import Any, List, Optional

import numpy as np

from .. import dde
from ..utils.visualization import plot_loss_curves

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    dde.config.set_default_float("float64")

# (optimizer name, needs lr / iteration count) in the order a train_step runs them
_PHASES = (("adam", True), ("L-BFGS", False))


def _output_directory(solver_name, case_name, path, add_time):
    base = path if path is not None else os.path.join(".", "models", f"{solver_name}_{case_name}")
    if add_time:
        base = f"{base}_{time.strftime('%Y%m%d_%H%M%S', time.localtime())}"
    base += "/"
    os.makedirs(base, exist_ok=True)
    return base


@dataclass
class _Record:
    """What a run leaves behind; exposed on the solver under the reference's attribute names."""
    error: List[Any] = field(default_factory=list)
    losshistory: Optional[Any] = None
    train_state: Optional[Any] = None
    train_cost: Optional[float] = None


class PINNSolver:
    def __init__(self, name, PDECase):
        self.name = name
        self.PDECase = PDECase
        self.model = dde.Model(PDECase.data, PDECase.net)
        self._rec = _Record()

    # ---- the reference's public attributes, backed by the record ------------------------------
    error = property(lambda self: self._rec.error, lambda self, v: setattr(self._rec, "error", v))
    losshistory = property(lambda self: self._rec.losshistory, lambda self, v: setattr(self._rec, "losshistory", v))
    train_state = property(lambda self: self._rec.train_state, lambda self, v: setattr(self._rec, "train_state", v))
    train_cost = property(lambda self: self._rec.train_cost, lambda self, v: setattr(self._rec, "train_cost", v))

    # ---- training -----------------------------------------------------------------------------
    def _run_phase(self, optimizer, lr, iterations, callbacks):
        if optimizer == "adam":
            self.PDECase.compile(self.model, optimizer, lr=lr)
            return self.model.train(iterations=iterations, callbacks=callbacks)
        self.PDECase.compile(self.model, optimizer)
        return self.model.train()

    def train_step(self, lr=1e-3, iterations=15000, callbacks=None, eval=False):
        result = None
        for optimizer, first_order in _PHASES:
            result = self._run_phase(optimizer, lr if first_order else None,
                                     iterations if first_order else None, callbacks if first_order else None)
        self._rec.losshistory, self._rec.train_state = result
        if eval:
            self.eval()

    def closure(self):
        self.train_step()

    def train(self, **kwargs):
        print(f"Model {self.name} is training...")
        started = time.time()
        self.closure(**kwargs)
        self._rec.train_cost = time.time() - started

    # ---- evaluation / persistence ---------------------------------------------------------------
    def eval(self):
        X, y_true = self.PDECase.get_testdata()
        err = dde.metrics.l2_relative_error(y_true, self.model.predict(X))
        print("L2 relative error:", err)
        self._rec.error.append(np.array(err))

    def save(self, path=None, add_time=False):
        out = _output_directory(self.name, self.PDECase.name, path, add_time)
        self.model.save(out)
        dde.saveplot(self._rec.losshistory, self._rec.train_state, issave=True, isplot=False, output_dir=out)
        np.savetxt(f"{out}{self.name}_error.txt", self._rec.error)

    def plot_loss_history(self, axes=None, train=False, use_time=False):
        return plot_loss_curves(self.name, self._rec.losshistory, self._rec.train_cost, axes=axes, train=train,
                                use_time=use_time)
