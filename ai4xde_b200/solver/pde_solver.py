"""PINNSolver: the frozen solver API of /root/reference/AI4XDE/solver/PDESolver.py:9-97
(`train`, `closure`, `train_step(lr, iterations, callbacks, eval)`, `eval`, `save`,
`plot_loss_history`) driving the fused dde.Model.  Two-phase schedule per train_step: Adam for
`iterations`, then L-BFGS until it stalls or hits its maxiter (PDESolver.py:79-85).

The reference switches DeepXDE to float64 at import (PDESolver.py:6); the fused path is float32 /
TF32 by contract (ai4xde_b200/dde/config.py) and keeps only host point sets in float64."""
import os
import time
import warnings

import numpy as np

from .. import dde

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    dde.config.set_default_float("float64")


class PINNSolver:
    def __init__(self, name, PDECase):
        self.name = name
        self.PDECase = PDECase
        self.model = dde.Model(PDECase.data, PDECase.net)
        self.error = []
        self.losshistory = None
        self.train_state = None
        self.train_cost = None

    def save(self, path=None, add_time=False):
        if path is None:
            path = f"./models/{self.name}_{self.PDECase.name}"
        path += f'_{time.strftime("%Y%m%d_%H%M%S", time.localtime())}/' if add_time else "/"
        os.makedirs(path, exist_ok=True)
        self.model.save(path)
        dde.saveplot(self.losshistory, self.train_state, issave=True, isplot=False, output_dir=path)
        np.savetxt(f"{path}{self.name}_error.txt", self.error)

    def plot_loss_history(self, axes=None, train=False, use_time=False):
        import matplotlib.pyplot as plt  # optional dependency, plotting is off the hot path

        fig, axes = (plt.subplots() if axes is None else (axes.get_figure(), axes))
        loss_train = np.sum(self.losshistory.loss_train, axis=1)
        loss_test = np.sum(self.losshistory.loss_test, axis=1)
        if use_time:
            x = np.array(self.losshistory.steps) / self.losshistory.steps[-1] * self.train_cost
            axes.set_xlabel("Time(s)")
        else:
            x = self.losshistory.steps
            axes.set_xlabel("Steps")
        if train:
            axes.semilogy(x, loss_train, label="Train loss")
            axes.semilogy(x, loss_test, label="Test loss")
        else:
            axes.semilogy(x, loss_test, label=self.name)
        axes.set_title(self.name)
        axes.set_ylabel("Loss")
        axes.legend()
        return fig, axes

    def eval(self):
        X, y = self.PDECase.get_testdata()
        y_pred = self.model.predict(X)
        error = dde.metrics.l2_relative_error(y, y_pred)
        print("L2 relative error:", error)
        self.error.append(np.array(error))

    def train_step(self, lr=1e-3, iterations=15000, callbacks=None, eval=False):
        self.PDECase.compile(self.model, "adam", lr=lr)
        self.model.train(iterations=iterations, callbacks=callbacks)
        self.PDECase.compile(self.model, "L-BFGS")
        self.losshistory, self.train_state = self.model.train()
        if eval:
            self.eval()

    def closure(self):
        self.train_step()

    def train(self, **kwargs):
        t_start = time.time()
        print(f"Model {self.name} is training...")
        self.closure(**kwargs)
        self.train_cost = time.time() - t_start
