"""RAR_G: greedy residual-based refinement (/root/reference/AI4XDE/algorithm/RAR_G.py:7-29).
Only the residual probe is on the hot path; the selection stays NumPy as in the reference --
including its behaviour of adding ALL candidates in argpartition order (the partition result is
not sliced there)."""
import numpy as np

from .. import dde
from ..solver import PINNSolver


class RAR_G(PINNSolver):
    def __init__(self, PDECase, iter=100, add_num=1):
        dde.optimizers.config.set_LBFGS_options(maxiter=1000)
        super().__init__(name="RAR_G", PDECase=PDECase)
        self.iter, self.add_num = iter, add_num

    def closure(self):
        self.train_step()
        for _ in range(self.iter):
            X = self.PDECase.geomtime.random_points(100000)
            Y = np.abs(self.model.predict(X, operator=self.PDECase.pde))[:, 0]
            ids = np.argpartition(Y, self.add_num)
            self.PDECase.data.add_anchors(X[ids])
            self.train_step(iterations=1000)
