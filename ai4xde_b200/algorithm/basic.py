"""Algorithms without arithmetic of their own on the hot path: they only decide how often
`train_step` runs and what happens to the point set in between.

  PINN      vanilla schedule                      (/root/reference/AI4XDE/algorithm/PINN.py:5-15)
  Random_R  uniform re-draw every P iterations    (/root/reference/AI4XDE/algorithm/Random_R.py:6-17)
  RAR_G     greedy residual-based refinement      (/root/reference/AI4XDE/algorithm/RAR_G.py:7-29)

RAR_G's selection stays NumPy as in the reference -- including its behaviour of adding ALL 100 000
candidates in `argpartition` order (the partition result is never sliced there).
Only its residual probe runs on the fused path (forward-only jets through `model.predict`)."""
import numpy as np

from .. import dde
from ..solver import PINNSolver


class PINN(PINNSolver):
    def __init__(self, PDECase, iteration=15000):
        self.iteration = iteration
        super().__init__("PINN", PDECase)

    def closure(self):
        self.train_step(iterations=self.iteration)


class Random_R(PINNSolver):
    def __init__(self, PDECase, P):
        self.P = P
        self.resampler = dde.callbacks.PDEPointResampler(period=P)
        super().__init__(f"Random_R_P_{P}", PDECase)

    def closure(self):
        self.train_step(callbacks=[self.resampler])


class RAR_G(PINNSolver):
    CANDIDATES = 100000

    def __init__(self, PDECase, iter=100, add_num=1):
        dde.optimizers.config.set_LBFGS_options(maxiter=1000)
        self.iter, self.add_num = iter, add_num
        super().__init__("RAR_G", PDECase)

    def refine(self):
        case = self.PDECase
        candidates = case.geomtime.random_points(self.CANDIDATES)
        residual = np.abs(self.model.predict(candidates, operator=case.pde))[:, 0]
        order = np.argpartition(residual, self.add_num)
        case.data.add_anchors(candidates[order])

    def closure(self):
        self.train_step()
        for _ in range(self.iter):
            self.refine()
            self.train_step(iterations=1000)
