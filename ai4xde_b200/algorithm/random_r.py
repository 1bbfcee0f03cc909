"""Random_R: periodic uniform resampling (/root/reference/AI4XDE/algorithm/Random_R.py:6-17)."""
from .. import dde
from ..solver import PINNSolver


class Random_R(PINNSolver):
    def __init__(self, PDECase, P):
        super().__init__(name=f"Random_R_P_{P}", PDECase=PDECase)
        self.P = P
        self.resampler = dde.callbacks.PDEPointResampler(period=self.P)

    def closure(self):
        self.train_step(callbacks=[self.resampler])
