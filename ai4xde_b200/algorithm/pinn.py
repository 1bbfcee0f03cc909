"""Vanilla PINN (/root/reference/AI4XDE/algorithm/PINN.py:5-15)."""
from ..solver import PINNSolver


class PINN(PINNSolver):
    def __init__(self, PDECase, iteration=15000):
        super().__init__(name="PINN", PDECase=PDECase)
        self.iteration = iteration

    def closure(self):
        self.train_step(iterations=self.iteration)
