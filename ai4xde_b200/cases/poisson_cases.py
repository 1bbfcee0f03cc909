"""1-D Poisson family on the fused path  (/root/reference/AI4XDE/cases/PoissonCases.py)."""
from abc import abstractmethod

import numpy as np

from .. import dde
from ..dde import backend as bkd
from ..utils import Visualization_1D
from .base import PDECases


class PoissonCase1D(PDECases):
    """u_xx = f(x) template  (PoissonCases.py:9-57)."""

    def __init__(self, name, NumDomain=2000, Interval=None, use_output_transform=False,
                 layer_size=[2] + [32] * 3 + [1], activation="tanh", initializer="Glorot uniform",
                 metrics=["l2 relative error"], loss_weights=None, external_trainable_variables=None):
        self.Interval = Interval
        super().__init__(name=name, NumDomain=NumDomain, use_output_transform=use_output_transform,
                         layer_size=layer_size, activation=activation, initializer=initializer, metrics=metrics,
                         loss_weights=loss_weights, external_trainable_variables=external_trainable_variables,
                         visualization=Visualization_1D(x_label="x", y_label="y"))

    @abstractmethod
    def func(self, x):
        ...

    def gen_geomtime(self):
        return dde.geometry.Interval(self.Interval[0], self.Interval[1])

    def gen_pde(self):
        def pde(x, y):
            return dde.grad.hessian(y, x) - self.func(x)

        return pde


class Poisson_1D_Dirichlet(PoissonCase1D):
    """-u'' = pi^2 sin(pi x) on [-1, 1], u(+-1) = 0  (PoissonCases.py:133-175)."""

    def __init__(self, NumDomain=16, layer_size=[1] + [50] * 3 + [1], activation="tanh",
                 initializer="Glorot uniform"):
        super().__init__(name="Poisson equation in 1D with Dirichlet boundary conditions", NumDomain=NumDomain,
                         Interval=[-1, 1], layer_size=layer_size, activation=activation, initializer=initializer)

    def func(self, x):
        return -np.pi**2 * bkd.sin(np.pi * x)

    def sol(self, x):
        return np.sin(np.pi * x)

    def gen_data(self):
        bc = dde.icbc.DirichletBC(self.geomtime, self.sol, lambda _, on_boundary: on_boundary)
        return dde.data.PDE(self.geomtime, self.pde, bc, self.NumDomain, 2, solution=self.sol, num_test=100)


class Poisson_1D_Fourier_Net(PoissonCase1D):
    """Multiscale 1-D Poisson: u = sin(2 pi x) + 0.1 sin(50 pi x)  (PoissonCases.py:419-493).
    The reference builds an MsFFN for it; the multiscale Fourier net is not a plain FNN and is not in
    the fused kernels yet, so this class uses dde.nn.FNN with the given layer_size (the FBPINN
    BASELINE cfg5 runs 1-[64]x3-1 sub-nets on it).  Learning-rate decay as in the reference."""

    def __init__(self, NumDomain=1280, layer_size=[1] + [100] * 3 + [1], activation="tanh",
                 initializer="Glorot uniform", num_boundary=2):
        # [ext] num_boundary: the reference passes none (PoissonCases.py:485-493), which leaves the
        # Dirichlet term with zero rows and a NaN mean; 2 is what the DeepXDE demo it cites uses.
        self.A, self.B = 2, 50
        self.num_boundary = num_boundary
        super().__init__(name="Poisson equation in 1D with Multi-scale Fourier feature networks",
                         NumDomain=NumDomain, Interval=[0, 1], use_output_transform=True, layer_size=layer_size,
                         activation=activation, initializer=initializer)

    def gen_compile(self, metrics=None, loss_weights=None, external_trainable_variables=None):
        def compile(model, optimizer, lr=None, loss="MSE", decay=("inverse time", 2000, 0.9)):
            model.compile(optimizer, lr, loss, metrics, decay, loss_weights, external_trainable_variables)

        return compile

    def func(self, x):
        a, b = np.pi * self.A, np.pi * self.B
        return -(a**2) * bkd.sin(a * x) - 0.1 * b**2 * bkd.sin(b * x)

    def sol(self, x):
        return np.sin(np.pi * self.A * x) + 0.1 * np.sin(np.pi * self.B * x)

    def gen_net(self, layer_size, activation, initializer):
        return dde.nn.FNN(layer_size, activation, initializer)

    def gen_data(self):
        bc = dde.icbc.DirichletBC(self.geomtime, self.sol, lambda _, on_boundary: on_boundary)
        return dde.data.PDE(self.geomtime, self.pde, bc, self.NumDomain, self.num_boundary,
                            train_distribution="pseudo", solution=self.sol, num_test=10000)


def multiscale_fbpinn_case(NumDomain=1280, layer_size=(1, 64, 64, 64, 1), activation="tanh",
                           initializer="Glorot uniform"):
    """The parent case of BASELINE config 5: the 1-D multiscale Poisson problem (forcing and solution of
    PoissonCases.py:469-476) on a dde.nn.FNN built by the BASE `gen_net` with `use_output_transform=True`, so that
    FBPINN's sub-cases get their sigmoid window wired (FBPINN.py:35-42 -- with the reference's own
    Poisson_1D_Fourier_Net the flag is set but its `gen_net` override never applies the transform, SURVEY.md
    App. C #8 / #11).  FBPINN builds every sub-case with the CLASS defaults (FBPINN.py:42 `PDE()`), so the
    requested point count and net shape are baked into a fresh class; returns an instance of it."""
    nd, ls, act, ini = int(NumDomain), [int(w) for w in layer_size], activation, initializer

    class Poisson_1D_Multiscale(PoissonCase1D):
        def __init__(self, NumDomain=nd, layer_size=ls, activation=act, initializer=ini):
            self.A, self.B = 2, 50
            super().__init__(name="Poisson equation in 1D, multiscale forcing (FBPINN parent)", NumDomain=NumDomain,
                             Interval=[0, 1], use_output_transform=True, layer_size=layer_size,
                             activation=activation, initializer=initializer)

        def func(self, x):
            a, b = np.pi * self.A, np.pi * self.B
            return -(a**2) * bkd.sin(a * x) - 0.1 * b**2 * bkd.sin(b * x)

        def sol(self, x):
            return np.sin(np.pi * self.A * x) + 0.1 * np.sin(np.pi * self.B * x)

        def gen_data(self):
            bc = dde.icbc.DirichletBC(self.geomtime, self.sol, lambda _, on_boundary: on_boundary)
            return dde.data.PDE(self.geomtime, self.pde, bc, self.NumDomain, 2, train_distribution="pseudo",
                                solution=self.sol, num_test=1000)

    return Poisson_1D_Multiscale()
