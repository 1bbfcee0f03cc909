"""Initial / boundary conditions restated from DeepXDE semantics."""
import numbers

import numpy as np
import torch

from . import config, gradients as grad, utils


def _npfunc_range(func):
    """DeepXDE's npfunc_range_autocache without the cache: f(X, beg, end, aux) = func(X[beg:end])."""

    def wrapper(X, beg, end, _=None):
        return func(X[beg:end])

    return wrapper


class BC:
    def __init__(self, geom, on_boundary, component):
        self.geom = geom
        self.on_boundary = lambda x, on: np.array([on_boundary(x[i], on[i]) for i in range(len(x))])
        self.component = component
        self.boundary_normal = _npfunc_range(utils.return_tensor(self.geom.boundary_normal))

    def filter(self, X):
        return X[self.on_boundary(X, self.geom.on_boundary(X))]

    def collocation_points(self, X):
        return self.filter(X)

    def normal_derivative(self, X, inputs, outputs, beg, end):
        dydx = grad.jacobian(outputs, inputs, i=self.component, j=None)[beg:end]
        n = self.boundary_normal(X, beg, end, None)
        return torch.sum(dydx * n, 1, keepdim=True)

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        raise NotImplementedError


class DirichletBC(BC):
    def __init__(self, geom, func, on_boundary, component=0):
        super().__init__(geom, on_boundary, component)
        self.func = _npfunc_range(utils.return_tensor(func))

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        values = self.func(X, beg, end, aux_var)
        if values.dim() == 2 and values.shape[1] != 1:
            raise RuntimeError("DirichletBC function should return an array of shape N by 1")
        return outputs[beg:end, self.component : self.component + 1] - values


class NeumannBC(BC):
    def __init__(self, geom, func, on_boundary, component=0):
        super().__init__(geom, on_boundary, component)
        self.func = _npfunc_range(utils.return_tensor(func))

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        values = self.func(X, beg, end, aux_var)
        return self.normal_derivative(X, inputs, outputs, beg, end) - values


class RobinBC(BC):
    def __init__(self, geom, func, on_boundary, component=0):
        super().__init__(geom, on_boundary, component)
        self.func = func

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        return self.normal_derivative(X, inputs, outputs, beg, end) - self.func(
            X[beg:end], outputs[beg:end]
        )


class PeriodicBC(BC):
    def __init__(self, geom, component_x, on_boundary, derivative_order=0, component=0):
        super().__init__(geom, on_boundary, component)
        self.component_x = component_x
        self.derivative_order = derivative_order
        if derivative_order > 1:
            raise NotImplementedError("PeriodicBC only supports derivative_order 0 or 1.")

    def collocation_points(self, X):
        X1 = self.filter(X)
        X2 = self.geom.periodic_point(X1, self.component_x)
        return np.vstack((X1, X2))

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        mid = beg + (end - beg) // 2
        if self.derivative_order == 0:
            yleft = outputs[beg:mid, self.component : self.component + 1]
            yright = outputs[mid:end, self.component : self.component + 1]
        else:
            dydx = grad.jacobian(outputs, inputs, i=self.component, j=self.component_x)
            yleft = dydx[beg:mid]
            yright = dydx[mid:end]
        return yleft - yright


class OperatorBC(BC):
    def __init__(self, geom, func, on_boundary):
        super().__init__(geom, on_boundary, 0)
        self.func = func

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        return self.func(inputs, outputs, X)[beg:end]


class IC:
    def __init__(self, geom, func, on_initial, component=0):
        self.geom = geom
        self.func = _npfunc_range(utils.return_tensor(func))
        self.on_initial = lambda x, on: np.array([on_initial(x[i], on[i]) for i in range(len(x))])
        self.component = component

    def filter(self, X):
        return X[self.on_initial(X, self.geom.on_initial(X))]

    def collocation_points(self, X):
        return self.filter(X)

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        values = self.func(X, beg, end, aux_var)
        if values.dim() == 2 and values.shape[1] != 1:
            raise RuntimeError("IC function should return an array of shape N by 1")
        return outputs[beg:end, self.component : self.component + 1] - values


class PointSetBC:
    def __init__(self, points, values, component=0, batch_size=None, shuffle=True):
        self.points = np.array(points, dtype=config.real(np))
        self.values = torch.as_tensor(values, dtype=config.real(torch))
        self.component = component
        if batch_size is not None:
            raise NotImplementedError("PointSetBC batch_size is outside the oracle's scope")

    def __len__(self):
        return self.points.shape[0]

    def collocation_points(self, X):
        return self.points

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        if isinstance(self.component, numbers.Number):
            return outputs[beg:end, self.component : self.component + 1] - self.values
        return outputs[beg:end, self.component] - self.values


class PointSetOperatorBC:
    def __init__(self, points, values, func):
        self.points = np.array(points, dtype=config.real(np))
        if not isinstance(values, numbers.Number) and values.shape[1] != 1:
            raise RuntimeError("PointSetOperatorBC should output 1D values")
        self.values = torch.as_tensor(values, dtype=config.real(torch))
        self.func = func

    def collocation_points(self, X):
        return self.points

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        return self.func(inputs, outputs, X)[beg:end] - self.values
