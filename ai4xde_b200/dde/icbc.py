"""Initial / boundary conditions of the fused path.

`collocation_points` / `filter` run on the host once per point set; `error` runs on the device
every step on slices of the batch.  Target values (`func(X[beg:end])`) are evaluated on the host
once per (point set, slice) and kept resident on the device."""
import numbers

import numpy as np
import torch

from . import config, gradients as grad


class _DeviceValues:
    """func(X[beg:end]) -> device tensor, cached per (X, beg, end).  The entry holds X itself and compares with
    `is`: an id() alone could be recycled by a later array of the same shape and serve stale targets."""

    def __init__(self, func):
        self.func = func
        self._X = None
        self._key = None
        self._val = None

    def __call__(self, X, beg, end):
        key = (X.shape, beg, end)
        if X is not self._X or key != self._key:
            v = self.func(X[beg:end])
            self._val = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v,
                                        dtype=config.real(torch), device=config.device())
            self._X, self._key = X, key
        return self._val


def _rowwise(pred):
    """DeepXDE calls user predicates one row at a time: pred(x_row, flag_row)."""
    return lambda X, flags: np.array([pred(X[i], flags[i]) for i in range(len(X))], dtype=bool).reshape(-1)


class BC:
    def __init__(self, geom, on_boundary, component):
        self.geom = geom
        self.on_boundary = _rowwise(on_boundary)
        self.component = component
        self._normal = _DeviceValues(self.geom.boundary_normal)

    def filter(self, X):
        return X[self.on_boundary(X, self.geom.on_boundary(X))]

    def collocation_points(self, X):
        return self.filter(X)

    def normal_derivative(self, X, inputs, outputs, beg, end):
        dydx = grad.jacobian(outputs, inputs, i=self.component, j=None)[beg:end]
        return torch.sum(dydx * self._normal(X, beg, end), 1, keepdim=True)

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        raise NotImplementedError


class DirichletBC(BC):
    def __init__(self, geom, func, on_boundary, component=0):
        super().__init__(geom, on_boundary, component)
        self.func = _DeviceValues(func)

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        values = self.func(X, beg, end)
        if values.dim() == 2 and values.shape[1] != 1:
            raise RuntimeError("DirichletBC function should return an array of shape N by 1 for each "
                               "component. Use argument 'component' for different output components.")
        return outputs[beg:end, self.component : self.component + 1] - values


class NeumannBC(BC):
    def __init__(self, geom, func, on_boundary, component=0):
        super().__init__(geom, on_boundary, component)
        self.func = _DeviceValues(func)

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        return self.normal_derivative(X, inputs, outputs, beg, end) - self.func(X, beg, end)


class RobinBC(BC):
    def __init__(self, geom, func, on_boundary, component=0):
        super().__init__(geom, on_boundary, component)
        self.func = func

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        return self.normal_derivative(X, inputs, outputs, beg, end) - self.func(X[beg:end], outputs[beg:end])


class PeriodicBC(BC):
    def __init__(self, geom, component_x, on_boundary, derivative_order=0, component=0):
        super().__init__(geom, on_boundary, component)
        self.component_x = component_x
        self.derivative_order = derivative_order
        if derivative_order > 1:
            raise NotImplementedError("PeriodicBC only supports derivative_order 0 or 1 now.")

    def collocation_points(self, X):
        X1 = self.filter(X)
        return np.vstack((X1, self.geom.periodic_point(X1, self.component_x)))

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        mid = beg + (end - beg) // 2
        if self.derivative_order == 0:
            left = outputs[beg:mid, self.component : self.component + 1]
            right = outputs[mid:end, self.component : self.component + 1]
        else:
            dydx = grad.jacobian(outputs, inputs, i=self.component, j=self.component_x)
            left, right = dydx[beg:mid], dydx[mid:end]
        return left - right


class OperatorBC(BC):
    def __init__(self, geom, func, on_boundary):
        super().__init__(geom, on_boundary, 0)
        self.func = func

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        return self.func(inputs, outputs, X)[beg:end]


class IC:
    def __init__(self, geom, func, on_initial, component=0):
        self.geom = geom
        self.func = _DeviceValues(func)
        self.on_initial = _rowwise(on_initial)
        self.component = component

    def filter(self, X):
        return X[self.on_initial(X, self.geom.on_initial(X))]

    def collocation_points(self, X):
        return self.filter(X)

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        values = self.func(X, beg, end)
        if values.dim() == 2 and values.shape[1] != 1:
            raise RuntimeError("IC function should return an array of shape N by 1 for each component.")
        return outputs[beg:end, self.component : self.component + 1] - values


class PointSetBC:
    def __init__(self, points, values, component=0, batch_size=None, shuffle=True):
        self.points = np.array(points, dtype=config.real(np))
        self._values_host = np.asarray(values)
        self._values = None
        self.component = component
        if batch_size is not None:
            raise NotImplementedError("PointSetBC mini-batching is outside the fused path")

    def __len__(self):
        return self.points.shape[0]

    @property
    def values(self):
        if self._values is None:
            self._values = torch.as_tensor(self._values_host, dtype=config.real(torch), device=config.device())
        return self._values

    def collocation_points(self, X):
        return self.points

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        if isinstance(self.component, numbers.Number):
            return outputs[beg:end, self.component : self.component + 1] - self.values
        return outputs[beg:end, self.component] - self.values


class PointSetOperatorBC:
    def __init__(self, points, values, func):
        self.points = np.array(points, dtype=config.real(np))
        if not isinstance(values, numbers.Number) and np.asarray(values).shape[1] != 1:
            raise RuntimeError("PointSetOperatorBC should output 1D values")
        self._values_host = np.asarray(values)
        self._values = None
        self.func = func

    @property
    def values(self):
        if self._values is None:
            self._values = torch.as_tensor(self._values_host, dtype=config.real(torch), device=config.device())
        return self._values

    def collocation_points(self, X):
        return self.points

    def error(self, X, inputs, outputs, beg, end, aux_var=None):
        return self.func(inputs, outputs, X)[beg:end] - self.values
