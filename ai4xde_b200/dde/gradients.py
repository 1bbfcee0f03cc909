"""dde.grad.jacobian / hessian / clear of the fused path.

Same signatures, return shapes, caching-until-clear() and ValueErrors as DeepXDE's gradients
module.  Two routes:
  * `ys` is the raw output of a fused network evaluation of `xs` -> the requested column is
    read straight from the kernel's jet streams (no autograd pass at all);
  * anything else (output transforms, derived tensors such as `jacobian(mu, x)` at
    /root/reference/AI4XDE/cases/PDECases.py:2578-2583, residuals in gPINN) -> torch.autograd
    with create_graph=True exactly like DeepXDE; the graph bottoms out in jet leaves
    (ai4xde_b200/jet.py), so the MLP itself is never re-differentiated.
"""
import torch

from .. import residual as _residual

_records = {}  # id(y) -> (y, JetOutput): alive until clear()


def register(y, record):
    _records[id(y)] = (y, record)


def _record_for(ys, xs):
    hit = _records.get(id(ys))
    if hit is None or hit[0] is not ys or hit[1].x is not xs:
        return None
    return hit[1]


class Jacobian:
    def __init__(self, ys, xs):
        self.ys, self.xs = ys, xs
        self.dim_y, self.dim_x = ys.shape[1], xs.shape[1]
        self.J = {}
        self.rec = _record_for(ys, xs)

    def __call__(self, i=0, j=None):
        if not 0 <= i < self.dim_y:
            raise ValueError("i={} is not valid.".format(i))
        if j is not None and not 0 <= j < self.dim_x:
            raise ValueError("j={} is not valid.".format(j))
        if self.rec is not None and i not in self.J:
            if j is not None:
                col = self.rec.first(i, j)
                if col is not None:
                    return col
            else:
                row = self.rec.first_all(i)
                if row is not None:
                    self.J[i] = row
        if i not in self.J:
            y = self.ys[:, i : i + 1] if self.dim_y > 1 else self.ys
            self.J[i] = torch.autograd.grad(y, self.xs, grad_outputs=torch.ones_like(y), create_graph=True)[0]
        return self.J[i] if j is None or self.dim_x == 1 else self.J[i][:, j : j + 1]


class Hessian:
    def __init__(self, y, xs, component=None, grad_y=None):
        dim_y = y.shape[1]
        if dim_y > 1:
            if component is None:
                raise ValueError("The component of y is missing.")
            if component >= dim_y:
                raise ValueError(
                    "The component of y={} cannot be larger than the dimension={}.".format(component, dim_y)
                )
        else:
            if component is not None:
                raise ValueError("Do not use component for 1D y.")
            component = 0
        self.component = component
        self.rec = _record_for(y, xs) if grad_y is None else None
        self.y, self.xs, self.grad_y = y, xs, grad_y
        self.H = None

    def __call__(self, i=0, j=0):
        dim_x = self.xs.shape[1]
        if not 0 <= i < dim_x:
            raise ValueError("i={} is not valid.".format(i))
        if not 0 <= j < dim_x:
            raise ValueError("j={} is not valid.".format(j))
        if self.rec is not None:
            col = self.rec.second(self.component, i, j)
            if col is not None:
                return col
        if self.H is None:
            g = self.grad_y
            if g is None:
                g = jacobian(self.y, self.xs, i=self.component, j=None)
            self.H = Jacobian(g, self.xs)
        return self.H(i, j)


_jacobians, _hessians = {}, {}


def jacobian(ys, xs, i=0, j=None):
    if isinstance(ys, _residual.Sym):       # the loss code is being traced into a residual program
        return _residual.jacobian(ys, xs, i=i, j=j)
    key = (id(ys), id(xs))
    J = _jacobians.get(key)
    if J is None or J.ys is not ys or J.xs is not xs:
        J = _jacobians[key] = Jacobian(ys, xs)
    return J(i, j)


def hessian(ys, xs, component=None, i=0, j=0, grad_y=None):
    if isinstance(ys, _residual.Sym):
        return _residual.hessian(ys, xs, component=component, i=i, j=j, grad_y=grad_y)
    key = (id(ys), id(xs), component)
    H = _hessians.get(key)
    if H is None or H.y is not ys or H.xs is not xs:
        H = _hessians[key] = Hessian(ys, xs, component=component, grad_y=grad_y)
    return H(i, j)


def clear():
    _jacobians.clear()
    _hessians.clear()
    _records.clear()
