"""`deepxde.backend` names used by AI4XDE cases, on the torch device."""
import numpy as np
import torch as _torch

from . import config

backend_name = "pytorch"
torch = _torch
tf = jax = paddle = None
lib = _torch


def is_tensor(obj):
    return _torch.is_tensor(obj)


def shape(t):
    return list(t.shape)


def size(t):
    return _torch.numel(t)


def ndim(t):
    return t.dim()


def transpose(t, axes=None):
    return _torch.permute(t, axes if axes is not None else tuple(range(t.dim())[::-1]))


def reshape(t, shape):
    return _torch.reshape(t, shape)


def Variable(initial_value, dtype=None):
    """A trainable scalar of an inverse problem (cases/InverseCases.py:61-63): an ordinary leaf tensor for autograd and
    the optimiser, and recognised by the residual-program tracer (ai4xde_b200/residual.py:VariableTensor)."""
    from ..residual import VariableTensor

    data = _torch.tensor(initial_value, dtype=dtype or config.real(_torch), device=config.device())
    return VariableTensor.make(data)


def as_tensor(data, dtype=None):
    if isinstance(data, _torch.Tensor):
        return data if dtype is None or data.dtype == dtype else data.type(dtype=dtype)
    return _torch.as_tensor(data, dtype=dtype, device=config.device())


def from_numpy(a):
    if isinstance(a, _torch.Tensor):      # DeepXDE's is torch.as_tensor: a tensor passes through, graph intact
        return a                          # (the reference's IDE cases call it on the network output)
    t = _torch.as_tensor(np.asarray(a))
    if t.is_floating_point():
        t = t.to(config.real(_torch))
    return t.to(config.device())


def to_numpy(t):
    return t.detach().cpu().numpy()


def sparse_tensor(indices, values, shape):
    return _torch.sparse_coo_tensor(list(zip(*indices)), values, shape, requires_grad=False,
                                    device=config.device())


def concat(values, axis):
    return _torch.cat(values, axis)


def stack(values, axis):
    return _torch.stack(values, axis)


def expand_dims(t, axis):
    return _torch.unsqueeze(t, axis)


def reverse(t, axis):
    return _torch.flip(t, axis)


def roll(t, shift, axis):
    return _torch.roll(t, shift, axis)


lgamma, sigmoid, sin, cos, exp, square, tanh, pow, minimum = (
    _torch.lgamma, _torch.sigmoid, _torch.sin, _torch.cos, _torch.exp, _torch.square, _torch.tanh,
    _torch.pow, _torch.minimum,
)
abs = _torch.abs
relu, elu, selu, silu = (_torch.nn.functional.relu, _torch.nn.functional.elu,
                         _torch.nn.functional.selu, _torch.nn.functional.silu)


def mean(t, dim, keepdims=False):
    return _torch.mean(t, dim, keepdim=keepdims)


def reduce_mean(t):
    return _torch.mean(t)


def sum(t, dim, keepdims=False):
    return _torch.sum(t, dim, keepdim=keepdims)


def reduce_sum(t):
    return _torch.sum(t)


def prod(t, dim, keepdims=False):
    return _torch.prod(t, dim, keepdim=keepdims)


def norm(t, ord=None, axis=None, keepdims=False):
    return _torch.linalg.norm(t, ord=ord, dim=axis, keepdim=keepdims)


def zeros(shape, dtype):
    return _torch.zeros(shape, dtype=dtype, device=config.device())


def zeros_like(t):
    return _torch.zeros_like(t)


def matmul(x, y):
    return _torch.mm(x, y)


def sparse_dense_matmul(x, y):
    return _torch.sparse.mm(x, y)
