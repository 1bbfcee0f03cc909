"""Backend shim: the `deepxde.backend` names AI4XDE touches."""
import numpy as np
import torch as _torch

from . import config

backend_name = "pytorch"
torch = _torch
tf = None
jax = None
paddle = None
lib = _torch


def data_type_dict():
    return {"float32": _torch.float32, "float64": _torch.float64}


def is_tensor(obj):
    return _torch.is_tensor(obj)


def shape(t):
    return list(t.shape)


def size(t):
    return _torch.numel(t)


def ndim(t):
    return t.dim()


def transpose(t, axes=None):
    if axes is None:
        axes = tuple(range(t.dim())[::-1])
    return _torch.permute(t, axes)


def reshape(t, shape):
    return _torch.reshape(t, shape)


def Variable(initial_value, dtype=None):
    return _torch.tensor(initial_value, dtype=dtype or config.real(_torch), requires_grad=True)


def as_tensor(data, dtype=None):
    if isinstance(data, _torch.Tensor):
        if dtype is None or data.dtype == dtype:
            return data
        return data.type(dtype=dtype)
    return _torch.as_tensor(data, dtype=dtype)


def sparse_tensor(indices, values, shape):
    return _torch.sparse_coo_tensor(list(zip(*indices)), values, shape, requires_grad=False)


def from_numpy(a):
    return _torch.as_tensor(a)


def to_numpy(t):
    return t.detach().cpu().numpy()


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


def lgamma(x):
    return _torch.lgamma(x)


def elu(x):
    return _torch.nn.functional.elu(x)


def relu(x):
    return _torch.nn.functional.relu(x)


def selu(x):
    return _torch.nn.functional.selu(x)


def sigmoid(x):
    return _torch.sigmoid(x)


def silu(x):
    return _torch.nn.functional.silu(x)


def sin(x):
    return _torch.sin(x)


def cos(x):
    return _torch.cos(x)


def exp(x):
    return _torch.exp(x)


def square(x):
    return _torch.square(x)


def abs(x):
    return _torch.abs(x)


def minimum(x, y):
    return _torch.minimum(x, y)


def tanh(x):
    return _torch.tanh(x)


def pow(x, y):
    return _torch.pow(x, y)


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
    return _torch.zeros(shape, dtype=dtype)


def zeros_like(t):
    return _torch.zeros_like(t)


def matmul(x, y):
    return _torch.mm(x, y)


def sparse_dense_matmul(x, y):
    return _torch.sparse.mm(x, y)
