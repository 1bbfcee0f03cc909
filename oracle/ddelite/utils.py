"""deepxde.utils subset: get_num_args (algorithm/gPINN.py:17, R3Sampling.py:31),
to_numpy for predict(), decorators used by data.PDE caching."""
import inspect
from functools import wraps

import numpy as np
import torch

from . import config


def get_num_args(func):
    params = inspect.signature(func).parameters
    return len(params) - ("self" in params)


def to_numpy(tensors):
    if isinstance(tensors, (list, tuple)):
        return [to_numpy(t) for t in tensors]
    if torch.is_tensor(tensors):
        return tensors.detach().cpu().numpy()
    return np.asarray(tensors)


def run_if_all_none(*attr):
    def decorator(func):
        @wraps(func)
        def wrapper(self, *args, **kwargs):
            x = [getattr(self, a) for a in attr]
            if all(i is None for i in x):
                return func(self, *args, **kwargs)
            return x if len(x) > 1 else x[0]

        return wrapper

    return decorator


def run_if_any_none(*attr):
    def decorator(func):
        @wraps(func)
        def wrapper(self, *args, **kwargs):
            x = [getattr(self, a) for a in attr]
            if any(i is None for i in x):
                return func(self, *args, **kwargs)
            return x if len(x) > 1 else x[0]

        return wrapper

    return decorator


def return_tensor(func):
    @wraps(func)
    def wrapper(*args, **kwargs):
        return torch.as_tensor(func(*args, **kwargs), dtype=config.real(torch))

    return wrapper


def list_to_str(nums, precision=2):
    if nums is None:
        return ""
    if not isinstance(nums, (list, tuple, np.ndarray)):
        return "{:.{}e}".format(nums, precision)
    return "[{:s}]".format(", ".join(["{:.{}e}".format(x, precision) for x in nums]))
