"""deepxde.utils subset used by AI4XDE (get_num_args at algorithm/gPINN.py:17 etc.)."""
import inspect

import numpy as np
import torch


def get_num_args(func):
    params = inspect.signature(func).parameters
    return len(params) - ("self" in params)


def to_numpy(tensors):
    if isinstance(tensors, (list, tuple)):
        return [to_numpy(t) for t in tensors]
    if torch.is_tensor(tensors):
        return tensors.detach().cpu().numpy()
    return np.asarray(tensors)


def list_to_str(nums, precision=2):
    if nums is None:
        return ""
    if not isinstance(nums, (list, tuple, np.ndarray)):
        return "{:.{}e}".format(nums, precision)
    return "[{:s}]".format(", ".join(["{:.{}e}".format(x, precision) for x in nums]))
