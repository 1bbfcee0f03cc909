"""Default float handling. Restates deepxde.config (set_default_float / real)
as used by /root/reference/AI4XDE/solver/PDESolver.py:6."""
import numpy as np
import torch


class _Real:
    def __init__(self):
        self.set_float32()

    def set_float32(self):
        self.precision = 32
        self.reals = {np: np.float32, torch: torch.float32}

    def set_float64(self):
        self.precision = 64
        self.reals = {np: np.float64, torch: torch.float64}

    def __call__(self, package):
        return self.reals[package]


real = _Real()


def set_default_float(value):
    if value == "float32":
        real.set_float32()
        torch.set_default_dtype(torch.float32)
    elif value == "float64":
        real.set_float64()
        torch.set_default_dtype(torch.float64)
    else:
        raise ValueError(f"{value} not supported")


def default_float():
    return "float64" if real.precision == 64 else "float32"


def set_random_seed(seed):
    import random

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
