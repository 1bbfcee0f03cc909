"""Seeding and cache helpers for the PyTorch-only fused path.

`setup_seed(seed)` seeds exactly what the reference's helper seeds for its PyTorch backend
(/root/reference/AI4XDE/utils/backend.py:19-50): NumPy's global legacy RNG (which the geometry samplers
and the RAD / RAR_D draws consume), Python's `random`, torch CPU and every CUDA device, and turns on
deterministic cuDNN; the parity tests rely on that order of streams."""
import os
import random

import numpy as np
import torch


def setup_seed(seed=0):
    for seeder in (np.random.seed, random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    os.environ["PYTHONHASHSEED"] = str(seed)


def empty_cache():
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
