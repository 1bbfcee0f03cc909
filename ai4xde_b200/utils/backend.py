"""setup_seed / empty_cache with the semantics of /root/reference/AI4XDE/utils/backend.py:19-50
for the PyTorch backend (the only backend of the fused path)."""
import os
import random

import numpy as np
import torch


def empty_cache():
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def setup_seed(seed=0):
    np.random.seed(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    os.environ["PYTHONHASHSEED"] = str(seed)
