import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def oracle_f64():
    """ddelite in float64 for the duration of a test (the reference's default float,
    /root/reference/AI4XDE/solver/PDESolver.py:6), torch default dtype restored afterwards."""
    import torch
    from oracle import ddelite

    ddelite.config.set_default_float("float64")
    yield ddelite
    ddelite.config.set_default_float("float32")
    torch.set_default_dtype(torch.float32)
