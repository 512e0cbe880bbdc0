import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def make_field(values, start="2000-01-01", calendar="standard", units="K", dims=None, **attrs):
    """A Field with a daily time axis on dim 0 (the `*_series` fixtures of the reference's
    tests/conftest.py, without xarray)."""
    from xclim_b200 import Field, TimeAxis
    values = np.asarray(values)
    ta = TimeAxis.daily(start, values.shape[0], calendar)
    if dims is None:
        dims = ("time",) + tuple(f"d{i}" for i in range(values.ndim - 1))
    return Field(values, dims, ta, {}, {"units": units, **attrs})
