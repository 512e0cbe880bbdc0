"""Shared test helpers (imported as a top-level module: the tests directory is on sys.path)."""
import numpy as np


def make_field(values, start="2000-01-01", calendar="standard", units="K", dims=None, **attrs):
    """A Field with a daily time axis on dim 0 (the `*_series` fixtures of the reference's
    tests/conftest.py, without xarray)."""
    from xclim_b200 import Field, TimeAxis
    values = np.asarray(values)
    ta = TimeAxis.daily(start, values.shape[0], calendar)
    if dims is None:
        dims = ("time",) + tuple(f"d{i}" for i in range(values.ndim - 1))
    return Field(values, dims, ta, {}, {"units": units, **attrs})
