"""TEST INFRASTRUCTURE: the few dozen lines of the xarray.DataArray interface that the xclim_b200 host layer
touches (xarray itself is absent from this image).  ``install(monkeypatch)`` plugs this module in as
``xclim_b200.field.xr`` so that the xarray branches of field.py / generic.py / calendar.py / streaming.py
(``is_xarray``, ``TimeAxis.from_xarray``, ``wrap_like``, ``_period_time``, ``_assemble``) are executed by
the test-suite.  tests/test_xarray_real.py runs the same scenarios against the real package when it is
importable.  Nothing here is a product path."""
import numpy as np


class _Values:
    def __init__(self, v):
        self.values = np.asarray(v)


class _DT:
    """``time.dt``: integer date fields + calendar of a daily axis (built from a TimeAxis)."""

    def __init__(self, ta):
        self.year, self.month, self.day, self.dayofyear = (_Values(ta.year), _Values(ta.month), _Values(ta.day),
                                                             _Values(ta.doy))
        self.calendar = ta.calendar


class DataArray:
    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        self.values = data if hasattr(data, "shape") else np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(self.values.ndim))
        self.coords = {}
        for k, v in (coords or {}).items():
            self.coords[k] = v if isinstance(v, (DataArray, TimeCoord)) else DataArray(np.asarray(v), dims=(k,))
        self.attrs = dict(attrs or {})
        self.name = name

    # ---- the interface used by the host layer
    @property
    def data(self):
        return self.values

    @property
    def shape(self):
        return tuple(self.values.shape)

    @property
    def dtype(self):
        return self.values.dtype

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[key]
        raise TypeError("mini_xarray: positional indexing is not part of the stand-in")

    def __len__(self):
        return self.values.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def assign_attrs(self, **kw):
        return DataArray(self.values, self.dims, self.coords, {**self.attrs, **kw}, self.name)

    def copy(self, data=None):
        return DataArray(self.values.copy() if data is None else data, self.dims, self.coords, self.attrs, self.name)

    def astype(self, dtype):
        return DataArray(self.values.astype(dtype), self.dims, self.coords, self.attrs, self.name)

    def sel(self, **kw):
        (dim, val), = kw.items()
        ax = self.dims.index(dim)
        i = int(np.nonzero(np.asarray(self.coords[dim].values) == val)[0][0])
        coords = {k: v for k, v in self.coords.items() if k != dim}
        coords[dim] = DataArray(np.asarray(self.coords[dim].values)[i], dims=())
        return DataArray(np.take(self.values, i, axis=ax), tuple(d for d in self.dims if d != dim), coords, self.attrs,
                         self.name)

    def squeeze(self, dim):
        return self.sel(**{dim: np.asarray(self.coords[dim].values).reshape(-1)[0]})


class TimeCoord:
    """The ``time`` coordinate: ``.dt`` fields, slicing (``TimeAxis.isel``) and a failing ``resample`` so
    that ``_period_time`` takes its label fallback, as it does for exotic calendars."""

    def __init__(self, ta):
        self._ta = ta
        self.dt = _DT(ta)
        self.dims = ("time",)

    @property
    def values(self):
        return np.array(self._ta.date_strings(slice(None)) if False else [f"{y:04d}-{m:02d}-{d:02d}" for y, m, d in
                                                                        zip(self._ta.year, self._ta.month, self._ta.day)])

    def __len__(self):
        return len(self._ta)

    def __getitem__(self, sl):
        return TimeCoord(self._ta.isel(sl))

    def resample(self, **kw):
        raise NotImplementedError("mini_xarray has no resample: the host layer falls back to period labels")


def daily(values, start, calendar="noleap", units="K", dims=("time", "lat", "lon"), **attrs):
    """A DataArray with a daily ``time`` coordinate and plain integer spatial coordinates."""
    from xclim_b200 import TimeAxis
    values = np.asarray(values)
    ta = TimeAxis.daily(start, values.shape[0], calendar)
    coords = {"time": TimeCoord(ta)}
    for ax, d in enumerate(dims[1:], start=1):
        coords[d] = DataArray(np.arange(values.shape[ax]) * 0.25, dims=(d,))
    return DataArray(values, dims=dims, coords=coords, attrs={"units": units, **attrs})


def install(monkeypatch):
    import sys

    import xclim_b200.field as field
    this = sys.modules[__name__]
    monkeypatch.setattr(field, "xr", this)
    monkeypatch.setitem(sys.modules, "xarray", this)       # streaming._assemble does `import xarray as xr`
