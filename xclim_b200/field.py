"""Input/output containers: unwrap a DataArray to a contiguous (time, cell) float32 device buffer.

The reference's operator interface is ``f(*DataArrays, **params) -> DataArray`` (core/indicator.py
:884-886 calls ``self.compute(**args)``).  This module is the L1 replacement described in SURVEY.md
section 1: *unwrap DataArray -> contiguous (time, lat, lon) float32 device buffer -> C-ABI call ->
wrap result*.  When xarray is not installed (the authoring container) the same functions accept
and return :class:`Field`, a minimal labelled array with the few attributes the wrappers need.
"""
from __future__ import annotations

from dataclasses import dataclass, field as _dc_field

import numpy as np

from .timeaxis import TimeAxis

try:  # xarray is optional: present wherever the real xclim is installed
    import xarray as xr  # type: ignore
except Exception:  # pragma: no cover - absent in the authoring container
    xr = None


@dataclass
class Field:
    """Minimal stand-in for ``xarray.DataArray``: values + dims + a daily time axis + attrs.

    ``values`` is a numpy array or a CUDA ``torch.Tensor`` (device-resident data stay in HBM between
    calls).  ``time`` is the :class:`TimeAxis` of dimension ``"time"`` (or ``None``).
    """

    values: object
    dims: tuple
    time: TimeAxis | None = None
    coords: dict = _dc_field(default_factory=dict)
    attrs: dict = _dc_field(default_factory=dict)
    name: str | None = None

    @property
    def shape(self):
        return tuple(self.values.shape)

    @property
    def dtype(self):
        return self.values.dtype

    def numpy(self) -> np.ndarray:
        v = self.values
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return np.asarray(v)

    def assign_attrs(self, **kw) -> "Field":
        return Field(self.values, self.dims, self.time, dict(self.coords), {**self.attrs, **kw}, self.name)

    def isel_time(self, sl: slice) -> "Field":
        ax = self.dims.index("time")
        idx = [slice(None)] * len(self.dims)
        idx[ax] = sl
        return Field(self.values[tuple(idx)], self.dims, self.time.isel(sl) if self.time else None,
                     dict(self.coords), dict(self.attrs), self.name)


def is_xarray(obj) -> bool:
    return xr is not None and isinstance(obj, xr.DataArray)


def time_axis_of(obj) -> TimeAxis:
    if isinstance(obj, Field):
        if obj.time is None:
            raise ValueError("input has no time axis")
        return obj.time
    if is_xarray(obj):
        cache = obj.attrs.get("_xclim_b200_timeaxis") if False else None  # attrs are user data: do not pollute
        return cache or TimeAxis.from_xarray(obj["time"])
    raise TypeError(f"expected an xarray.DataArray or xclim_b200.Field, got {type(obj).__name__}")


def attrs_of(obj) -> dict:
    return dict(obj.attrs)


def dims_of(obj) -> tuple:
    return tuple(obj.dims)


def raw_values(obj):
    """The underlying array (numpy, or torch tensor for device-resident Fields)."""
    if isinstance(obj, Field):
        return obj.values
    return obj.values  # xarray: loads dask-backed data


def wrap_like(template, values, dims, *, time=None, coords_extra=None, attrs=None, name=None):
    """Build the output container of the same family as ``template``.

    ``dims`` are the output dims; non-time coords of ``template`` along kept dims are carried over;
    ``time`` (a TimeAxis-derived list of ISO labels, a TimeAxis, or an xarray coordinate) labels a
    "time" dim if present.
    """
    attrs = dict(attrs or {})
    if is_xarray(template):
        coords = {}
        for d in dims:
            if d == "time":
                continue
            if d in template.coords:
                coords[d] = template.coords[d]
        if coords_extra:
            coords.update(coords_extra)
        if "time" in dims and time is not None:
            coords["time"] = time
        if hasattr(values, "detach"):
            values = values.detach().cpu().numpy()
        return xr.DataArray(values, dims=dims, coords=coords, attrs=attrs, name=name)
    coords = {d: template.coords[d] for d in dims if d != "time" and d in getattr(template, "coords", {})}
    if coords_extra:
        coords.update(coords_extra)
    t = time if isinstance(time, TimeAxis) else None
    if time is not None and not isinstance(time, TimeAxis):
        coords["time"] = time
    return Field(values, tuple(dims), t, coords, attrs, name)
