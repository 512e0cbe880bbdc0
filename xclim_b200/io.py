"""The I/O step either side of the hot path.

The reference's CLI opens chunked files lazily (``xr.open_mfdataset(..., chunks=...)``, cli.py:68-71), runs
the indicators chunk by chunk under dask and writes with ``to_netcdf(compute=False)`` (cli.py:483-497).
NetCDF / zarr readers are not part of this image, so the on-disk format here is the ``.npy`` array (plus a
JSON side-car for the time axis and attrs) opened as a MEMORY MAP: the returned :class:`Field` is host-backed
and lazy, and the slab streamer (:mod:`xclim_b200.streaming`) then does what the dask graph does --
a reader thread gathers lat slab k+2 from the file into a pinned buffer while slab k+1 crosses PCIe and
slab k is in the kernels.  A real deployment plugs its own reader in the same way: anything that exposes a
``(time, lat, lon)`` float32 buffer interface (numpy memmap, ``zarr.Array[...]``, ``netCDF4.Variable[...]``
materialised per slab) can back a Field.
"""
from __future__ import annotations

import json
import os

import numpy as np

from .field import Field
from .timeaxis import TimeAxis


def save_npy(path: str, field, start: str | None = None) -> str:
    """Write ``field`` (values + daily time axis + attrs) as ``path`` (.npy) and ``path + ".json"``."""
    vals = field.numpy() if isinstance(field, Field) else np.asarray(field.values)
    np.save(path, vals)
    meta = {"dims": list(field.dims), "attrs": {k: v for k, v in dict(field.attrs).items() if _plain(v)}}
    ta = getattr(field, "time", None)
    if isinstance(ta, TimeAxis) and len(ta):
        meta["time"] = {"start": ta.date_strings(0)[0], "periods": len(ta), "calendar": ta.calendar}
    elif "time" in getattr(field, "coords", {}):
        meta["time_labels"] = [str(v) for v in np.asarray(field.coords["time"]).tolist()]
    with open(path + ".json", "w") as f:
        json.dump(meta, f)
    return path


def open_npy(path: str, mmap: bool = True) -> Field:
    """Open an array written by :func:`save_npy` as a lazy, memory-mapped, host-backed Field."""
    vals = np.load(path, mmap_mode="r" if mmap else None)
    meta = {}
    if os.path.exists(path + ".json"):
        with open(path + ".json") as f:
            meta = json.load(f)
    dims = tuple(meta.get("dims", ("time",) + tuple(f"d{i}" for i in range(vals.ndim - 1))))
    ta = None
    coords = {}
    if "time" in meta:
        t = meta["time"]
        ta = TimeAxis.daily(t["start"], int(t["periods"]), t["calendar"])
    elif "time_labels" in meta:
        coords["time"] = np.array(meta["time_labels"])
    return Field(vals, dims, ta, coords, dict(meta.get("attrs", {})))


def _plain(v) -> bool:
    return isinstance(v, (str, int, float, bool, list)) or v is None
