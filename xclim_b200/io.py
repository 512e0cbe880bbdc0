"""The I/O step either side of the hot path.

The reference's CLI opens chunked files lazily (``xr.open_mfdataset(..., chunks=...)``, cli.py:68-71), runs
the indicators chunk by chunk under dask and writes with ``to_netcdf(compute=False)`` (cli.py:483-497).
Here a file-backed variable is opened as a host-backed, LAZY :class:`Field`; the slab streamer
(:mod:`xclim_b200.streaming`) then does what the dask graph does -- a reader thread gathers lat slab k+2
from the file into a page-locked buffer (decoding it on the way) while slab k+1 crosses PCIe and slab k is
in the kernels.  Slab edges are aligned with the file's chunk edges along the leading spatial dimension, so
no chunk is read or inflated twice.

Formats (none of netCDF4 / h5py / zarr / xarray is in this image, so the readers are written here):

* ``.npy`` + JSON side-car, memory-mapped (:func:`open_npy` / :func:`save_npy`);
* NetCDF-3 classic and 64-bit-offset files through ``scipy.io.netcdf_file`` (memory-mapped; big-endian on
  disk, record variables interleaved per time step): :func:`open_netcdf3` / :func:`save_netcdf3`;
* zarr v2 directory stores with ``C``-order chunks, compressor ``null`` / ``zlib`` / ``gzip`` / ``bz2`` /
  ``lzma`` and the xarray ``_ARRAY_DIMENSIONS`` convention: :func:`open_zarr` / :func:`save_zarr`.

CF decoding follows what ``xr.open_dataset`` does by default: ``_FillValue`` / ``missing_value`` -> NaN,
``scale_factor`` / ``add_offset`` unpacking to float32, ``"days since ..."`` + ``calendar`` -> the daily time
axis.  NetCDF-4 / HDF5 files are outside what can be read here (no HDF5 library); a deployment that has
netCDF4 or zarr plugs them in through the same :class:`LazyGrid` protocol (``shape``, ``lead_chunk``,
``read_rows``).
"""
from __future__ import annotations

import json
import os
import re

import numpy as np

from .field import Field
from .timeaxis import TimeAxis

# ------------------------------------------------------------------------------------------------ .npy


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
    meta["coords"] = {d: np.asarray(c).tolist() for d, c in getattr(field, "coords", {}).items()
                      if d in field.dims and d != "time" and np.asarray(c).dtype.kind in "fiu"}
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
    for d, c in meta.get("coords", {}).items():
        coords[d] = np.asarray(c)
    return Field(vals, dims, ta, coords, dict(meta.get("attrs", {})))


def _plain(v) -> bool:
    return isinstance(v, (str, int, float, bool, list)) or v is None


# ------------------------------------------------------------------------------------------------ parallel host copies
#: host copies / decodes of at least this many bytes are split over the copy threads
PARALLEL_MIN_BYTES = 32 << 20
_copy_pool = None


def copy_threads() -> int:
    return max(1, min(16, (os.cpu_count() or 1)))


def parallel_rows(fn, n_rows: int, nbytes: int) -> None:
    """Run ``fn(a, b)`` over contiguous ranges ``[a, b)`` that cover ``range(n_rows)``, on the copy threads when
    the work is large.  One thread moves a strided (time, lat-slab, lon) box at 4-6 GB/s (big-endian data:
    byte-swapped on the way; ``tools/time_staging_copy.py``), an order of magnitude under what PCIe takes; numpy
    releases the GIL inside the copy, so time ranges of the box can go in parallel.  (The authoring container
    gives its threads about one core in total, so the gain is not measurable there; on the GPU box the e2e
    figures of bench.py use page-locked inputs and bypass this path.)"""
    global _copy_pool
    n = copy_threads()
    if n == 1 or nbytes < PARALLEL_MIN_BYTES or n_rows < 2 * n:
        fn(0, n_rows)
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(n, thread_name_prefix="xclim_b200-copy")
    edges = np.linspace(0, n_rows, n + 1).astype(np.int64)
    futs = [_copy_pool.submit(fn, int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    for f in futs:
        f.result()


# ------------------------------------------------------------------------------------------------ lazy sources


class LazyGrid:
    """A file-backed ``(time, lead, ...)`` variable decoded to float32 on demand.

    The streamer needs three things from a source: its ``shape``, the chunk length along the leading
    spatial dimension (``lead_chunk``; slab edges are multiples of it) and ``read_rows(r0, r1, out)``.
    Everything else (``__array__``, ``[...]``) materialises through ``read_rows`` so that small inputs take
    the direct path of the index functions unchanged.
    """

    dtype = np.dtype(np.float32)
    lead_chunk = 1

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape, dtype=np.int64)))
    nbytes = property(lambda self: self.size * 4)

    def read_rows(self, r0: int, r1: int, out: np.ndarray | None = None) -> np.ndarray:
        """``self[:, r0:r1]`` as a C-contiguous native float32 array (written into ``out`` when given)."""
        raise NotImplementedError

    def _out(self, r0, r1, out):
        shp = (self.shape[0], r1 - r0) + self.shape[2:]
        if out is None:
            return np.empty(shp, np.float32)
        if out.shape != shp or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError(f"read_rows: the output buffer must be C-contiguous float32 {shp}")
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.read_rows(0, self.shape[1])
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if len(idx) >= 2 and isinstance(idx[1], slice) and idx[1].step in (None, 1):
            r0, r1, _ = idx[1].indices(self.shape[1])
            rest = (idx[0], slice(None)) + tuple(idx[2:])
            return self.read_rows(r0, max(r0, r1))[rest]
        return np.asarray(self)[idx]

    def __len__(self):
        return self.shape[0]


def _unpack_dtype(raw_dtype, scale, offset):
    """The float type xarray's CF decoding computes ``raw * scale_factor + add_offset`` in (the type of the
    two attributes when they agree, float64 for 4-byte integers or when only an offset is given)."""
    st = np.asarray(scale).dtype if scale is not None else None
    ot = np.asarray(offset).dtype if offset is not None else None
    if st is not None and ot is not None and st == ot and st in (np.float32, np.float64):
        if raw_dtype.kind in "iu" and raw_dtype.itemsize == 4:
            return np.dtype(np.float64)
        return st
    if ot is not None:
        return np.dtype(np.float64)
    return st if st in (np.float32, np.float64) else np.dtype(np.float64)


def _cf_decoder(attrs: dict, raw_dtype: np.dtype):
    """``decode(raw, out)``: fill values -> NaN, then ``raw * scale_factor + add_offset`` (in the float type
    xarray would use), delivered as float32."""
    fills = []
    for key in ("_FillValue", "missing_value"):
        if key in attrs:
            fills.extend(np.atleast_1d(np.asarray(attrs[key])).astype(np.float64).tolist())
    fills = [f for f in fills if not np.isnan(f)]
    scale = np.asarray(attrs["scale_factor"]).reshape(-1)[0] if "scale_factor" in attrs else None
    offset = np.asarray(attrs["add_offset"]).reshape(-1)[0] if "add_offset" in attrs else None
    plain = not fills and scale is None and offset is None
    work = _unpack_dtype(np.dtype(raw_dtype), scale, offset) if (scale is not None or offset is not None) else None

    def decode(raw: np.ndarray, out: np.ndarray) -> None:
        if plain:
            np.copyto(out, raw, casting="unsafe")        # byte order and integer -> float conversion
            return
        mask = None
        for f in fills:
            m = raw == np.asarray(f).astype(raw.dtype)
            mask = m if mask is None else (mask | m)
        if work is None:
            np.copyto(out, raw, casting="unsafe")
        else:
            v = raw.astype(work)
            if scale is not None:
                v *= work.type(scale)
            if offset is not None:
                v += work.type(offset)
            np.copyto(out, v, casting="unsafe")
        if mask is not None:
            out[mask] = np.nan

    return decode


_CF_DROP = ("_FillValue", "missing_value", "scale_factor", "add_offset", "_ARRAY_DIMENSIONS")


def _clean_attrs(attrs: dict) -> dict:
    out = {}
    for k, v in attrs.items():
        if k in _CF_DROP:
            continue
        if isinstance(v, bytes):
            v = v.decode("utf-8", "replace")
        elif isinstance(v, np.ndarray):
            v = v.tolist() if v.size != 1 else v.reshape(-1)[0].item()
        elif isinstance(v, np.generic):
            v = v.item()
        out[k] = v
    return out


_SINCE = re.compile(r"\s*(\w+)\s+since\s+(-?\d{1,4})-(\d{1,2})-(\d{1,2})(?:[T\s]+(\d{1,2}):(\d{1,2})(?::(\d{1,2}(?:\.\d*)?))?)?")


def decode_cf_time(values, units: str, calendar: str = "standard") -> TimeAxis:
    """The daily :class:`TimeAxis` of a CF time coordinate (``"days since YYYY-MM-DD"``, any CF calendar).

    The hot path works on gap-free daily series (as the reference's indicators require for their
    ``freq`` checks, core/indicator.py ``datachecks``), so anything else is an error here.
    """
    m = _SINCE.match(units or "")
    if not m:
        raise ValueError(f"cannot decode time units {units!r}")
    unit = m.group(1).lower().rstrip("s")
    per_day = {"day": 1.0, "d": 1.0, "hour": 24.0, "hr": 24.0, "h": 24.0, "minute": 1440.0, "min": 1440.0,
               "second": 86400.0, "sec": 86400.0, "s": 86400.0}
    if unit not in per_day:
        raise ValueError(f"cannot decode time units {units!r}")
    v = np.asarray(values, dtype=np.float64) / per_day[unit]
    hh, mi, ss = (float(g) if g else 0.0 for g in m.group(5, 6, 7))
    v = v + (hh + mi / 60.0 + ss / 3600.0) / 24.0
    if v.size == 0:
        raise ValueError("empty time coordinate")
    first = int(np.floor(v[0]))
    if first < 0:
        raise ValueError("time values before the reference date are not supported")
    if v.size > 1 and not np.allclose(np.diff(v), 1.0, rtol=0, atol=1e-6):
        raise ValueError("the time coordinate is not a gap-free daily axis")
    ref = f"{int(m.group(2)):04d}-{int(m.group(3)):02d}-{int(m.group(4)):02d}"
    cal = (calendar or "standard").lower()
    full = TimeAxis.daily(ref, first + v.size, cal)
    return full.isel(slice(first, first + v.size))


def encode_cf_time(ta: TimeAxis):
    """``(values, units)`` of a daily axis: days since its first date."""
    return np.arange(len(ta), dtype=np.float64), f"days since {ta.date_strings(0)[0]} 00:00:00"


def _field_of(source, dims, time_vals, time_attrs, coords, attrs, name):
    ta = None
    coords = dict(coords)
    if "time" in dims and time_vals is not None:
        units = time_attrs.get("units")
        try:
            ta = decode_cf_time(time_vals, units, time_attrs.get("calendar", "standard"))
        except ValueError:
            coords["time"] = np.asarray(time_vals)       # not daily: labels only (no freq-based index will run)
    return Field(source, tuple(dims), ta, coords, attrs, name)


# ------------------------------------------------------------------------------------------------ NetCDF-3


class _Nc3Var(LazyGrid):
    """One variable of a memory-mapped NetCDF-3 file (``scipy.io.netcdf_file``; data big-endian on disk)."""

    def __init__(self, ncfile, var):
        super().__init__(var.shape)
        self._data = var.data                # a view into the file's memory map (which it keeps alive)
        # scipy's close() warns when views of the mapping outlive the file object; the views own the mapping from
        # here on, and the file object may go quietly
        if getattr(ncfile, "_mm_buf", None) is not None:
            ncfile._mm_buf = None
        self._decode = _cf_decoder(dict(var._attributes), self._data.dtype)

    def read_rows(self, r0, r1, out=None):
        out = self._out(r0, r1, out)
        parallel_rows(lambda a, b: self._decode(self._data[a:b, r0:r1], out[a:b]), self.shape[0], out.nbytes)
        return out



def _str_attr(v):
    return v.decode("utf-8", "replace") if isinstance(v, bytes) else v


def open_netcdf3(path: str, variable: str | None = None, mmap: bool = True) -> Field:
    """Open one variable of a NetCDF-3 (classic / 64-bit offset) file as a lazy host-backed Field.

    ``variable`` defaults to the only variable with a ``time`` dimension that is not a coordinate.
    Mirrors ``xr.open_dataset(path)[variable]`` (cli.py:68-71) with the default CF decoding.
    """
    from scipy.io import netcdf_file

    nc = netcdf_file(path, "r", mmap=mmap, maskandscale=False)
    names = [n for n, v in nc.variables.items() if n not in nc.dimensions and v.dimensions
             and not n.endswith("_bnds") and not n.endswith("_bounds")]
    if variable is None:
        cand = [n for n in names if "time" in nc.variables[n].dimensions] or names
        if len(cand) != 1:
            raise ValueError(f"{path}: give `variable`, one of {sorted(names)}")
        variable = cand[0]
    if variable not in nc.variables:
        raise KeyError(f"{path} has no variable {variable!r}")
    var = nc.variables[variable]
    dims = tuple(var.dimensions)
    if "time" in dims and dims[0] != "time":
        raise ValueError(f"{variable}: `time` must be the first dimension, found {dims}")
    coords, tvals, tattrs = {}, None, {}
    for d in dims:
        if d in nc.variables:
            cv = nc.variables[d]
            if d == "time":
                tvals = np.array(cv.data, dtype=np.float64)
                tattrs = {k: _str_attr(v) for k, v in cv._attributes.items()}
            else:
                coords[d] = np.array(cv.data).astype(cv.data.dtype.newbyteorder("="))
    return _field_of(_Nc3Var(nc, var), dims, tvals, tattrs, coords, _clean_attrs(dict(var._attributes)), variable)


def save_netcdf3(path: str, field, name: str | None = None, calendar: str | None = None) -> str:
    """Write a Field / DataArray as a NetCDF-3 64-bit-offset file (``time`` = the record dimension) --
    the ``to_netcdf`` leg of cli.py:483-497 for the results of the hot path."""
    from scipy.io import netcdf_file

    vals = field.numpy() if isinstance(field, Field) else np.asarray(field.values)
    dims = tuple(field.dims)
    name = name or getattr(field, "name", None) or "data"
    with netcdf_file(path, "w", version=2) as nc:
        for d, n in zip(dims, vals.shape):
            nc.createDimension(d, None if d == "time" else int(n))
        tv = _time_values(field, calendar) if "time" in dims else None
        if tv is not None:
            t = nc.createVariable("time", "f8", ("time",))
            t[:] = tv[0]
            t.units, t.calendar = tv[1], tv[2]
        for d in dims:
            c = getattr(field, "coords", {}).get(d) if d != "time" else None
            if c is not None and np.asarray(c).dtype.kind in "fiu":
                c = np.asarray(c)
                cv = nc.createVariable(d, c.dtype.newbyteorder("=").char, (d,))
                cv[:] = c
        v = nc.createVariable(name, vals.dtype.newbyteorder("=").char if vals.dtype.kind != "b" else "b", dims)
        v[:] = vals
        for k, a in dict(field.attrs).items():
            if isinstance(a, (str, int, float)) and k not in _CF_DROP:
                setattr(v, k, a)
    return path


def _time_values(field, calendar=None):
    """``(values, units, calendar)`` of the time dimension of a result, or None.  Period results carry ISO
    date labels only: ``calendar`` (default "standard") says in which calendar they are to be counted."""
    ta = getattr(field, "time", None)
    if isinstance(ta, TimeAxis) and len(ta):
        vals, units = encode_cf_time(ta)
        return vals, units, ta.calendar
    labels = getattr(field, "coords", {}).get("time") if isinstance(field, Field) else None
    if labels is None:
        return None
    labels = [str(s)[:10] for s in np.asarray(labels).tolist()]
    if not labels:
        return None
    cal = (calendar or "standard").lower()
    if cal not in _known_calendars():
        raise ValueError(f"Unknown calendar {calendar!r}")
    y0, m0, d0 = (int(x) for x in labels[0].split("-"))
    ref = f"{y0:04d}-{m0:02d}-{d0:02d}"
    y1 = int(labels[-1].split("-")[0])
    span = TimeAxis.daily(ref, (y1 - y0 + 1) * 366 + 1, cal)
    key = {(int(y), int(m), int(d)): i for i, (y, m, d) in enumerate(zip(span.year, span.month, span.day))}
    vals = np.array([key[tuple(int(x) for x in s.split("-"))] for s in labels], dtype=np.float64)
    return vals, f"days since {ref} 00:00:00", span.calendar


def _known_calendars():
    from .timeaxis import MAX_DOY
    return MAX_DOY


# ------------------------------------------------------------------------------------------------ zarr v2


def _decompressor(spec):
    if spec is None:
        return lambda b: b
    cid = spec.get("id")
    if cid in ("zlib", "gzip"):
        import zlib
        wbits = 15 if cid == "zlib" else 31
        return lambda b: zlib.decompress(b, wbits)
    if cid == "bz2":
        import bz2
        return bz2.decompress
    if cid == "lzma":
        import lzma
        return lzma.decompress
    raise NotImplementedError(f"zarr compressor {cid!r} needs a codec library that is not in this image "
                              "(supported: null, zlib, gzip, bz2, lzma)")


def _compressor(spec):
    if spec is None:
        return lambda b: b
    cid, level = spec.get("id"), int(spec.get("level", 1))
    if cid == "zlib":
        import zlib
        return lambda b: zlib.compress(b, level)
    if cid == "gzip":
        import gzip
        return lambda b: gzip.compress(b, compresslevel=level, mtime=0)
    if cid == "bz2":
        import bz2
        return lambda b: bz2.compress(b, level)
    if cid == "lzma":
        import lzma
        return lzma.compress
    raise NotImplementedError(f"zarr compressor {cid!r} is not supported")


class _ZarrArray(LazyGrid):
    """One array of a zarr v2 directory store.  ``read_rows`` touches each chunk that intersects the row
    range exactly once; chunks are read and inflated by a small thread pool (zlib releases the GIL)."""

    def __init__(self, path, meta, attrs, workers=None):
        if meta.get("zarr_format") != 2:
            raise NotImplementedError("only zarr format 2 stores are supported")
        if meta.get("order", "C") != "C":
            raise NotImplementedError("only C-order zarr chunks are supported")
        if meta.get("filters"):
            raise NotImplementedError("zarr filters are not supported")
        super().__init__(meta["shape"])
        self.path = path
        self.chunks = tuple(int(c) for c in meta["chunks"])
        self.raw_dtype = np.dtype(meta["dtype"])
        self.sep = meta.get("dimension_separator", ".")
        self._inflate = _decompressor(meta.get("compressor"))
        fv = meta.get("fill_value")
        self.fill = np.nan if fv in ("NaN", None) else (np.inf if fv == "Infinity" else (-np.inf if fv == "-Infinity" else fv))
        self._decode = _cf_decoder(attrs, self.raw_dtype)
        self.lead_chunk = self.chunks[1] if len(self.chunks) > 1 else 1
        self.workers = workers or min(8, os.cpu_count() or 1)

    def _chunk(self, idx):
        fn = os.path.join(self.path, self.sep.join(str(i) for i in idx))
        if not os.path.exists(fn):
            return None
        with open(fn, "rb") as f:
            raw = self._inflate(f.read())
        return np.frombuffer(raw, dtype=self.raw_dtype).reshape(self.chunks)

    def read_rows(self, r0, r1, out=None):
        out = self._out(r0, r1, out)
        if r1 <= r0:
            return out
        nd = len(self.shape)
        grid = [range(-(-self.shape[d] // self.chunks[d])) for d in range(nd)]
        grid[1] = range(r0 // self.chunks[1], -(-r1 // self.chunks[1]))
        jobs = [idx for idx in np.ndindex(*[len(g) for g in grid])]

        def one(k):
            idx = tuple(grid[d][k[d]] for d in range(nd))
            lo = [idx[d] * self.chunks[d] for d in range(nd)]
            hi = [min(self.shape[d], lo[d] + self.chunks[d]) for d in range(nd)]
            lo1, hi1 = max(lo[1], r0), min(hi[1], r1)
            src = [slice(0, hi[d] - lo[d]) for d in range(nd)]
            src[1] = slice(lo1 - lo[1], hi1 - lo[1])
            dst = [slice(lo[d], hi[d]) for d in range(nd)]
            dst[1] = slice(lo1 - r0, hi1 - r0)
            c = self._chunk(idx)
            if c is None:                         # absent chunk = fill_value (a RAW value: decoded like data)
                fill = self.fill if self.raw_dtype.kind == "f" or not isinstance(self.fill, float) or \
                    np.isfinite(self.fill) else 0
                c = np.full(self.chunks, fill, dtype=self.raw_dtype)
            self._decode(c[tuple(src)], out[tuple(dst)])

        if self.workers > 1 and len(jobs) > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(self.workers) as ex:
                list(ex.map(one, jobs))
        else:
            for k in jobs:
                one(k)
        return out


def _zarr_json(path, name):
    fn = os.path.join(path, name)
    if not os.path.exists(fn):
        return None
    with open(fn) as f:
        return json.load(f)


def _zarr_small(path):
    """A whole (small) zarr array as numpy: coordinates."""
    meta = _zarr_json(path, ".zarray")
    if meta is None:
        return None, {}
    attrs = _zarr_json(path, ".zattrs") or {}
    shape = tuple(meta["shape"])
    chunks = tuple(meta["chunks"])
    dt = np.dtype(meta["dtype"])
    inflate = _decompressor(meta.get("compressor"))
    sep = meta.get("dimension_separator", ".")
    out = np.zeros(shape, dt.newbyteorder("="))
    for k in np.ndindex(*[-(-s // c) for s, c in zip(shape, chunks)]):
        fn = os.path.join(path, sep.join(str(i) for i in k)) if shape else os.path.join(path, "0")
        if not os.path.exists(fn):
            continue
        with open(fn, "rb") as f:
            c = np.frombuffer(inflate(f.read()), dtype=dt).reshape(chunks)
        sl = tuple(slice(i * cs, min(s, (i + 1) * cs)) for i, cs, s in zip(k, chunks, shape))
        out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
    return out, attrs


def open_zarr(path: str, variable: str | None = None, workers: int | None = None) -> Field:
    """Open one array of a zarr v2 directory store (group or single array) as a lazy host-backed Field --
    ``xr.open_zarr(path)[variable]`` with the default CF decoding."""
    if os.path.exists(os.path.join(path, ".zarray")):
        root, apath = os.path.dirname(os.path.abspath(path)), path
        variable = variable or os.path.basename(os.path.abspath(path))
    else:
        root = path
        names = sorted(n for n in os.listdir(path) if os.path.exists(os.path.join(path, n, ".zarray")))
        if variable is None:
            def _dims(n):
                return (_zarr_json(os.path.join(path, n), ".zattrs") or {}).get("_ARRAY_DIMENSIONS", [])
            cand = [n for n in names if len(_dims(n)) > 1 and "time" in _dims(n)]
            if len(cand) != 1:
                raise ValueError(f"{path}: give `variable`, one of {names}")
            variable = cand[0]
        apath = os.path.join(path, variable)
        if not os.path.exists(os.path.join(apath, ".zarray")):
            raise KeyError(f"{path} has no array {variable!r}")
    meta = _zarr_json(apath, ".zarray")
    attrs = _zarr_json(apath, ".zattrs") or {}
    dims = tuple(attrs.get("_ARRAY_DIMENSIONS") or (("time",) + tuple(f"d{i}" for i in range(len(meta["shape"]) - 1))))
    if "time" in dims and dims[0] != "time":
        raise ValueError(f"{variable}: `time` must be the first dimension, found {dims}")
    coords, tvals, tattrs = {}, None, {}
    for d in dims:
        cpath = os.path.join(root, d)
        if cpath != os.path.abspath(apath) and os.path.exists(os.path.join(cpath, ".zarray")):
            cv, cattrs = _zarr_small(cpath)
            if d == "time":
                tvals, tattrs = cv, cattrs
            else:
                coords[d] = cv
    src = _ZarrArray(apath, meta, attrs, workers)
    return _field_of(src, dims, tvals, tattrs, coords, _clean_attrs(attrs), variable)


def _write_zarr_array(path, vals, chunks, compressor, attrs):
    os.makedirs(path, exist_ok=True)
    vals = np.asarray(vals)
    chunks = tuple(int(min(max(1, c), max(1, s))) for c, s in zip(chunks, vals.shape))
    dt = vals.dtype.newbyteorder("<") if vals.dtype.byteorder != "|" else vals.dtype
    fill = "NaN" if vals.dtype.kind == "f" else 0
    meta = {"zarr_format": 2, "shape": list(vals.shape), "chunks": list(chunks), "dtype": dt.str, "order": "C",
            "compressor": compressor, "fill_value": fill, "filters": None}
    with open(os.path.join(path, ".zarray"), "w") as f:
        json.dump(meta, f)
    with open(os.path.join(path, ".zattrs"), "w") as f:
        json.dump(attrs, f)
    deflate = _compressor(compressor)
    for k in np.ndindex(*[-(-s // c) for s, c in zip(vals.shape, chunks)]):
        sl = tuple(slice(i * c, min(s, (i + 1) * c)) for i, c, s in zip(k, chunks, vals.shape))
        block = np.full(chunks, np.nan if vals.dtype.kind == "f" else 0, dtype=dt)      # edge chunks are padded
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = vals[sl]
        with open(os.path.join(path, ".".join(str(i) for i in k)), "wb") as f:
            f.write(deflate(block.tobytes()))


def save_zarr(path: str, field, name: str | None = None, chunks=None, compressor={"id": "zlib", "level": 1},
              calendar: str | None = None) -> str:
    """Write a Field / DataArray into a zarr v2 group ``path`` (array ``name`` + its coordinates)."""
    vals = field.numpy() if isinstance(field, Field) else np.asarray(field.values)
    dims = tuple(field.dims)
    name = name or getattr(field, "name", None) or "data"
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, ".zgroup"), "w") as f:
        json.dump({"zarr_format": 2}, f)
    if chunks is None:
        chunks = vals.shape
    attrs = {k: v for k, v in dict(field.attrs).items() if _plain(v) and k not in _CF_DROP}
    attrs["_ARRAY_DIMENSIONS"] = list(dims)
    _write_zarr_array(os.path.join(path, name), vals, chunks, compressor, attrs)
    tv = _time_values(field, calendar) if "time" in dims else None
    if tv is not None:
        _write_zarr_array(os.path.join(path, "time"), tv[0], (len(tv[0]),), None,
                          {"_ARRAY_DIMENSIONS": ["time"], "units": tv[1], "calendar": tv[2]})
    for d in dims:
        c = getattr(field, "coords", {}).get(d) if d != "time" else None
        if c is not None and np.asarray(c).dtype.kind in "fiu":
            _write_zarr_array(os.path.join(path, d), np.asarray(c), (len(c),), None, {"_ARRAY_DIMENSIONS": [d]})
    return path


# ------------------------------------------------------------------------------------------------ dispatch


def open_field(path: str, variable: str | None = None, **kwargs) -> Field:
    """Open ``path`` by its kind: zarr directory store, ``.npy`` or NetCDF-3 file."""
    if os.path.isdir(path):
        return open_zarr(path, variable, **kwargs)
    if path.endswith(".npy"):
        return open_npy(path, **kwargs)
    with open(path, "rb") as f:
        magic = f.read(4)
    if magic[:3] == b"CDF":
        return open_netcdf3(path, variable, **kwargs)
    if magic[1:4] == b"HDF":
        raise NotImplementedError(f"{path} is a NetCDF-4 / HDF5 file: no HDF5 library in this image "
                                  "(convert with `nccopy -k cdf5`/`-k 64-bit-offset`, or to zarr)")
    raise ValueError(f"{path}: unknown file kind")
