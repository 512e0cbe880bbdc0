"""End-to-end slab streaming: host arrays -> lat slabs in HBM -> index kernels -> host result.

The reference hands host (numpy / dask) arrays to its index functions (core/indicator.py:884-886
``self.compute(**args)``); dask walks them chunk by chunk (indices/run_length.py:209-216,
indices/helpers.py:949-974).  Every value on this path depends on ONE grid cell's series, so the
device analogue is: cut the leading spatial dimension (lat) into slabs, copy slab k+1 host -> device
on a copy stream (``xc_copy_box_async``: one strided box per input) while the kernels of slab k run,
and copy each slab's small result back.  :func:`streamed` wraps an index function with exactly that;
inputs that are already device-resident, or small, take the direct path unchanged.

PyTorch is plumbing here (device buffers, streams, events); the copies and the kernels go through
the C ABI.
"""
from __future__ import annotations

import functools

import numpy as np

from .field import Field, attrs_of, dims_of, is_xarray, raw_values, time_axis_of
from .options import OPTIONS, set_options

#: inputs below this many bytes are unwrapped in one piece (device.to_time_cell)
STREAM_MIN_BYTES = 256 << 20
#: target size of one slab of one input in HBM (two slabs per input are resident)
SLAB_BYTES = 1 << 30

_staging: dict = {}


def plan_slabs(n_rows: int, bytes_per_row: int, slab_bytes: int = None, align: int = 1) -> list[tuple[int, int]]:
    """Contiguous [r0, r1) ranges of the leading spatial dimension, each about ``slab_bytes``.
    With ``align`` > 1 (the chunk length of a file-backed input along that dimension) every slab edge
    but the last is a multiple of it, so that no chunk of the file is read or inflated twice."""
    slab_bytes = SLAB_BYTES if slab_bytes is None else slab_bytes
    rows = int(max(1, min(n_rows, slab_bytes // max(1, bytes_per_row))))
    n = -(-n_rows // rows)
    rows = -(-n_rows // n)            # even out: no tiny last slab
    if align > 1:
        rows = max(align, rows // align * align)
    return [(r, min(n_rows, r + rows)) for r in range(0, n_rows, rows)]


def _host_array(obj):
    """The numpy view (or lazy file-backed source, io.LazyGrid) of a host-backed input, or None
    (device tensor, unsupported container)."""
    from .io import LazyGrid
    v = raw_values(obj)
    if hasattr(v, "is_cuda"):
        if v.is_cuda:
            return None
        return v.numpy()
    return v if isinstance(v, (np.ndarray, LazyGrid)) else None


def _is_lazy(h):
    from .io import LazyGrid
    return isinstance(h, LazyGrid)


def _is_labelled(a):
    return isinstance(a, Field) or is_xarray(a)


def _classify(args, kwargs):
    """Split the labelled inputs of a call into series (time first) and per-cell tables."""
    series, tables = {}, {}
    items = [(("a", i), a) for i, a in enumerate(args)] + [(("k", k), v) for k, v in kwargs.items()]
    for key, a in items:
        if not _is_labelled(a):
            continue
        dims = dims_of(a)
        if "time" in dims:
            series[key] = a
        else:
            tables[key] = a
    return series, tables


def _streamable(series, tables):
    if not series:
        return None
    lead, shape = None, None
    total = 0
    for a in series.values():
        dims = dims_of(a)
        h = _host_array(a)
        if h is None or dims[0] != "time" or len(dims) < 2 or h.dtype != np.float32 or \
                not (_is_lazy(h) or h.flags.c_contiguous):
            return None
        if lead is None:
            lead, shape = dims[1], h.shape
        elif dims[1] != lead or h.shape != shape:
            return None
        total = max(total, h.nbytes)
    if total < OPTIONS.get("stream_min_bytes", STREAM_MIN_BYTES):
        return None
    for a in tables.values():
        dims = dims_of(a)
        h = _host_array(a)
        if h is None or _is_lazy(h) or lead not in dims or dims[0] != lead or h.shape[0] != shape[1]:
            return None
    return lead, shape


def _pinned(name, nbytes):
    import torch
    buf = _staging.get(name)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        _staging[name] = buf
    return buf


def _slice_lead(a, r0, r1, values):
    """The slab [r0, r1) of a labelled input as a Field around ``values``."""
    dims = dims_of(a)
    coords = {}
    if isinstance(a, Field):
        coords = {k: v for k, v in a.coords.items() if k not in dims}
    else:   # xarray: scalar / non-dimension coordinates the wrappers look at
        for k in ("percentiles",):
            if k in a.coords and k not in dims:
                coords[k] = a.coords[k].values
    ta = time_axis_of(a) if "time" in dims else None
    return Field(values, dims, ta, coords, attrs_of(a), getattr(a, "name", None))


def run_streamed(fn, args, kwargs, lead, shape, series, tables):
    import torch

    from . import device
    from ._lib import check, load
    from .io import parallel_rows

    device._require_cuda()
    lib = load()
    T, n_lead = shape[0], shape[1]
    rest = int(np.prod(shape[2:], dtype=np.int64)) if len(shape) > 2 else 1
    keys = list(series)
    hosts = {k: _host_array(series[k]) for k in keys}
    align = int(np.lcm.reduce([int(getattr(h, "lead_chunk", 1)) for h in hosts.values()]))
    slabs = plan_slabs(n_lead, T * rest * 4, OPTIONS.get("stream_slab_bytes", SLAB_BYTES), align)
    rows_max = max(b - a for a, b in slabs)
    comp = torch.cuda.current_stream()
    s_copy, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    s_copy.wait_stream(comp)
    bufs = {k: [torch.empty(T * rows_max * rest, dtype=torch.float32, device="cuda") for _ in range(2)] for k in keys}
    copied = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    # Inputs that are not page-locked (plain numpy arrays, memory-mapped files: the I/O step of SURVEY.md 8f.3)
    # are staged by a reader thread: slab k+2 is gathered from the file / pageable memory into a pinned buffer
    # while slab k+1 crosses PCIe and slab k is computed, so disk, PCIe and kernels overlap.
    # File-backed lazy sources (io.LazyGrid: NetCDF-3, zarr) are read and decoded by the same thread.
    staged = {k: _is_lazy(hosts[k]) or not lib.xc_host_pinned(hosts[k].ctypes.data) for k in keys}
    stage_bufs, pool, futures = {}, None, {}
    if any(staged.values()):
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=1)
        for key in keys:
            if staged[key]:
                stage_bufs[key] = [_pinned(f"in{len(stage_bufs)}_{i}", T * rows_max * rest * 4) for i in range(2)]

    def fill_stage(k):
        r0, r1 = slabs[k]
        b = k & 1
        if k >= 2:
            copied[b].synchronize()               # the H2D copy of slab k-2 has drained this staging buffer
        for key in keys:
            if staged[key]:
                dst = stage_bufs[key][b][: T * (r1 - r0) * rest * 4].numpy().view(np.float32)
                dst = dst.reshape((T, r1 - r0) + tuple(shape[2:]))
                if _is_lazy(hosts[key]):
                    hosts[key].read_rows(r0, r1, out=dst)
                else:      # strided box of pageable / memory-mapped memory: time ranges on the copy threads
                    src = hosts[key]
                    parallel_rows(lambda a, b, s_=src, d_=dst: np.copyto(d_[a:b], s_[a:b, r0:r1]), T, dst.nbytes)

    def submit_stage(k):
        if pool is not None and k < len(slabs):
            futures[k] = pool.submit(fill_stage, k)

    def issue_copy(k):
        r0, r1 = slabs[k]
        b = k & 1
        if k in futures:
            futures.pop(k).result()
        if k >= 2:
            s_copy.wait_event(done[b])            # the kernels of slab k-2 have released the buffer
        w = (r1 - r0) * rest * 4
        for key in keys:
            if staged[key]:
                check(lib.xc_copy_box_async(bufs[key][b].data_ptr(), w, stage_bufs[key][b].data_ptr(), w, w, T, 1,
                                            s_copy.cuda_stream))
            else:
                h = hosts[key]
                check(lib.xc_copy_box_async(bufs[key][b].data_ptr(), w, h.ctypes.data + r0 * rest * 4,
                                            n_lead * rest * 4, w, T, 1, s_copy.cuda_stream))
        copied[b].record(s_copy)

    template = next(iter(series.values()))
    outs: dict = {}                               # one _SlabOutput per result of the call
    shape_of = None                               # how to rebuild the result container (Field / tuple / dict)
    keep = []                                     # device results stay alive until their D2H has run
    submit_stage(0)
    submit_stage(1)
    issue_copy(0)
    with set_options(device_outputs=True):
        for k, (r0, r1) in enumerate(slabs):
            b = k & 1
            if k + 1 < len(slabs):
                issue_copy(k + 1)                 # prefetch before the (possibly synchronising) index call
            submit_stage(k + 2)
            comp.wait_event(copied[b])
            a2, k2 = list(args), dict(kwargs)
            for key in keys:
                v = bufs[key][b][: T * (r1 - r0) * rest].view((T, r1 - r0) + tuple(shape[2:]))
                f = _slice_lead(series[key], r0, r1, v)
                if key[0] == "a":
                    a2[key[1]] = f
                else:
                    k2[key[1]] = f
            for key, tab in tables.items():
                f = _slice_lead(tab, r0, r1, np.ascontiguousarray(_host_array(tab)[r0:r1]))
                if key[0] == "a":
                    a2[key[1]] = f
                else:
                    k2[key[1]] = f
            res = fn(*a2, **k2)
            done[b].record(comp)
            parts, shape_of = _result_parts(fn, res)
            for name, part in parts:
                if name not in outs:
                    outs[name] = _SlabOutput(fn, part, lead, n_lead, name)
                outs[name].put(part, r0, r1, lib, comp, s_out, keep)
    s_out.synchronize()
    comp.synchronize()
    s_copy.synchronize()
    if pool is not None:
        pool.shutdown(wait=True)
    del keep
    done_parts = {name: o.finish(template, lead) for name, o in outs.items()}
    return shape_of(done_parts)


def _result_parts(fn, res):
    """``[(name, Field), ...]`` of what an index returned and the function that rebuilds the same kind of
    container from ``{name: assembled}``: a single Field, a (named) tuple of Fields (``cffwis_indices``) or a
    dict of Fields (``fire_weather_ufunc``)."""
    if isinstance(res, Field):
        return [(None, res)], lambda d: d[None]
    if isinstance(res, dict) and res and all(isinstance(v, Field) for v in res.values()):
        names = list(res)
        return list(res.items()), lambda d: {n: d[n] for n in names}
    if isinstance(res, tuple) and res and all(isinstance(v, Field) for v in res):
        n = len(res)
        rebuild = (lambda d: type(res)(*(d[i] for i in range(n)))) if hasattr(res, "_fields") else \
            (lambda d: tuple(d[i] for i in range(n)))
        return list(enumerate(res)), rebuild
    raise TypeError(f"{getattr(fn, '__name__', fn)} returned {type(res).__name__}: cannot stream it")


def _on_device(vals) -> bool:
    return bool(getattr(vals, "is_cuda", False))


class _SlabOutput:
    """The host array one result of a streamed call is assembled in, slab by slab."""

    def __init__(self, fn, res, lead, n_lead, name):
        import torch
        vals = res.values
        rdims = tuple(res.dims)
        if lead not in rdims:
            raise ValueError(f"the result of {fn.__name__} has no `{lead}` dimension: cannot stream it")
        self.axis = rdims.index(lead)
        self.n_lead = n_lead
        vshape = tuple(vals.shape)
        full = vshape[:self.axis] + (n_lead,) + vshape[self.axis + 1:]
        np_dtype = np.dtype(str(vals.dtype).replace("torch.", "")) if hasattr(vals, "is_cuda") else vals.dtype
        nbytes = int(np.prod(full, dtype=np.int64)) * np_dtype.itemsize
        self.meta = res
        if nbytes >= (256 << 20):
            # large results (percentile tables, daily series): the result array itself is page-locked, the slabs
            # land in it directly (numpy keeps the torch storage alive through the buffer protocol)
            own = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self.host = self.stage = own.numpy().view(np_dtype).reshape(full)
        else:
            self.host = np.empty(full, dtype=np_dtype)
            out_stage = _pinned("out" if name is None else f"out_{name}", nbytes)
            self.stage = out_stage[:nbytes].numpy().view(np_dtype).reshape(full)

    def put(self, res, r0, r1, lib, comp, s_out, keep):
        from ._lib import check
        vals = res.values
        A = int(np.prod(self.host.shape[:self.axis], dtype=np.int64))
        B = int(np.prod(self.host.shape[self.axis + 1:], dtype=np.int64)) * self.host.itemsize
        if _on_device(vals):
            vals = vals.contiguous()
            keep.append(vals)
            s_out.wait_stream(comp)
            dst = self.stage.ctypes.data + r0 * B
            check(lib.xc_copy_box_async(dst, self.n_lead * B, vals.data_ptr(), (r1 - r0) * B, (r1 - r0) * B, A, 0,
                                        s_out.cuda_stream))
        else:   # the index returned host values (e.g. percentile tables)
            idx = [slice(None)] * self.host.ndim
            idx[self.axis] = slice(r0, r1)
            self.stage[tuple(idx)] = np.asarray(vals)

    def finish(self, template, lead):
        if self.host is not self.stage:
            np.copyto(self.host, self.stage)
        return _assemble(template, self.meta, self.host, lead)


def _assemble(template, res, values, lead):
    """The full-grid container of the same family as the inputs (labels of the sliced dimension and
    the other spatial coordinates come from the template; the rest from the slab results)."""
    dims = tuple(res.dims)
    if is_xarray(template):
        import xarray as xr
        coords = {d: template.coords[d] for d in dims if d != "time" and d in template.coords}
        for k, v in res.coords.items():
            if k not in coords:
                coords[k] = v
        return xr.DataArray(values, dims=dims, coords=coords, attrs=dict(res.attrs), name=res.name)
    coords = dict(res.coords)
    for d in dims:
        if d != "time" and d in getattr(template, "coords", {}):
            coords[d] = template.coords[d]
    return Field(values, dims, res.time, coords, dict(res.attrs), res.name)


def streamed(fn):
    """Route host-backed inputs of an index function through the slab streamer (see module doc)."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if OPTIONS.get("device_outputs") or OPTIONS.get("_in_stream"):
            return fn(*args, **kwargs)
        series, tables = _classify(args, kwargs)
        plan = _streamable(series, tables)
        if plan is None:
            return fn(*args, **kwargs)
        OPTIONS["_in_stream"] = True
        try:
            return run_streamed(fn, args, kwargs, plan[0], plan[1], series, tables)
        finally:
            OPTIONS["_in_stream"] = False

    wrapper.__wrapped_index__ = fn
    return wrapper
