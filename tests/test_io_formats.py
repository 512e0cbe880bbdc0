"""SURVEY.md 8f.3 (host side): NetCDF-3 and zarr v2 readers / writers, CF decoding, chunk-aligned slab
plans, and the slab streamer fed by lazy file-backed sources (CPU stand-ins: tests/fake_stream.py)."""
import gzip
import json
import os
import zlib

import numpy as np
import pytest

from xb_helpers import make_field


def _pr(rng, T=365 * 2, shape=(11, 16)):
    pr = rng.gamma(0.4, 6.0, size=(T,) + shape).astype(np.float32)
    pr[rng.random(pr.shape) < 0.5] = 0
    pr[rng.random(pr.shape) < 0.002] = np.nan
    return pr


def test_netcdf3_round_trip_is_lazy_and_big_endian_on_disk(tmp_path):
    from xclim_b200 import io
    rng = np.random.default_rng(5)
    pr = _pr(rng)
    f = make_field(pr, "1981-01-01", calendar="noleap", units="mm/d", dims=("time", "lat", "lon"))
    f.coords["lat"] = np.linspace(-45, 45, pr.shape[1])
    f.coords["lon"] = np.arange(pr.shape[2], dtype=np.float32) * 0.25
    path = io.save_netcdf3(str(tmp_path / "pr.nc"), f, name="pr")
    with open(path, "rb") as fh:
        assert fh.read(4) == b"CDF\x02"                 # 64-bit offset: variables beyond 2 GiB
    lazy = io.open_netcdf3(path)
    assert isinstance(lazy.values, io.LazyGrid) and lazy.values._data.dtype == np.dtype(">f4")
    assert lazy.dims == ("time", "lat", "lon") and lazy.shape == pr.shape and lazy.name == "pr"
    assert lazy.attrs["units"] == "mm/d" and lazy.time.calendar == "noleap" and len(lazy.time) == pr.shape[0]
    assert lazy.time.date_strings(0)[0] == "1981-01-01" and lazy.time.date_strings(364)[0] == "1981-12-31"
    np.testing.assert_array_equal(lazy.coords["lat"], f.coords["lat"])
    np.testing.assert_array_equal(np.asarray(lazy.values), pr)
    rows = lazy.values.read_rows(3, 7)
    assert rows.dtype == np.float32 and rows.flags.c_contiguous and rows.dtype.isnative
    np.testing.assert_array_equal(rows, pr[:, 3:7])
    np.testing.assert_array_equal(lazy.values[10:20, 2:5, 1], pr[10:20, 2:5, 1])
    assert io.open_field(path).shape == pr.shape


def test_cf_packed_variable_is_unpacked_like_xarray(tmp_path):
    """int16 packed data with scale_factor / add_offset / _FillValue (the usual ERA5 NetCDF layout):
    fill -> NaN, value = raw * scale + offset computed in float64 and rounded once to float32."""
    from scipy.io import netcdf_file

    from xclim_b200 import io
    rng = np.random.default_rng(6)
    raw = rng.integers(-32000, 32000, size=(30, 5, 6)).astype(np.int16)
    raw[rng.random(raw.shape) < 0.05] = -32767
    path = str(tmp_path / "t2m.nc")
    with netcdf_file(path, "w") as nc:
        nc.createDimension("time", None); nc.createDimension("lat", 5); nc.createDimension("lon", 6)
        t = nc.createVariable("time", "i4", ("time",)); t[:] = np.arange(30) * 24 + 12
        t.units = "hours since 2000-02-25 00:00:00"; t.calendar = "gregorian"
        v = nc.createVariable("t2m", "h", ("time", "lat", "lon")); v[:] = raw
        v.scale_factor = np.float64(0.0011); v.add_offset = np.float64(271.3); v._FillValue = np.int16(-32767)
        v.units = "K"
        w = nc.createVariable("t2m32", "h", ("time", "lat", "lon")); w[:] = raw
        w.scale_factor = np.float32(0.0011); w.add_offset = np.float32(271.3)
    f = io.open_netcdf3(path, "t2m")
    exp = (raw.astype(np.float64) * 0.0011 + 271.3).astype(np.float32)
    exp[raw == -32767] = np.nan
    np.testing.assert_array_equal(np.asarray(f.values), exp)
    # float32 attributes on 2-byte integers: xarray unpacks in float32 (coding/variables.py _choose_float_dtype)
    g = io.open_netcdf3(path, "t2m32")
    np.testing.assert_array_equal(np.asarray(g.values), raw.astype(np.float32) * np.float32(0.0011) + np.float32(271.3))
    assert "scale_factor" not in f.attrs and f.attrs["units"] == "K"
    # 2000 is a leap year in the gregorian calendar: Feb 25 + 5 days = Mar 1
    assert f.time.date_strings(0)[0] == "2000-02-25" and f.time.date_strings(5)[0] == "2000-03-01"


@pytest.mark.parametrize("compressor", [None, {"id": "zlib", "level": 1}, {"id": "gzip", "level": 1}, {"id": "bz2", "level": 1}])
def test_zarr_v2_round_trip_with_edge_chunks(tmp_path, compressor):
    from xclim_b200 import io
    rng = np.random.default_rng(7)
    pr = _pr(rng, T=200, shape=(11, 16))
    f = make_field(pr, "1981-01-01", calendar="360_day", units="mm/d", dims=("time", "lat", "lon"))
    f.coords["lat"] = np.linspace(-45, 45, 11)
    store = io.save_zarr(str(tmp_path / "ds.zarr"), f, name="pr", chunks=(64, 4, 10), compressor=compressor)
    meta = json.load(open(os.path.join(store, "pr", ".zarray")))
    assert meta["chunks"] == [64, 4, 10] and meta["zarr_format"] == 2 and meta["dtype"] == "<f4"
    assert sorted(n for n in os.listdir(os.path.join(store, "pr")) if n[0] != ".")[:3] == ["0.0.0", "0.0.1", "0.1.0"]
    lazy = io.open_zarr(store)
    assert lazy.values.lead_chunk == 4 and lazy.dims == ("time", "lat", "lon") and lazy.time.calendar == "360_day"
    np.testing.assert_array_equal(np.asarray(lazy.values), pr)
    np.testing.assert_array_equal(lazy.values.read_rows(4, 11), pr[:, 4:11])      # chunk-aligned start, ragged end
    np.testing.assert_array_equal(lazy.values.read_rows(5, 6), pr[:, 5:6])        # inside one chunk
    np.testing.assert_array_equal(lazy.coords["lat"], f.coords["lat"])
    assert io.open_field(store, "pr").shape == pr.shape


def test_zarr_chunks_written_by_hand_missing_chunk_and_slash_separator(tmp_path):
    """A store laid out as the zarr v2 spec says (not by our writer): '/' separator, zlib chunks, one
    chunk absent (-> fill_value), int16 packed with CF attributes."""
    from xclim_b200 import io
    rng = np.random.default_rng(8)
    raw = rng.integers(-1000, 1000, size=(6, 4, 4)).astype("<i2")
    root = tmp_path / "g.zarr"
    arr = root / "tas"
    os.makedirs(arr)
    json.dump({"zarr_format": 2}, open(root / ".zgroup", "w"))
    json.dump({"zarr_format": 2, "shape": [6, 4, 4], "chunks": [6, 2, 4], "dtype": "<i2", "order": "C",
               "compressor": {"id": "zlib", "level": 5}, "fill_value": -9999, "filters": None,
               "dimension_separator": "/"}, open(arr / ".zarray", "w"))
    json.dump({"_ARRAY_DIMENSIONS": ["time", "lat", "lon"], "scale_factor": 0.01, "add_offset": 280.0,
               "_FillValue": -9999, "units": "K"}, open(arr / ".zattrs", "w"))
    os.makedirs(arr / "0" / "0")
    open(arr / "0" / "0" / "0", "wb").write(zlib.compress(np.ascontiguousarray(raw[:, 0:2]).tobytes()))
    # chunk 0/1/0 is absent: fill_value
    tdir = root / "time"
    os.makedirs(tdir)
    json.dump({"zarr_format": 2, "shape": [6], "chunks": [6], "dtype": "<i8", "order": "C", "compressor": {"id": "gzip"},
               "fill_value": 0, "filters": None}, open(tdir / ".zarray", "w"))
    json.dump({"_ARRAY_DIMENSIONS": ["time"], "units": "days since 1999-12-30", "calendar": "noleap"},
              open(tdir / ".zattrs", "w"))
    open(tdir / "0", "wb").write(gzip.compress(np.arange(6, dtype="<i8").tobytes()))
    f = io.open_zarr(str(root))
    exp = np.full((6, 4, 4), np.nan, np.float32)
    exp[:, 0:2] = (raw[:, 0:2].astype(np.float64) * 0.01 + 280.0).astype(np.float32)
    np.testing.assert_array_equal(np.asarray(f.values), exp)
    assert f.time.date_strings(2)[0] == "2000-01-01" and f.attrs == {"units": "K"}


def test_unsupported_stores_say_so(tmp_path):
    from xclim_b200 import io
    arr = tmp_path / "b.zarr"
    os.makedirs(arr)
    json.dump({"zarr_format": 2, "shape": [2, 2], "chunks": [2, 2], "dtype": "<f4", "order": "C",
               "compressor": {"id": "blosc", "cname": "lz4"}, "fill_value": "NaN", "filters": None}, open(arr / ".zarray", "w"))
    with pytest.raises(NotImplementedError, match="blosc"):
        io.open_zarr(str(arr))
    h5 = tmp_path / "x.nc"
    open(h5, "wb").write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(NotImplementedError, match="HDF5"):
        io.open_field(str(h5))
    with pytest.raises(ValueError, match="gap-free daily"):
        io.decode_cf_time([0, 1, 3], "days since 2000-01-01", "standard")


def test_period_results_are_written_with_their_dates(tmp_path):
    """A (periods, lat, lon) result with ISO labels goes to disk with a CF time coordinate (to_netcdf leg)."""
    from scipy.io import netcdf_file

    from xclim_b200 import Field, io
    vals = np.arange(3 * 2 * 2, dtype=np.float32).reshape(3, 2, 2)
    res = Field(vals, ("time", "lat", "lon"), None, {"time": ["1981-01-01", "1982-01-01", "1983-01-01"]}, {"units": "days"})
    p = io.save_netcdf3(str(tmp_path / "cdd.nc"), res, name="cdd", calendar="noleap")
    with netcdf_file(p, "r", mmap=False) as nc:
        np.testing.assert_array_equal(nc.variables["time"][:], [0, 365, 730])
        assert nc.variables["time"].calendar == b"noleap" and nc.variables["cdd"].units == b"days"
        np.testing.assert_array_equal(nc.variables["cdd"][:], vals)
    z = io.save_zarr(str(tmp_path / "cdd.zarr"), res, name="cdd")
    tv, ta = io._zarr_small(os.path.join(z, "time"))
    np.testing.assert_array_equal(tv, [0, 365, 730])       # standard calendar: 1981 and 1982 are not leap years
    assert ta["calendar"] == "standard" and ta["units"].startswith("days since 1981-01-01")


def test_plan_slabs_follow_chunk_edges():
    from xclim_b200 import streaming
    for n, align in ((721, 32), (11, 4), (90, 7), (5, 8)):
        for slab in (1, 10 ** 6, 10 ** 12):
            pl = streaming.plan_slabs(n, 10 ** 5, slab, align)
            assert pl[0][0] == 0 and pl[-1][1] == n and all(a[1] == b[0] for a, b in zip(pl, pl[1:]))
            assert all(a % align == 0 for a, _ in pl)


@pytest.mark.parametrize("kind", ["netcdf3", "zarr", "npy"])
def test_streamer_reads_lazy_sources_slab_by_slab(tmp_path, monkeypatch, kind):
    """File on disk -> lazy Field -> slab streamer (reader thread decodes each slab into the staging
    buffer) -> index function per slab -> assembled host result -> file on disk; equal to the in-memory
    call.  CPU stand-ins for the device layer and the CUDA plumbing (tests/fake_stream.py)."""
    import fake_stream
    import xclim_b200
    from xclim_b200 import atmos, indices, io, streaming
    lib = fake_stream.install(monkeypatch)
    rng = np.random.default_rng(9)
    pr = _pr(rng, T=365 * 2, shape=(11, 16))
    f_pr = make_field(pr, "1981-01-01", calendar="noleap", units="mm/d", dims=("time", "lat", "lon"))
    f_pr.coords["lat"] = np.linspace(-50, 50, 11)
    if kind == "netcdf3":
        lazy = io.open_field(io.save_netcdf3(str(tmp_path / "pr.nc"), f_pr, name="pr"))
    elif kind == "zarr":
        lazy = io.open_field(io.save_zarr(str(tmp_path / "pr.zarr"), f_pr, name="pr", chunks=(365, 4, 16)))
    else:
        lazy = io.open_field(io.save_npy(str(tmp_path / "pr.npy"), f_pr))
    reads = []
    if kind != "npy":
        real = lazy.values.read_rows
        monkeypatch.setattr(lazy.values, "read_rows", lambda r0, r1, out=None: (reads.append((r0, r1)), real(r0, r1, out))[1])
    ref_cdd = atmos.maximum_consecutive_dry_days(f_pr)
    ref_wet = indices.wetdays(f_pr)
    row = pr.shape[0] * pr.shape[2] * 4
    with xclim_b200.set_options(stream_min_bytes=0, stream_slab_bytes=2 * row):
        series, tables = streaming._classify((lazy,), {})
        assert streaming._streamable(series, tables) == ("lat", pr.shape)
        out = atmos.maximum_consecutive_dry_days(lazy)
        n_boxes = len(lib.boxes)
        wet = indices.wetdays(lazy)
    np.testing.assert_array_equal(out.values, ref_cdd.values)
    np.testing.assert_array_equal(wet.values, ref_wet.values)
    assert out.dims == ("time", "lat", "lon") and out.attrs["units"] == "days"
    np.testing.assert_array_equal(out.coords["lat"], f_pr.coords["lat"])
    if kind == "zarr":      # slabs follow the 4-row chunks: 11 rows -> [0,4) [4,8) [8,11), each chunk read once per call
        assert reads[:3] == [(0, 4), (4, 8), (8, 11)] and n_boxes == 3
    elif kind == "netcdf3":
        assert reads[:6] == [(0, 2), (2, 4), (4, 6), (6, 8), (8, 10), (10, 11)] and n_boxes == 6
    # below the streaming threshold the lazy source is materialised once and takes the direct path
    small = atmos.maximum_consecutive_dry_days(lazy)
    np.testing.assert_array_equal(small.values, ref_cdd.values)
    # the to_netcdf / to_zarr leg
    back = io.open_field(io.save_netcdf3(str(tmp_path / "cdd.nc"), out, name="cdd", calendar="noleap"))
    np.testing.assert_array_equal(np.asarray(back.values), ref_cdd.values)
    assert back.attrs["units"] == "days"


def test_slab_output_device_leg_and_multi_output_containers(monkeypatch):
    """The D2H leg of the streamer (one strided box per slab result into the page-locked result array) and the
    rebuilding of dict / namedtuple results, on the CPU stand-ins."""
    import collections

    import torch

    import fake_stream
    from xclim_b200 import Field, streaming
    lib = fake_stream.install(monkeypatch)
    rng = np.random.default_rng(12)
    full = rng.standard_normal((5, 11, 7)).astype(np.float32)          # (periods, lat, lon)
    flat = rng.integers(0, 9, size=(11, 7)).astype(np.int32)           # (lat, lon): no time dimension
    template = Field(np.zeros((3, 11, 7), np.float32), ("time", "lat", "lon"), None, {"lat": np.arange(11.0)}, {})

    def part(r0, r1):
        return {"a": Field(torch.from_numpy(full[:, r0:r1].copy()), ("time", "lat", "lon"), None, {"time": list("abcde")},
                           {"units": "K"}),
                "b": Field(torch.from_numpy(flat[r0:r1].copy()), ("lat", "lon"), None, {}, {"units": ""})}

    outs = {}
    keep = []
    for r0, r1 in ((0, 4), (4, 8), (8, 11)):
        parts, rebuild = streaming._result_parts(test_plan_slabs_follow_chunk_edges, part(r0, r1))
        for name, p in parts:
            if name not in outs:
                outs[name] = streaming._SlabOutput(test_plan_slabs_follow_chunk_edges, p, "lat", 11, name)
            outs[name].put(p, r0, r1, lib, fake_stream._Stream(), fake_stream._Stream(), keep)
    res = rebuild({n: o.finish(template, "lat") for n, o in outs.items()})
    assert list(res) == ["a", "b"] and res["a"].dims == ("time", "lat", "lon") and res["b"].dims == ("lat", "lon")
    np.testing.assert_array_equal(res["a"].values, full)
    np.testing.assert_array_equal(res["b"].values, flat)
    assert res["b"].values.dtype == np.int32 and res["a"].attrs == {"units": "K"}
    np.testing.assert_array_equal(res["a"].coords["lat"], np.arange(11.0))
    # device -> host boxes: 5 rows of (r1 - r0) * 7 * 4 bytes for "a", one row for "b"
    assert [b for b in lib.boxes if b[2] == 0][:2] == [(4 * 7 * 4, 5, 0), (4 * 7 * 4, 1, 0)]
    # containers
    Pair = collections.namedtuple("Pair", ["x", "y"])
    fa, fb = part(0, 11)["a"], part(0, 11)["b"]
    parts, rebuild = streaming._result_parts(len, Pair(fa, fb))
    assert [n for n, _ in parts] == [0, 1] and isinstance(rebuild({0: 1, 1: 2}), Pair)
    parts, rebuild = streaming._result_parts(len, fa)
    assert parts == [(None, fa)] and rebuild({None: 7}) == 7
    with pytest.raises(TypeError, match="cannot stream"):
        streaming._result_parts(len, 3)


def test_non_daily_time_axes_are_labels_only(tmp_path, monkeypatch):
    """A monthly file opens, but carries no daily TimeAxis: the hot path (which resamples daily series) refuses
    it with a clear error instead of mis-reading the steps as days."""
    from scipy.io import netcdf_file

    import fake_device
    from xclim_b200 import indices, io
    fake_device.install(monkeypatch)
    path = str(tmp_path / "monthly.nc")
    with netcdf_file(path, "w") as nc:
        nc.createDimension("time", None); nc.createDimension("lat", 2); nc.createDimension("lon", 3)
        t = nc.createVariable("time", "f8", ("time",)); t[:] = np.array([15.5, 45.0, 74.5, 105.0])
        t.units = "days since 2001-01-01"; t.calendar = "noleap"
        v = nc.createVariable("pr", "f4", ("time", "lat", "lon")); v[:] = np.ones((4, 2, 3), np.float32); v.units = "mm/d"
        w = nc.createVariable("orog", "f4", ("lat", "lon")); w[:] = np.zeros((2, 3), np.float32)
    f = io.open_netcdf3(path)                      # the only variable with a time dimension
    assert f.name == "pr" and f.time is None and list(f.coords["time"]) == [15.5, 45.0, 74.5, 105.0]
    with pytest.raises(ValueError, match="no time axis"):
        indices.wetdays(f)
    with pytest.raises(KeyError):
        io.open_netcdf3(path, "tas")
    o = io.open_netcdf3(path, "orog")
    assert o.dims == ("lat", "lon") and o.time is None and np.asarray(o.values).shape == (2, 3)


def test_parallel_rows_covers_every_row_once(monkeypatch):
    from xclim_b200 import io
    monkeypatch.setattr(io, "PARALLEL_MIN_BYTES", 0)
    monkeypatch.setattr(io, "copy_threads", lambda: 5)
    monkeypatch.setattr(io, "_copy_pool", None)
    for n in (1, 7, 10, 11, 97, 1000):
        hits = np.zeros(n, np.int32)

        def fn(a, b):
            hits[a:b] += 1
        io.parallel_rows(fn, n, 1)
        assert (hits == 1).all(), n
    with pytest.raises(ZeroDivisionError):          # a failing range is not swallowed
        io.parallel_rows(lambda a, b: 1 // 0, 100, 1)


@pytest.mark.parametrize("kind", ["netcdf3", "npy"])
def test_streamer_with_parallel_staging_copies(tmp_path, monkeypatch, kind):
    """The staging copy of a slab split over the copy threads (time ranges) gives the same bytes."""
    import fake_stream
    import xclim_b200
    from xclim_b200 import indices, io
    fake_stream.install(monkeypatch)
    monkeypatch.setattr(io, "PARALLEL_MIN_BYTES", 0)
    monkeypatch.setattr(io, "copy_threads", lambda: 4)
    monkeypatch.setattr(io, "_copy_pool", None)
    rng = np.random.default_rng(15)
    pr = _pr(rng, T=365 * 2, shape=(9, 16))
    f_pr = make_field(pr, "1981-01-01", calendar="noleap", units="mm/d", dims=("time", "lat", "lon"))
    if kind == "netcdf3":
        lazy = io.open_field(io.save_netcdf3(str(tmp_path / "pr.nc"), f_pr, name="pr"))
        np.testing.assert_array_equal(lazy.values.read_rows(2, 7), pr[:, 2:7])
    else:
        lazy = io.open_field(io.save_npy(str(tmp_path / "pr.npy"), f_pr))
    ref = indices.maximum_consecutive_dry_days(f_pr)
    with xclim_b200.set_options(stream_min_bytes=0, stream_slab_bytes=3 * pr.shape[0] * pr.shape[2] * 4):
        got = indices.maximum_consecutive_dry_days(lazy)
        got2 = indices.maximum_consecutive_dry_days(f_pr)       # a plain (pageable) numpy array is staged the same way
    np.testing.assert_array_equal(got.values, ref.values)
    np.testing.assert_array_equal(got2.values, ref.values)
