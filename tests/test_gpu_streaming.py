"""GPU: the end-to-end slab streamer (xclim_b200/streaming.py) gives the results of the one-piece path."""
import numpy as np
import pytest

from xb_helpers import make_field

pytestmark = pytest.mark.gpu


def _inputs(rng, T=365 * 4, shape=(13, 24)):
    t = np.arange(T)
    tas = (285 + 10 * np.sin(2 * np.pi * (t % 365 - 110) / 365)[:, None, None]
           + 3 * rng.standard_normal((T,) + shape)).astype(np.float32)
    pr = rng.gamma(0.4, 6.0, size=(T,) + shape).astype(np.float32)
    pr[rng.random(pr.shape) < 0.5] = 0
    for a in (tas, pr):
        a[rng.random(a.shape) < 0.002] = np.nan
    return tas, pr


def test_streamed_calls_match_one_piece_calls(cuda):
    import xclim_b200
    from xclim_b200 import atmos, calendar as xcal, indices, streaming
    rng = np.random.default_rng(71)
    tas, pr = _inputs(rng)
    dims = ("time", "lat", "lon")
    f_tas = make_field(tas, "1981-01-01", calendar="noleap", units="K", dims=dims)
    f_pr = make_field(pr, "1981-01-01", calendar="noleap", units="mm/d", dims=dims)
    # reference results: one-piece unwrap (inputs far below the streaming threshold)
    ref = {
        "cdd": atmos.maximum_consecutive_dry_days(f_pr).values,
        "cdd_idx": indices.maximum_consecutive_dry_days(f_pr).values,
        "tg": atmos.tg_mean(f_tas, freq="MS").values,
        "wet": indices.wetdays(f_pr).values,
        "dsf": indices.dry_spell_frequency(f_pr).values,
    }
    per_ref = xcal.percentile_doy(f_tas, window=5, per=90.0)
    ref["tx90p"] = indices.tx90p(f_tas, xcal.select_percentile(per_ref, 90.0)).values
    row_bytes = tas.shape[0] * tas.shape[2] * 4
    with xclim_b200.set_options(stream_min_bytes=0, stream_slab_bytes=3 * row_bytes):   # 13 rows -> 5 slabs
        assert len(streaming.plan_slabs(13, row_bytes, 3 * row_bytes)) == 5
        got = {
            "cdd": atmos.maximum_consecutive_dry_days(f_pr),
            "cdd_idx": indices.maximum_consecutive_dry_days(f_pr),
            "tg": atmos.tg_mean(f_tas, freq="MS"),
            "wet": indices.wetdays(f_pr),
            "dsf": indices.dry_spell_frequency(f_pr),
        }
        per = xcal.percentile_doy(f_tas, window=5, per=90.0)
        got["tx90p"] = indices.tx90p(f_tas, xcal.select_percentile(per, 90.0))
    np.testing.assert_array_equal(per.values, per_ref.values)
    assert per.dims == per_ref.dims and per.attrs["climatology_bounds"] == per_ref.attrs["climatology_bounds"]
    for k, v in got.items():
        assert isinstance(v.values, np.ndarray), k
        assert v.values.dtype == ref[k].dtype and v.dims[0] == "time", k
        np.testing.assert_array_equal(v.values, ref[k], err_msg=k)
    assert got["cdd"].attrs["units"] == "days" and got["wet"].attrs["units"] == "d"
    assert not xclim_b200.options.OPTIONS["device_outputs"] and not xclim_b200.options.OPTIONS["_in_stream"]


def test_streaming_is_skipped_for_device_and_small_inputs(cuda):
    import torch
    from xclim_b200 import Field, indices, streaming
    rng = np.random.default_rng(72)
    tas, pr = _inputs(rng, T=365 * 2, shape=(3, 8))
    f = make_field(pr, "1981-01-01", calendar="noleap", units="mm/d", dims=("time", "lat", "lon"))
    series, tables = streaming._classify((f,), {})
    assert streaming._streamable(series, tables) is None            # below stream_min_bytes
    d = Field(torch.from_numpy(pr).cuda(), f.dims, f.time, {}, dict(f.attrs))
    out = indices.wetdays(d)
    np.testing.assert_array_equal(out.values, indices.wetdays(f).values)


def test_plan_slabs_covers_every_row():
    from xclim_b200 import streaming
    for n in (1, 7, 90, 721):
        for slab in (1, 10 ** 6, 10 ** 12):
            pl = streaming.plan_slabs(n, 10 ** 5, slab)
            assert pl[0][0] == 0 and pl[-1][1] == n and all(a[1] == b[0] for a, b in zip(pl, pl[1:]))
            assert all(b > a for a, b in pl)


def test_memory_mapped_file_in_file_out(cuda, tmp_path):
    """.npy on disk -> memory map -> reader thread -> pinned slabs -> kernels -> .npy on disk;
    the result equals the in-memory call."""
    import xclim_b200
    from xclim_b200 import atmos, indices, io
    rng = np.random.default_rng(73)
    tas, pr = _inputs(rng, T=365 * 3, shape=(11, 16))
    f_pr = make_field(pr, "1981-01-01", calendar="noleap", units="mm/d", dims=("time", "lat", "lon"))
    path = str(tmp_path / "pr.npy")
    io.save_npy(path, f_pr)
    lazy = io.open_npy(path)
    assert isinstance(lazy.values, np.memmap) and lazy.dims == ("time", "lat", "lon") and lazy.attrs["units"] == "mm/d"
    assert lazy.time.calendar == "noleap" and len(lazy.time) == pr.shape[0]
    ref = atmos.maximum_consecutive_dry_days(f_pr).values
    row = pr.shape[0] * pr.shape[2] * 4
    with xclim_b200.set_options(stream_min_bytes=0, stream_slab_bytes=2 * row):     # 11 rows -> 6 slabs
        out = atmos.maximum_consecutive_dry_days(lazy)
        wet = indices.wetdays(lazy)
    np.testing.assert_array_equal(out.values, ref)
    np.testing.assert_array_equal(wet.values, indices.wetdays(f_pr).values)
    out_path = str(tmp_path / "cdd.npy")
    io.save_npy(out_path, out)
    back = io.open_npy(out_path, mmap=False)
    np.testing.assert_array_equal(back.values, ref)
    assert back.attrs["units"] == "days" and back.dims == ("time", "lat", "lon")
