"""The DataArray-in / DataArray-out boundary (north_star: "keeping the xclim.indices.* / Indicator.compute
xarray-DataArray-in/out signatures"): the xarray branches of the host layer executed through
tests/mini_xarray.py (xarray is absent from the image; tests/test_xarray_real.py repeats this with the real
package when importable).  CPU: oracle-backed device functions; GPU: the real kernels."""
import numpy as np
import pytest

import fake_device
import mini_xarray as mx
from oracle import xclim_oracle as O


def _data():
    rng = np.random.default_rng(91)
    T, shape = 365 * 3, (3, 5)
    t = np.arange(T)
    tas = (285 + 10 * np.sin(2 * np.pi * (t % 365 - 110) / 365)[:, None, None]
           + 3 * rng.standard_normal((T,) + shape)).astype(np.float32)
    pr = rng.gamma(0.4, 6.0, size=(T,) + shape).astype(np.float32)
    pr[rng.random(pr.shape) < 0.5] = 0
    pr[400:403, 0, 0] = np.nan
    return tas, pr


def _scenarios():
    from xclim_b200 import TimeAxis, atmos, calendar as xcal, indices
    tas, pr = _data()
    da_t = mx.daily(tas, "2001-01-01", units="K", standard_name="air_temperature")
    da_p = mx.daily(pr, "2001-01-01", units="mm/d")
    ta = TimeAxis.daily("2001-01-01", tas.shape[0], "noleap")
    poff = ta.period_offsets("YS")
    # (1) the headline index + indicator
    cdd = indices.maximum_consecutive_dry_days(da_p, thresh="1 mm/day", freq="YS")
    assert isinstance(cdd, mx.DataArray) and cdd.dims == ("time", "lat", "lon")
    np.testing.assert_array_equal(cdd.values, O.maximum_consecutive_dry_days(pr, 1.0, poff))
    assert cdd.attrs["units"] == "d" and set(cdd.coords) >= {"time", "lat", "lon"}
    np.testing.assert_array_equal(cdd.coords["lon"].values, da_p.coords["lon"].values)
    assert len(cdd.coords["time"].values) == 3 and str(cdd.coords["time"].values[1]).startswith("2002-01-01")
    ind = atmos.maximum_consecutive_dry_days(da_p, freq="YS")
    assert isinstance(ind, mx.DataArray) and np.isnan(ind.values[1, 0, 0]) and ind.attrs["units"] == "days"
    # (2) monthly mean (config 0 of BASELINE.json: tg_mean freq=MS)
    tg = atmos.tg_mean(da_t, freq="MS")
    assert tg.values.shape == (36, 3, 5) and tg.attrs["units"] == "K"
    np.testing.assert_allclose(tg.values, O.select_resample_op(tas.astype(np.float64), "mean", ta.period_offsets("MS")),
                               rtol=1e-5)
    # (3) percentile_doy -> .sel(percentiles=) -> tx90p
    per = xcal.percentile_doy(da_t, window=5, per=[10.0, 90.0])
    assert isinstance(per, mx.DataArray) and per.dims == ("lat", "lon", "dayofyear", "percentiles")
    assert per.values.shape == (3, 5, 365, 2) and per.attrs["window"] == 5 and "climatology_bounds" in per.attrs
    tab = O.percentile_doy(tas, ta.year, ta.doy, 5, [10.0, 90.0])
    np.testing.assert_array_equal(np.moveaxis(per.values, (2, 3), (0, 1)), tab)
    p90 = per.sel(percentiles=90.0)
    tx = indices.tx90p(da_t, p90, freq="YS")
    np.testing.assert_array_equal(tx.values, O.doy_threshold_count(tas, tab[:, 1], ta.doy, poff, ">"))
    assert tx.values.dtype == np.int64 and tx.attrs["units"] == "d"


def test_dataarray_boundary_host_layer(monkeypatch):
    fake_device.install(monkeypatch)
    mx.install(monkeypatch)
    _scenarios()


@pytest.mark.gpu
def test_dataarray_boundary_on_device(cuda, monkeypatch):
    mx.install(monkeypatch)
    _scenarios()


@pytest.mark.gpu
def test_dataarray_inputs_stream_in_slabs(cuda, monkeypatch):
    """DataArray inputs above the streaming threshold go through the slab streamer and come back as
    DataArrays with the template's coordinates."""
    import xclim_b200
    from xclim_b200 import TimeAxis, indices
    mx.install(monkeypatch)
    tas, pr = _data()
    da_p = mx.daily(pr, "2001-01-01", units="mm/d")
    ta = TimeAxis.daily("2001-01-01", pr.shape[0], "noleap")
    row = pr.shape[0] * pr.shape[2] * 4
    with xclim_b200.set_options(stream_min_bytes=0, stream_slab_bytes=row):
        out = indices.maximum_consecutive_dry_days(da_p, thresh="1 mm/day", freq="YS")
    assert isinstance(out, mx.DataArray) and out.dims == ("time", "lat", "lon")
    np.testing.assert_array_equal(out.values, O.maximum_consecutive_dry_days(pr, 1.0, ta.period_offsets("YS")))
    np.testing.assert_array_equal(out.coords["lat"].values, da_p.coords["lat"].values)


def _indexer_scenario():
    """Indicator-level `select_time` indexers on DataArray inputs (ResamplingIndicatorWithIndexing,
    core/indicator.py:1611-1673): DataArray in -> DataArray out, same numbers as the Field path."""
    from xclim_b200 import atmos
    from xb_helpers import make_field
    tas, pr = _data()
    da_t = mx.daily(tas - 273.15 + 4, "2001-01-01", units="degC")
    f_t = make_field(tas - 273.15 + 4, "2001-01-01", calendar="noleap", units="degC", dims=("time", "lat", "lon"))
    for kw in (dict(date_bounds=("09-01", "12-31")), dict(month=[1, 2, 12]), dict(season="JJA")):
        got = atmos.frost_days(da_t, freq="YS", **kw)
        ref = atmos.frost_days(f_t, freq="YS", **kw)
        assert isinstance(got, mx.DataArray) and got.dims == ("time", "lat", "lon"), kw
        np.testing.assert_array_equal(np.asarray(got.values), np.asarray(ref.values), err_msg=str(kw))
        assert got.attrs["units"] == ref.attrs["units"]
        np.testing.assert_array_equal(got.coords["lon"].values, da_t.coords["lon"].values)
    da_p = mx.daily(pr, "2001-01-01", units="mm/d")
    f_p = make_field(pr, "2001-01-01", calendar="noleap", units="mm/d", dims=("time", "lat", "lon"))
    got = atmos.maximum_consecutive_dry_days(da_p, freq="YS", season="DJF")
    ref = atmos.maximum_consecutive_dry_days(f_p, freq="YS", season="DJF")
    assert isinstance(got, mx.DataArray)
    np.testing.assert_array_equal(np.asarray(got.values), np.asarray(ref.values))


def test_indicator_indexers_on_dataarrays(monkeypatch):
    fake_device.install(monkeypatch)
    mx.install(monkeypatch)
    _indexer_scenario()
