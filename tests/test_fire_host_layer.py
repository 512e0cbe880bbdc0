"""Host layer of the fire-weather path (xclim_b200/fire.py) on the oracle-backed device stand-in: argument
handling, unit conversion folded into the kernel parameters, output wrapping, reference errors.  The same
bodies run on the GPU in tests/test_zzz_gpu_fire.py."""
import numpy as np
import pytest

import fake_device
from oracle import fire_oracle as FO
from test_fire_oracle import assert_index_close, mg
from xb_helpers import make_field


def fields(C_shape=(4, 4), T=500, kelvin=False):
    inp = mg.cffwis_inputs(seed=9, C=16, T=T)
    dims = ("time", "lat", "lon")
    lat = np.array([-40.0, -5.0, 20.0, 55.0])

    def f(a, units):
        return make_field(np.ascontiguousarray(a.T).reshape((T,) + C_shape), "1990-01-01", calendar="noleap", units=units,
                          dims=dims)
    tas = inp["tas"] + np.float32(273.15) if kelvin else inp["tas"]
    out = dict(tas=f(tas, "K" if kelvin else "degC"), pr=f(inp["pr"] / np.float32(86400) if kelvin else inp["pr"],
                                                          "kg m-2 s-1" if kelvin else "mm/d"),
               hurs=f(inp["hurs"], "%"), ws=f(inp["ws"] / np.float32(3.6) if kelvin else inp["ws"], "m/s" if kelvin else "km/h"),
               snd=f(inp["snd"], "m"))
    for v in out.values():
        v.coords["lat"] = lat
    latf = make_field(np.repeat(lat[:, None], 4, 1), None, units="degrees_north", dims=("lat", "lon")) \
        if False else None
    return out, lat, latf


def check_bodies(fire, Field):
    """Shared with the GPU test: returns nothing, asserts."""
    f, lat, _ = fields()
    T = f["tas"].shape[0]
    month = np.asarray(f["tas"].time.month)
    lat2d = np.repeat(lat[:, None], 4, 1)
    tc = lambda x: np.asarray(x.values).reshape(T, -1)   # noqa: E731
    latF = Field(lat, ("lat",), None, {}, {"units": "degrees_north"})

    # 1. always-on, all indexes, through fire_weather_ufunc
    out = fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], hurs=f["hurs"], sfcWind=f["ws"], lat=latF)
    assert list(out) == ["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "DSR"]
    out_all = out
    exp = FO.fire_weather_calc(tc(f["tas"]), tc(f["pr"]), tc(f["hurs"]), tc(f["ws"]), None, month, lat2d.reshape(-1), None,
                               *(np.full(16, np.nan, np.float32),) * 3, np.zeros(16, np.float32), outputs=list(out))
    for k in out:
        assert out[k].dims == ("time", "lat", "lon") and out[k].values.dtype == np.float32
        assert_index_close(np.asarray(out[k].values).reshape(T, -1), exp[k], k, rtol=1e-5, atol=1e-6)

    # 2. only what is asked for (+ dependencies), computed season, overwintering, previous codes as Fields
    dc0 = Field(np.linspace(50, 400, 16, dtype=np.float32).reshape(4, 4), ("lat", "lon"), None, {}, {})
    out = fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], snd=f["snd"], lat=lat2d, dc0=dc0, indexes=["DC"],
                                  season_method="LA08", overwintering=True, temp_end_thresh=4)
    assert list(out) == ["DC", "season_mask", "winter_pr"]
    exp = FO.fire_weather_calc(tc(f["tas"]), tc(f["pr"]), None, None, tc(f["snd"]), month, lat2d.reshape(-1), None,
                               np.asarray(dc0.values).reshape(-1), np.full(16, np.nan, np.float32),
                               np.full(16, np.nan, np.float32), np.zeros(16, np.float32), outputs=list(out),
                               season_method="LA08", overwintering=True, temp_end_thresh=4)
    assert out["season_mask"].values.dtype == bool and out["winter_pr"].dims == ("lat", "lon")
    np.testing.assert_array_equal(np.asarray(out["season_mask"].values).reshape(T, -1), exp["season_mask"])
    np.testing.assert_allclose(np.asarray(out["DC"].values).reshape(T, -1), exp["DC"], rtol=2e-6, equal_nan=True)
    np.testing.assert_allclose(np.asarray(out["winter_pr"].values).reshape(-1), exp["winter_pr"], rtol=2e-6, equal_nan=True)

    # 3. a season mask handed in + the CFS dry start == the season computed on the fly
    mask = fire.fire_season(f["tas"], method="WF93")
    assert mask.values.dtype == bool and mask.attrs["units"] == ""
    a = fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], hurs=f["hurs"], lat=latF, season_mask=mask, dry_start="CFS",
                                indexes=["DC", "DMC"], dmc_dry_factor=5)
    b = fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], hurs=f["hurs"], lat=latF, season_method="WF93", dry_start="CFS",
                                indexes=["DC", "DMC"], dmc_dry_factor=5)
    assert list(a) == ["DC", "DMC"] and list(b) == ["DC", "DMC", "season_mask"]
    np.testing.assert_array_equal(np.asarray(b["season_mask"].values), np.asarray(mask.values))
    for k in ("DC", "DMC"):
        np.testing.assert_array_equal(np.asarray(a[k].values), np.asarray(b[k].values))
    assert np.isnan(np.asarray(a["DC"].values)[~np.asarray(mask.values)]).all()      # no code outside the season

    # 4. unit-aware entry points: K, kg m-2 s-1, m/s inputs give what degC, mm/d, km/h inputs give
    g, _, _ = fields(kelvin=True)
    ci = fire.cffwis_indices(g["tas"], g["pr"], g["ws"], g["hurs"], latF, season_method="WF93",
                             temp_start_thresh="285.15 K")
    cj = fire.cffwis_indices(f["tas"], f["pr"], f["ws"], f["hurs"], latF, season_method="WF93")
    assert ci._fields == ("DC", "DMC", "FFMC", "ISI", "BUI", "FWI") and ci.FWI.attrs["units"] == ""
    for k in ci._fields:       # the converted series differ from the native ones by float32 rounding of the conversion
        x, y = np.asarray(getattr(ci, k).values), np.asarray(getattr(cj, k).values)
        assert np.array_equal(np.isnan(x), np.isnan(y)), k
        np.testing.assert_allclose(x, y, rtol=2e-2, atol=2e-2, equal_nan=True, err_msg=k)
    dc = fire.drought_code(g["tas"], g["pr"], latF, snd=g["snd"], season_method="GFWED", dry_start="GFWED",
                           snow_cover_days=30)
    dmc = fire.duff_moisture_code(f["tas"], f["pr"], f["hurs"], latF, dmc0=np.full((4, 4), 20.0, np.float32))
    assert dc.dims == dmc.dims == ("time", "lat", "lon") and dc.attrs["units"] == ""
    exp = FO.fire_weather_calc(tc(f["tas"]), tc(f["pr"]), tc(f["hurs"]), None, None, month, lat2d.reshape(-1), None,
                               np.full(16, np.nan, np.float32), np.full(16, 20.0, np.float32),
                               np.full(16, np.nan, np.float32), np.zeros(16, np.float32), outputs=["DMC"])
    np.testing.assert_allclose(np.asarray(dmc.values).reshape(T, -1), exp["DMC"], rtol=2e-6, equal_nan=True)

    # 4b. the element-wise members on labelled arrays and bare arrays
    isi = fire.initial_spread_index(f["ws"], out_all["FFMC"])
    assert isi.dims == ("time", "lat", "lon") and isi.attrs["units"] == ""
    assert_index_close(np.asarray(isi.values), np.asarray(out_all["ISI"].values), "ISI")
    bui = fire.build_up_index(out_all["DMC"], out_all["DC"])
    assert_index_close(np.asarray(bui.values), np.asarray(out_all["BUI"].values), "BUI")
    fwi = fire.fire_weather_index(np.asarray(isi.values), np.asarray(bui.values))
    assert isinstance(fwi, np.ndarray)
    assert_index_close(fwi, np.asarray(out_all["FWI"].values), "FWI")
    assert_index_close(np.asarray(fire.daily_severity_rating(out_all["FWI"]).values), np.asarray(out_all["DSR"].values), "DSR")
    wdc = fire.overwintering_drought_code(Field(np.array([[300.0, 100.0]], np.float32), ("lat", "lon"), None, {}, {}),
                                          Field(np.array([[0.11, 0.05]], np.float32), ("lat", "lon"), None, {}, {"units": "m"}))
    np.testing.assert_allclose(np.asarray(wdc.values), [[109.4657, 105.176]], rtol=1e-5)

    # 5. the reference's errors
    with pytest.raises(TypeError, match="Missing input argument hurs"):
        fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], lat=latF, indexes=["DMC"])
    with pytest.raises(ValueError, match="overwintering"):
        fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], lat=latF, indexes=["DC"], overwintering=True)
    with pytest.raises(ValueError, match="dry_start"):
        fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], lat=latF, indexes=["DC"], dry_start="wet")
    with pytest.raises(ValueError, match="not a valid parameter"):
        fire.drought_code(f["tas"], f["pr"], latF, bad_param=3)
    with pytest.raises(ValueError, match="Invalid lat"):
        fire.drought_code(f["tas"], f["pr"], np.full((4, 4), 95.0))
    with pytest.raises(ValueError, match="Thresholds must be scalar"):
        fire.fire_season(f["tas"], temp_start_thresh=[1, 2])


def test_fire_host_layer_on_the_oracle_backed_device(monkeypatch):
    from xclim_b200 import Field, fire
    fake_device.install(monkeypatch)
    check_bodies(fire, Field)


def test_fire_weather_streams_in_lat_slabs(monkeypatch, tmp_path):
    import fake_stream
    check_streaming(tmp_path, fake_stream.install(monkeypatch))


def check_streaming(tmp_path, lib=None):
    """File-backed inputs -> slab streamer -> every output of the multi-output call assembled on the host.
    Shared with the GPU test (``lib``: the recording stand-in of the C ABI on the CPU, None on the GPU)."""
    import xclim_b200
    from xclim_b200 import Field, fire, io
    f, lat, _ = fields(T=400)
    lazy = {}
    for k in ("tas", "pr", "hurs", "ws", "snd"):
        f[k].coords["lat"] = lat
        lazy[k] = io.open_field(io.save_zarr(str(tmp_path / f"{k}.zarr"), f[k], name=k, chunks=(200, 2, 4), compressor=None))
    latF = Field(lat, ("lat",), None, {}, {"units": "degrees_north"})
    dc0 = np.linspace(50, 400, 16, dtype=np.float32).reshape(4, 4)
    kw = dict(lat=latF, dc0=dc0, season_method="LA08", overwintering=True)
    ref = fire.fire_weather_ufunc(tas=f["tas"], pr=f["pr"], hurs=f["hurs"], sfcWind=f["ws"], snd=f["snd"], **kw)
    row = 400 * 4 * 4
    with xclim_b200.set_options(stream_min_bytes=0, stream_slab_bytes=row):
        got = fire.fire_weather_ufunc(tas=lazy["tas"], pr=lazy["pr"], hurs=lazy["hurs"], sfcWind=lazy["ws"],
                                      snd=lazy["snd"], **kw)
        ci = fire.cffwis_indices(lazy["tas"], lazy["pr"], lazy["ws"], lazy["hurs"], latF)
    # slabs follow the 2-row chunks of the stores: 2 slabs x 5 inputs, then 2 slabs x 4 inputs (no snd)
    assert lib is None or len([b for b in lib.boxes if b[2] == 1]) == 2 * 5 + 2 * 4
    assert list(got) == list(ref) == ["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "DSR", "season_mask", "winter_pr"]
    for k in ref:
        assert got[k].dims == ref[k].dims and isinstance(got[k].values, np.ndarray), k
        np.testing.assert_array_equal(got[k].values, np.asarray(ref[k].values), err_msg=k)
    np.testing.assert_array_equal(got["DC"].coords["lat"], lat)
    cj = fire.cffwis_indices(f["tas"], f["pr"], f["ws"], f["hurs"], latF)
    assert type(ci).__name__ == "CFFWISIndices"
    for a, b in zip(ci, cj):
        np.testing.assert_array_equal(a.values, np.asarray(b.values))


def test_fire_dataarray_boundary(monkeypatch):
    """DataArray in -> DataArray out for the fire-weather functions (through tests/mini_xarray.py: xarray is
    absent from the image), latitude taken from a DataArray coordinate as `lat=ds.lat` is in the reference."""
    import mini_xarray as mx
    from xclim_b200 import fire
    fake_device.install(monkeypatch)
    mx.install(monkeypatch)
    inp = mg.cffwis_inputs(seed=13, C=15, T=420)
    T = 420

    def da(a, units):
        return mx.daily(np.ascontiguousarray(a.T).reshape(T, 3, 5), "1995-01-01", units=units)
    tas, pr, hurs, ws = da(inp["tas"], "degC"), da(inp["pr"], "mm/d"), da(inp["hurs"], "%"), da(inp["ws"], "km/h")
    lat = tas.coords["lat"]
    out = fire.cffwis_indices(tas, pr, ws, hurs, lat, season_method="WF93")
    assert all(isinstance(o, mx.DataArray) and o.dims == ("time", "lat", "lon") for o in out)
    assert out.FWI.attrs["units"] == "" and out.DC.values.dtype == np.float32
    np.testing.assert_array_equal(out.DC.coords["lon"].values, tas.coords["lon"].values)
    lat_cells = np.repeat(np.asarray(lat.values, dtype=np.float64), 5)
    x = lambda d: np.asarray(d.values).reshape(T, -1)   # noqa: E731
    exp = FO.fire_weather_calc(x(tas), x(pr), x(hurs), x(ws), None, np.asarray(mx_month(tas)), lat_cells, None,
                               *(np.full(15, np.nan, np.float32),) * 3, np.zeros(15, np.float32),
                               outputs=["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "season_mask"], season_method="WF93")
    for k in out._fields:
        assert_index_close(x(getattr(out, k)), exp[k], k)
    mask = fire.fire_season(tas, method="WF93")
    assert isinstance(mask, mx.DataArray) and mask.values.dtype == bool
    np.testing.assert_array_equal(x(mask), exp["season_mask"])
    dc = fire.drought_code(tas, pr, lat, season_mask=mask)
    np.testing.assert_allclose(x(dc), exp["DC"], rtol=2e-6, equal_nan=True)


def mx_month(da):
    from xclim_b200.timeaxis import TimeAxis
    return TimeAxis.from_xarray(da["time"]).month
