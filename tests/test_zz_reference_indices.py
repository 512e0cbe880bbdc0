"""Known answers of the reference's tests/test_indices.py (VALUES only; `*_series` fixtures become Fields
starting 2000-07-01 like xclim.testing.helpers.test_timeseries), asserted through the index entry points
on both backends: oracle stand-ins on the CPU (tests/fake_device.py), real kernels under -m gpu.
Precipitation inputs are given in mm/d (the reference converts kg m-2 s-1 rates to daily amounts)."""
import numpy as np
import pytest

import fake_device
from xb_helpers import make_field

K2C = 273.15


@pytest.fixture(params=["oracle-on-cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    if request.param == "cuda":
        import torch
        if not torch.cuda.is_available():
            pytest.skip("no CUDA device")
    else:
        fake_device.install(monkeypatch)
    return request.param


def series(values, units="K", start="2000-07-01", calendar="standard"):
    return make_field(np.asarray(values, dtype=np.float32), start, calendar=calendar, units=units)


def test_max_n_day_precipitation_amount(backend):          # :42-64
    from xclim_b200 import indices
    a = series([3, 4, 20, 20, 0, 6, 9, 25, 0, 0], "mm/d")
    assert indices.max_n_day_precipitation_amount(a, 2).values[0] == 40
    assert indices.max_n_day_precipitation_amount(a, 10).values[0] == 87
    b = series([3, 4, 20, 20, 0, 6, 15, 25, 0, 0], "mm/d")
    out = indices.max_n_day_precipitation_amount(b, 2)
    assert out.values[0] == 40 and len(out.values) == 1


def test_max_1day_precipitation_amount(backend):           # :66-103
    from xclim_b200 import indices
    for v in ([3, 4, 20, 0, 0], [20, 4, 20, 20, 0], [20, 20, 20, 20, 20]):
        out = indices.max_1day_precipitation_amount(series(v, "mm/day"))
        assert out.values[0] == 20 and len(out.values) == 1


def _cold(start="2000-07-01", second=True):
    a = np.zeros(365)
    a[10:20] -= 15
    a[40:43] -= 50
    if second:
        a[80:86] -= 30
        a[95:101] -= 30
    else:
        a[80:100] -= 30
    return series(a + K2C, "K", start)


def test_cold_spell_days(backend):                          # :119-129
    from xclim_b200 import indices
    out = indices.cold_spell_days(_cold(second=False), thresh="-10. degC", freq="MS")
    np.testing.assert_array_equal(out.values, [10, 0, 12, 8, 0, 0, 0, 0, 0, 0, 0, 0])
    assert out.attrs["units"] == "d"


def test_cold_spell_frequency_max_total(backend):           # :132-183
    from xclim_b200 import indices
    da = _cold("1971-01-01")
    np.testing.assert_array_equal(indices.cold_spell_frequency(da, thresh="-10. degC", freq="MS").values,
                                  [1, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(indices.cold_spell_frequency(da, thresh="-10. degC", freq="YS").values, [3])
    np.testing.assert_array_equal(indices.cold_spell_max_length(da, thresh="-10. degC", freq="MS").values,
                                  [10, 3, 6, 6, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(indices.cold_spell_max_length(da, thresh="-10. degC", freq="YS").values, [10])
    np.testing.assert_array_equal(indices.cold_spell_total_length(da, thresh="-10. degC", freq="MS").values,
                                  [10, 3, 6, 6, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(indices.cold_spell_total_length(da, thresh="-10. degC", freq="YS").values, [25])


def test_maximum_consecutive_frost_days(backend):           # :186-201
    from xclim_b200 import indices
    f = indices.maximum_consecutive_frost_days
    assert f(series(np.array([3, 4, 5, -1, 3]) + K2C)).values[0] == 1
    assert f(series(np.array([3, 4, 5, 1, 3]) + K2C)).values[0] == 0
    assert f(series(np.zeros(365) - 10 + K2C)).values[0] == 365        # default freq YS-JUL: one period


def test_maximum_consecutive_frost_free_days(backend):      # :204-229
    from xclim_b200 import indices
    f = indices.maximum_consecutive_frost_free_days
    assert f(series(np.array([3, 4, 5, -1, 3]) + K2C)).values[0] == 3
    assert f(series(np.array([3, 4, 5, -0.8, -2, 3]) + K2C), thresh="-1 degC").values[0] == 4
    assert f(series(np.array([3, 4, 5, 1, 3]) + K2C)).values[0] == 5
    assert (f(series(np.zeros(365) - 10 + K2C)).values == 0).all()
    assert f(series(np.array([-1, -1, 1, 1, 0, 2, -1]) + K2C)).values[0] == 4


def test_degree_days(backend):                              # :232-246, 1617-1622
    from xclim_b200 import indices
    assert indices.cooling_degree_days(series(np.array([10, 15, -5, 18]) + K2C)).values[0] == 0
    np.testing.assert_allclose(indices.cooling_degree_days(series(np.array([20, 25, -15, 19]) + K2C)).values[0], 10,
                               rtol=1e-5)
    a = np.zeros(365)
    a[0] = 5
    np.testing.assert_allclose(indices.growing_degree_days(series(a + K2C)).values[0], 1, rtol=1e-4)


def test_daily_pr_intensity(backend):                       # :1442-1454
    from xclim_b200 import indices
    pr = np.zeros(365)
    pr[3:8] += [0.5, 1, 2, 3, 4]
    np.testing.assert_allclose(indices.daily_pr_intensity(series(pr, "mm/d"), thresh="1 mm/day").values[0], 2.5)


def test_frost_dates(backend):                              # :1474-1572
    from xclim_b200 import indices
    a = np.zeros(365)
    a[180:270] = 303.15
    tas = series(a, start="2000-01-01")
    lsf = indices.last_spring_frost(tas)
    assert lsf.values[0] == 180 and lsf.attrs["is_dayofyear"] == 1 and lsf.attrs["is_dayofyear"].dtype == np.int32
    assert indices.first_day_temperature_below(tas).values[0] == 271
    assert np.isnan(indices.first_day_temperature_below(series(np.zeros(365) + 303.15, start="2000-01-01")).values[0])
    with pytest.raises(ValueError):
        indices.first_day_temperature_below(tas, op=">=")
    b = np.zeros(365) + 307
    b[180:270] = 270
    tas = series(b, start="2000-01-01")
    assert indices.first_day_temperature_above(tas).values[0] == 1
    assert indices.first_day_temperature_above(tas, after_date="07-01").values[0] == 271
    assert np.isnan(indices.first_day_temperature_above(series(np.zeros(365) + 270, start="2000-01-01")).values[0])
    with pytest.raises(ValueError):
        indices.first_day_temperature_above(tas, op="<")
    tg = np.zeros(365) - 1
    w = 5
    tg[10:10 + w - 1] += 6      # too short
    tg[20:20 + w] += 1          # does not cross the threshold
    tg[30:30 + w] += 6          # ok
    tg[40:40 + w + 1] += 6      # second valid run, ignored
    out = indices.first_day_temperature_above(series(tg + K2C, start="2000-01-01"), thresh="0 degC", window=w)
    assert out.values[0] == 31
    out = indices.first_day_temperature_above(series(np.zeros(365) - 1 + K2C, start="2000-01-01"), thresh="0 degC", window=5)
    assert np.isnan(out.values[0])


def _dated(base, inside, d1, d2, n=365, start="2000-01-01"):
    """`base` everywhere except `inside` between the two dates (inclusive), like tas.where(~isin(slice(d1, d2)))."""
    import pandas as pd
    idx = pd.date_range(start, periods=n, freq="D")
    v = np.full(n, base, np.float32)
    v[(idx >= d1) & (idx <= d2)] = inside
    return series(v, start=start)


@pytest.mark.parametrize("d1,d2,expected", [("1950-01-01", "1951-01-01", 0), ("2000-01-01", "2000-12-31", 365),
                                            ("2000-06-15", "2001-01-01", 199), ("2000-06-15", "2000-07-15", 31)])
def test_frost_season_length(backend, d1, d2, expected):    # :1710-1727
    from xclim_b200 import indices
    out = indices.frost_season_length(_dated(300, 270, d1, d2), freq="YS", mid_date="07-01")
    np.testing.assert_array_equal(out.values, [expected])


def test_frost_season_length_north_hemisphere(backend):     # :1729-1734
    from xclim_b200 import indices
    tas = _dated(300, 270, "2000-11-01", "2001-03-01", n=730)
    out = indices.frost_season_length(tas)
    labels = list(out.coords["time"])
    assert out.values[[str(t)[:10] for t in labels].index("2000-07-01")] == 121


def test_frost_free_season_start(backend):                  # :1738-1764
    from xclim_b200 import indices
    tn = np.zeros(365) - 1
    w = 5
    tn[10:10 + w - 1] += 2
    tn[20:20 + w] += 1
    tn[30:30 + w + 1] += 1
    out = indices.frost_free_season_start(series(tn + K2C, start="2000-01-01"), window=w)
    assert out.values[0] == 21 and out.attrs["is_dayofyear"] == 1
    out = indices.frost_free_season_start(series(np.zeros(365) - 1, start="2000-01-01"))
    assert np.isnan(out.values[0])


@pytest.mark.parametrize("d1,d2,mid_date,expected", [
    ("1950-01-01", "1951-01-01", "07-01", np.nan), ("2000-01-06", "2000-12-31", "07-01", 365),
    ("2000-07-10", "2001-01-01", "07-01", np.nan), ("2000-06-15", "2000-07-15", "07-01", 198),
    ("2000-06-15", "2000-07-25", "07-15", 208), ("2000-06-15", "2000-07-15", "10-01", 275),
    ("2000-06-15", "2000-07-15", "01-10", np.nan), ("2000-06-15", "2000-07-15", "06-15", np.nan)])
def test_frost_free_season_end(backend, d1, d2, mid_date, expected):   # :1768-1794
    from xclim_b200 import indices
    out = indices.frost_free_season_end(_dated(0, 0.1 + K2C, d1, d2), mid_date=mid_date)
    np.testing.assert_array_equal(out.values, [expected])
    assert out.attrs["is_dayofyear"] == 1


@pytest.mark.parametrize("d1,d2,expected", [("1950-01-01", "1951-01-01", 0), ("2000-01-01", "2000-12-31", 365),
                                            ("2000-06-15", "2001-01-01", 199), ("2000-06-15", "2000-07-15", 31)])
def test_frost_free_season_length(backend, d1, d2, expected):          # :1797-1814
    from xclim_b200 import indices
    out = indices.frost_free_season_length(_dated(270, 300, d1, d2), freq="YS", mid_date="07-01")
    np.testing.assert_array_equal(out.values, [expected])


def test_frost_free_season_length_south_hemisphere(backend):           # :1816-1822
    from xclim_b200 import indices
    tn = _dated(270, 300, "2000-11-01", "2001-03-01", n=730)
    out = indices.frost_free_season_length(tn, freq="YS-JUL", mid_date="01-01")
    labels = [str(t)[:10] for t in out.coords["time"]]
    assert out.values[labels.index("2000-07-01")] == 121


def test_frost_free_spell_max_length_and_hdd(backend):                 # :1825-1846
    from xclim_b200 import indices
    tn = np.zeros(365) - 1
    tn[10:12] = 1
    tn[20:30] = 1
    assert indices.frost_free_spell_max_length(series(tn + K2C, start="2000-01-01")).values[0] == 10
    a = np.zeros(365) + 17
    a[:7] += [-3, -2, -1, 0, 1, 2, 3]
    out = indices.heating_degree_days(series(a + K2C))
    np.testing.assert_allclose(out.values[:1], 6, rtol=1e-4)
    np.testing.assert_allclose(out.values[1:], 0, atol=1e-3)


TN_HW = np.asarray([20, 23, 23, 23, 23, 22, 23, 23, 23, 23]) + K2C
TX_HW = np.asarray([29, 31, 31, 31, 29, 31, 31, 31, 31, 31]) + K2C


@pytest.mark.parametrize("ttn,ttx,window,expected", [("22 C", "30 C", 3, 2), ("22 C", "30 C", 4, 1),
                                                     ("10 C", "10 C", 3, 1), ("40 C", "40 C", 3, 0)])
def test_heat_wave_frequency(backend, ttn, ttx, window, expected):      # :1859-1888
    from xclim_b200 import indices
    out = indices.heat_wave_frequency(series(TN_HW), series(TX_HW), thresh_tasmin=ttn, thresh_tasmax=ttx, window=window)
    np.testing.assert_allclose(out.values, expected)


@pytest.mark.parametrize("ttn,ttx,window,expected", [("22 C", "30 C", 3, 4), ("10 C", "10 C", 3, 10),
                                                     ("40 C", "40 C", 3, 0), ("22 C", "30 C", 5, 0)])
def test_heat_wave_max_length(backend, ttn, ttx, window, expected):     # :1891-1920
    from xclim_b200 import indices
    out = indices.heat_wave_max_length(series(TN_HW), series(TX_HW), thresh_tasmin=ttn, thresh_tasmax=ttx, window=window)
    np.testing.assert_allclose(out.values, expected)


@pytest.mark.parametrize("ttn,ttx,window,expected", [("22 C", "30 C", 3, 7), ("10 C", "10 C", 3, 10),
                                                     ("40 C", "40 C", 3, 0), ("22 C", "30 C", 5, 0)])
def test_heat_wave_total_length(backend, ttn, ttx, window, expected):   # :1923-1953
    from xclim_b200 import indices
    out = indices.heat_wave_total_length(series(TN_HW), series(TX_HW), thresh_tasmin=ttn, thresh_tasmax=ttx,
                                         window=window)
    np.testing.assert_allclose(out.values, expected)


def test_hot_days(backend):                                             # :2029-2037
    from xclim_b200 import indices
    a = np.zeros(365)
    a[:6] += [27, 28, 29, 30, 31, 32]
    out = indices.hot_days(series(a + K2C), thresh="30 C")
    np.testing.assert_array_equal(out.values[:1], [2])
    np.testing.assert_array_equal(out.values[1:], [0])


@pytest.mark.parametrize("thresh,window,op,expected", [("30 C", 3, ">", 2), ("30 C", 4, ">", 1), ("29 C", 3, ">", 2),
                                                       ("29 C", 3, ">=", 1), ("10 C", 3, ">", 1), ("40 C", 5, ">", 0)])
def test_hot_spell_frequency(backend, thresh, window, op, expected):    # :2040-2056
    from xclim_b200 import indices
    out = indices.hot_spell_frequency(series(TX_HW), thresh=thresh, window=window, op=op)
    np.testing.assert_allclose(out.values, expected)


@pytest.mark.parametrize("before,expected", [(True, 1), (False, 0)])
def test_hot_spell_frequency_resampling_order(backend, before, expected):   # :2058-2071
    from xclim_b200 import indices
    a = np.zeros(365)
    a[5:35] = 31
    out = indices.hot_spell_frequency(series(a + K2C), resample_before_rl=before, freq="MS")
    assert out.values[1] == expected


TX_HS = np.asarray([28, 31, 31, 31, 29, 31, 31, 31, 31, 31]) + K2C


@pytest.mark.parametrize("thresh,window,op,expected", [("30 C", 3, ">", 5), ("10 C", 3, ">", 10), ("29 C", 3, ">", 5),
                                                       ("29 C", 3, ">=", 9), ("40 C", 3, ">", 0), ("30 C", 5, ">", 5)])
def test_hot_spell_max_length(backend, thresh, window, op, expected):   # :2085-2101
    from xclim_b200 import indices
    out = indices.hot_spell_max_length(series(TX_HS), thresh=thresh, window=window, op=op)
    np.testing.assert_allclose(out.values, expected)


@pytest.mark.parametrize("thresh,window,op,expected", [("30 C", 3, ">", 8), ("10 C", 3, ">", 10), ("29 C", 3, ">", 8),
                                                       ("29 C", 3, ">=", 9), ("40 C", 3, ">", 0), ("30 C", 5, ">", 5)])
def test_hot_spell_total_length(backend, thresh, window, op, expected):  # :2104-2120
    from xclim_b200 import indices
    out = indices.hot_spell_total_length(series(TX_HS), thresh=thresh, window=window, op=op)
    np.testing.assert_allclose(out.values, expected)


def test_hot_spell_total_length_and_magnitude_monthly(backend):          # :2122-2142
    from xclim_b200 import indices
    a = np.zeros(365)
    a[10:20] += 30
    a[40:43] += 50
    a[80:100] += 30
    out = indices.hot_spell_total_length(series(a + K2C), window=5, thresh="25 C", freq="MS")
    np.testing.assert_array_equal(out.values, [10, 0, 12, 8, 0, 0, 0, 0, 0, 0, 0, 0])
    a = np.zeros(365)
    a[15:20] += 30
    a[40:42] += 50
    a[86:96] += 30
    out = indices.hot_spell_max_magnitude(series(a + K2C), thresh="25 C", freq="MS")
    np.testing.assert_allclose(out.values, [25, 0, 30, 20, 0, 0, 0, 0, 0, 0, 0, 0], atol=1e-3)


def _ramp(sign=1):
    a = np.zeros(365)
    a[:6] += sign * np.array([27, 28, 29, 30, 31, 32])
    return series(a + K2C)


def test_tn_tg_tx_days(backend):                                        # :2145-2281
    from xclim_b200 import indices
    for above, below in ((indices.tn_days_above, indices.tn_days_below), (indices.tg_days_above, indices.tg_days_below),
                         (indices.tx_days_above, indices.tx_days_below)):
        out = above(_ramp(), thresh="30 C")
        np.testing.assert_array_equal(out.values, [2, 0])
        np.testing.assert_array_equal(below(_ramp(-1), thresh="-10 C").values, [6, 0])
        np.testing.assert_array_equal(below(_ramp(-1), thresh="-30 C").values, [2, 0])
    with pytest.warns(UserWarning, match="renamed"):
        np.testing.assert_array_equal(indices.tn_days_above(_ramp(), thresh="30 C", op="gteq").values, [3, 0])
    np.testing.assert_array_equal(indices.tn_days_above(_ramp(), thresh="29 C", op=">=").values, [4, 0])
    np.testing.assert_array_equal(indices.tn_days_above(_ramp(), thresh="28 C", op=">=").values, [5, 0])
    np.testing.assert_array_equal(indices.tn_days_below(_ramp(-1), thresh="-31 C", op="<=").values, [2, 0])
    np.testing.assert_array_equal(indices.tn_days_below(_ramp(-1), thresh="-28 C", op="<=").values, [5, 0])
    with pytest.warns(UserWarning, match="renamed"):
        np.testing.assert_array_equal(indices.tn_days_below(_ramp(-1), thresh="-30 C", op="lteq").values, [3, 0])
    for bad in ("<=", "lt"):
        with pytest.raises(ValueError):
            indices.tn_days_above(_ramp(), thresh="30 C", op=bad)
    for bad in (">=", "gt"):
        with pytest.raises(ValueError):
            indices.tn_days_below(_ramp(-1), thresh="30 C", op=bad)


def test_maximum_consecutive_tx_days(backend):                          # :2384-2391
    from xclim_b200 import indices
    a = np.zeros(365) + 273.15
    a[5:15] += 30
    out = indices.maximum_consecutive_tx_days(series(a, start="2010-01-01"), thresh="25 C", freq="MS")
    assert out.values[0] == 10
    np.testing.assert_array_equal(out.values[1:], 0)


def test_precip_accumulation(backend):                                  # :2394-2420
    import calendar
    from xclim_b200 import indices
    pr = np.zeros(100)
    pr[5:10] = 1
    assert indices.precip_accumulation(series(pr, "mm/d"), freq="MS").values[0] == 5
    import pandas as pd
    t = pd.date_range("2000-01-01", "2010-12-31", freq="D")
    out = indices.precip_accumulation(series(t.year.values, "mm d-1", start="2000-01-01"))
    np.testing.assert_allclose(out.values, [(365 + calendar.isleap(y)) * y for y in range(2000, 2011)])


def test_tx_statistics(backend):                                        # :2640-2666
    from xclim_b200 import indices
    assert indices.tx_min(series([20, 25, -15, 19]), freq="YS").values[0] == -15
    assert indices.tx_max(series([20, 25, -15, 19]), freq="YS").values[0] == 25
    out = indices.tx_mean(series([320, 321, 322, 323, 324]), freq="YS")
    assert out.values[0] == 322 and out.attrs["units"] == "K"
    out = indices.tx_mean(series([20, 21, 22, 23, 24], units="°C"), freq="YS")
    assert out.values[0] == 22 and out.attrs["units"] == "°C"


def test_warm_day_and_night_frequency(backend):                         # :3088-3115
    from xclim_b200 import indices
    for fn, hot in ((indices.warm_day_frequency, 31), (indices.warm_night_frequency, 23)):
        a = np.zeros(35)
        a[25:] = hot
        da = series(a + K2C)
        np.testing.assert_allclose(fn(da, freq="MS").values, [6, 4])
        np.testing.assert_allclose(fn(da, freq="YS").values, [10])
        np.testing.assert_allclose(fn(da, thresh="-1 C").values, [35])
        np.testing.assert_allclose(fn(da, thresh="50 C").values, [0])


def test_wind_indices(backend):                                         # :3118-3136
    from xclim_b200 import indices
    a = np.full(365, 20.0)
    a[10:20] = 2
    a[40:50] = 3.1
    out = indices.calm_days(series(a, "km h-1"), thresh="3 km h-1", freq="MS")
    np.testing.assert_array_equal(out.values, [10, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0])
    assert out.attrs["units"] == "d"
    a = np.zeros(365)
    a[10:20] = 10.8
    a[40:50] = 12
    a[80:90] = 15
    out = indices.windy_days(series(a, "km h-1"), thresh="12 km h-1", freq="MS")
    np.testing.assert_array_equal(out.values, [0, 10, 10, 0, 0, 0, 0, 0, 0, 0, 0, 0])


def test_tx_tn_days_above(backend):                                     # :3139-3161
    from xclim_b200 import indices
    tn = series(np.asarray([20, 23, 23, 23, 23, 22, 23, 23, 23, 23]) + K2C)
    tx = series(np.asarray([29, 31, 31, 31, 29, 31, 30, 31, 31, 31]) + K2C)
    np.testing.assert_allclose(indices.tx_tn_days_above(tn, tx).values, [6])
    np.testing.assert_allclose(indices.tx_tn_days_above(tn, tx, thresh_tasmax="50 C").values, [0])
    np.testing.assert_allclose(indices.tx_tn_days_above(tn, tx, thresh_tasmax="0 C", thresh_tasmin="0 C").values, [10])
    np.testing.assert_allclose(indices.tx_tn_days_above(tn, tx, op=">=").values, [8])
    with pytest.raises(ValueError):
        indices.tx_tn_days_above(tn, tx, op="<")


def _wet():
    a = np.zeros(365)
    a[:7] += [4, 5.5, 6, 6, 2, 7, 5]
    a[100:106] += [1, 6, 7, 5, 2, 1]
    return series(a, "mm/day")


def test_wetdays_and_proportion(backend):                               # :4211-4240
    from xclim_b200 import indices
    np.testing.assert_allclose(indices.wetdays(_wet(), thresh="5 mm/day", freq="MS").values,
                               [5, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_allclose(indices.wetdays(_wet(), thresh="5 mm/day", freq="MS", op=">").values,
                               [4, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_allclose(indices.wetdays_prop(_wet(), thresh="5 mm/day", freq="MS").values,
                               [5 / 31, 0, 0, 3 / 31, 0, 0, 0, 0, 0, 0, 0, 0])
    np.testing.assert_allclose(indices.wetdays_prop(_wet(), thresh="5 mm/day", freq="MS", op=">").values,
                               [4 / 31, 0, 0, 2 / 31, 0, 0, 0, 0, 0, 0, 0, 0])


def test_surface_wind_statistics(backend):                              # :4443-4482
    from xclim_b200 import indices
    w = series([14.11, 15.27, 10.70], "m s-1")
    np.testing.assert_allclose(indices.sfcWind_max(w).values, [15.27], rtol=1e-6)
    np.testing.assert_allclose(indices.sfcWind_mean(w).values, [13.36], rtol=1e-6)
    np.testing.assert_allclose(indices.sfcWind_min(w).values, [10.70], rtol=1e-6)
    np.testing.assert_allclose(indices.sfcWindmax_max(w).values, [15.27], rtol=1e-6)
    np.testing.assert_allclose(indices.sfcWindmax_mean(w).values, [13.36], rtol=1e-6)
    np.testing.assert_allclose(indices.sfcWindmax_min(w).values, [10.70], rtol=1e-6)


def test_atmos_consecutive_frost_days(backend):          # tests/test_temperature.py:291-340
    """Indicator level (index + MissingAny): the answers of atmos.consecutive_frost_days."""
    from xclim_b200 import atmos
    f = atmos.maximum_consecutive_frost_days

    def ts(mod, units="K", base=K2C + 5.0):
        a = np.zeros(365) + base
        mod(a)
        return series(a, units)

    def one(a): a[2] -= 20
    def three(a): a[2:5] -= 20
    def two_equal(a): a[2:5] -= 20; a[6:9] -= 20
    def two_events(a): a[2:5] -= 20; a[6:10] -= 20
    np.testing.assert_array_equal(f(ts(one)).values, [1])
    np.testing.assert_array_equal(f(ts(three)).values, [3])
    np.testing.assert_array_equal(f(ts(two_equal)).values, [3])
    np.testing.assert_array_equal(f(ts(two_events)).values, [4])
    np.testing.assert_array_equal(f(ts(two_events, units="C", base=5.0)).values, [4])

    def one_nan(a): a[2] -= 20; a[-1] = np.nan
    np.testing.assert_array_equal(f(ts(one_nan)).values, [np.nan])


def test_atmos_cold_spell_days_and_frequency(backend):   # tests/test_temperature.py:368-403
    from xclim_b200 import atmos
    a = np.zeros(365)
    a[10:20] -= 15
    a[40:43] -= 50
    a[80:100] -= 30
    for ts in (series(a + K2C), series(a, "C")):
        np.testing.assert_array_equal(atmos.cold_spell_days(ts, thresh="-10 C", freq="MS").values,
                                      [10, 0, 12, 8, 0, 0, 0, 0, 0, 0, 0, 0])
        np.testing.assert_array_equal(atmos.cold_spell_frequency(ts, thresh="-10 C", freq="MS").values,
                                      [1, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0])
    b = a + K2C
    b[-1] = np.nan
    np.testing.assert_array_equal(atmos.cold_spell_days(series(b), thresh="-10 C", freq="MS").values,
                                  [10, 0, 12, 8, 0, 0, 0, 0, 0, 0, 0, np.nan])
    np.testing.assert_array_equal(atmos.cold_spell_frequency(series(b), thresh="-10 C", freq="MS").values,
                                  [1, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, np.nan])


def test_atmos_heat_wave_index(backend):                  # tests/test_temperature.py:821-845
    from xclim_b200 import atmos
    tx = np.zeros(366)
    tx[:10] = np.array([29, 31, 31, 31, 29, 31, 31, 31, 31, 31])
    np.testing.assert_array_equal(atmos.heat_wave_index(series(tx + K2C, start="2000-01-01"), freq="YS").values, [10])
    np.testing.assert_array_equal(atmos.heat_wave_index(series(tx, "C", start="2000-01-01"), freq="YS").values, [10])
    tx[-1] = np.nan
    np.testing.assert_array_equal(atmos.heat_wave_index(series(tx + K2C, start="2000-01-01"), freq="YS").values,
                                  [np.nan])


def test_atmos_dry_spell_total_and_max_length_with_missing_day(backend):   # tests/test_precip.py:645-674
    """Without the select_time indexer: the NaN on 1 January masks January, no 7-day window totalling
    less than 3.1 mm exists after it."""
    from xclim_b200 import atmos, indices
    pr = series([np.nan] + [1] * 4 + [0] * 10 + [1] * 350, "mm/d", start="1900-01-01")
    out = atmos.dry_spell_total_length(pr, window=7, op="sum", thresh="3.1 mm", freq="MS")
    np.testing.assert_allclose(out.values, [np.nan] + [0] * 11)
    out = atmos.dry_spell_max_length(pr, window=7, op="sum", thresh="3.1 mm", freq="MS")
    np.testing.assert_allclose(out.values, [np.nan] + [0] * 11)
    # the index itself (no missing-value mask) sees the dry spell of January: the ten zeros plus the
    # days whose 7-day windows still total less than 3.1 mm
    raw = indices.dry_spell_total_length(pr, window=7, op="sum", thresh="3.1 mm", freq="MS").values
    assert raw[0] > 0 and (raw[1:] == 0).all()


@pytest.mark.parametrize("name", ["tg90p", "tn90p", "tx90p"])
def test_atmos_t90p(backend, name):                       # tests/test_temperature.py:1090-1181
    from xclim_b200 import atmos, calendar as xcal
    arr = np.arange(366, dtype=np.float32)
    t90 = xcal.select_percentile(xcal.percentile_doy(series(arr, start="2000-01-01"), window=1, per=90), 90.0)
    x = arr.copy()
    x[175:180] = 1
    out = getattr(atmos, name)(series(x, start="2000-01-01"), t90, freq="MS").values
    assert out[0] == 30 and out[1] == 29 and out[5] == 25
    if True:                 # the same data in degC against the table in K (`outC` of the reference test)
        out_c = getattr(atmos, name)(series(x - K2C, "C", start="2000-01-01"), t90, freq="MS").values
        # (January apart: day 1 sits exactly ON its threshold, 0 K, and -273.15 is not a float32)
        np.testing.assert_array_equal(out_c[1:], out[1:])
    x[33] = np.nan
    out = getattr(atmos, name)(series(x, start="2000-01-01"), t90, freq="MS").values
    assert out[0] == 30 and np.isnan(out[1]) and out[5] == 25


@pytest.mark.parametrize("name", ["tg10p", "tn10p", "tx10p"])
def test_atmos_t10p(backend, name):                       # tests/test_temperature.py:1197-1288
    from xclim_b200 import atmos, calendar as xcal
    arr = np.arange(366, dtype=np.float32)
    t10 = xcal.select_percentile(xcal.percentile_doy(series(arr, start="2000-01-01"), per=10), 10.0)
    x = arr.copy()
    x[175:180] = 1
    out = getattr(atmos, name)(series(x, start="2000-01-01"), t10, freq="MS").values
    assert out[0] == 0 and out[5] == 5
    x[33] = np.nan
    out = getattr(atmos, name)(series(x, start="2000-01-01"), t10, freq="MS").values
    assert out[0] == 0 and np.isnan(out[1]) and out[5] == 5


def test_atmos_frost_season_length_incomplete_periods(backend):   # tests/test_temperature.py:351-365
    """The first and last YS-JUL periods of a 2000-01-01 .. 2001-12-31 series are incomplete: MissingAny
    compares the valid count with the length of the COMPLETE period (core/missing.py:64-160)."""
    from xclim_b200 import atmos
    a = np.zeros(731) + K2C + 15
    a[300:400] = K2C - 5
    a[404:407] = K2C - 5
    tasmin = series(a, start="2000-01-01")
    np.testing.assert_array_equal(atmos.frost_season_length(tasmin).values, [np.nan, 107, np.nan])
    np.testing.assert_array_equal(atmos.frost_season_length(tasmin, window=3).values, [np.nan, 100, np.nan])
    np.testing.assert_array_equal(atmos.frost_season_length(tasmin, mid_date="07-01", freq="YS").values, [0, 181])


def test_atmos_growing_season_length(backend):            # tests/test_temperature.py:904-958
    """Warm May..August (>= 5.5 degC above freezing), 0 degC otherwise: the season is those 123 days."""
    from xclim_b200 import atmos
    rng = np.random.default_rng(21)

    def year_series(n_years=1, units="K", nan_at=None):
        vals, starts = [], 0
        ta = series(np.zeros(366 * n_years), start="2000-01-01").time
        v = np.zeros(len(ta)) + (K2C if units == "K" else 0)
        tt = (ta.month >= 5) & (ta.month <= 8)
        v[tt] += rng.uniform(5.5, 23, size=int(tt.sum()))
        if nan_at is not None:
            v[nan_at] = np.nan
        return series(v, units, start="2000-01-01"), tt

    ts, tt = year_series()
    np.testing.assert_array_equal(atmos.growing_season_length(ts).values, [tt.sum()])
    ts, tt = year_series(units="C")
    np.testing.assert_array_equal(atmos.growing_season_length(ts).values, [tt.sum()])
    ts, _ = year_series(units="C", nan_at=50)
    np.testing.assert_array_equal(atmos.growing_season_length(ts).values, [np.nan])
    ts, tt = year_series(n_years=10, units="C")
    out = atmos.growing_season_length(ts).values
    assert out[3] == tt[:366].sum() and np.isnan(out[-1])   # 3660 days end on 2010-01-07: the last year is incomplete


@pytest.mark.parametrize("name", ["tg10p", "tx10p", "tn10p"])
def test_index_t10p_and_dayofyear_requirement(backend, name):   # :2529-2570
    from xclim_b200 import calendar as xcal, indices
    arr = np.arange(366, dtype=np.float32)
    tas = series(arr, start="2000-01-01")
    t10 = xcal.select_percentile(xcal.percentile_doy(tas, per=10), 10.0)
    x = arr.copy()
    x[175:180] = 1
    out = getattr(indices, name)(series(x, start="2000-01-01"), t10, freq="MS").values
    assert out[0] == 0 and out[5] == 5
    with pytest.raises(AttributeError, match="dayofyear"):
        getattr(indices, name)(tas, tas, freq="MS")


# ---- tests/test_missing.py (values only) --------------------------------------------------------------------
def test_missing_any_reference_known_answers(backend):     # :55-98
    from xclim_b200 import missing
    a = np.arange(360.0)
    a[5:10] = np.nan
    out = missing.missing_any(series(a), freq="MS").values
    assert out[0] and not out[1]
    np.testing.assert_array_equal(missing.missing_any(series(np.arange(66), "", "2001-12-30"), "MS").values,
                                  [True, False, False, True])
    np.testing.assert_array_equal(missing.missing_any(series(np.arange(378), "", "2001-12-31"), "YS").values,
                                  [True, False, True])
    np.testing.assert_array_equal(missing.missing_any(series(np.arange(378), "", "2001-12-31"), "QE-NOV").values,
                                  [True, False, False, False, True])
    b = np.zeros(365) + K2C + 5.0
    b[2] -= 20
    np.testing.assert_array_equal(missing.missing_any(series(b), freq="YS-JUL").values, [False])
    np.testing.assert_array_equal(missing.missing_any(series(b), freq="YE-JUN").values, [False])


def test_missing_wmo_reference_known_answers(backend):     # :165-197
    from xclim_b200 import missing
    a = np.arange(360.0)
    a[5:7] = np.nan
    a[40:45] = np.nan
    a[70:92:2] = np.nan
    out = missing.missing_wmo(series(a), freq="MS").values
    assert not out[0] and out[1] and out[2]
    a = np.arange(350.0)
    a[5:16] = np.nan
    np.testing.assert_array_equal(missing.missing_wmo(series(a), freq="QS-JAN").values, [True, False, False, True])
    np.testing.assert_array_equal(missing.missing_wmo(series(np.arange(31.0)), freq="YS").values, [True])


def test_missing_pct_and_at_least_n_reference_known_answers(backend):   # :199-230
    from xclim_b200 import missing
    a = np.arange(360.0)
    a[5:7] = np.nan
    a[40:45] = np.nan
    out = missing.missing_pct(series(a), freq="MS", tolerance=0.1).values
    assert not out[0] and out[1]
    a = np.arange(360.0)
    a[5:10] = np.nan
    a[40:55] = np.nan
    np.testing.assert_array_equal(missing.at_least_n_valid(series(a), freq="MS", n=20).values[:2], [False, True])


def _ar1(alpha, n, rng, positive=False):
    """tests/test_bootstrapping.py `ar1`: a red-noise series (values only matter statistically)."""
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):
        x[i] = alpha * x[i - 1] + rng.standard_normal()
    return np.abs(x) if positive else x + 280


@pytest.mark.parametrize("name,p,freq", [("tg90p", 98, "MS"), ("tn90p", 98, "YS-JUL"), ("tx90p", 98, "QS-APR"),
                                         ("tn10p", 2, "MS"), ("tx10p", 2, "YS"), ("tg10p", 2, "MS")])
def test_bootstrap_property_standard_calendar(backend, name, p, freq):   # tests/test_bootstrapping.py:22-74
    """Four years on the standard calendar, base 2000-2001 (a leap and a common year: unequal blocks; with
    anchored frequencies the year groups are partial at both ends): bootstrapping raises the in-base index
    more often than it lowers it and leaves the out-of-base periods untouched."""
    from xclim_b200 import calendar as xcal, indices
    rng = np.random.default_rng(int(p) + len(freq))
    n = int(4 * 365.25)
    da = series(_ar1(0.8, n, rng), start="2000-01-01")
    base = da.isel_time(da.time.sel_years(2000, 2001))
    per = xcal.select_percentile(xcal.percentile_doy(base, per=p), float(p))
    fn = getattr(indices, name)
    plain = fn(da, per, freq=freq, bootstrap=False).values.astype(np.float64)
    boot = fn(da, per, freq=freq, bootstrap=True).values
    labels = np.array([int(str(s)[:4]) * 100 + int(str(s)[5:7]) for s in da.time.period_labels(freq)])
    lengths = np.diff(da.time.period_offsets(freq))
    starts = da.time.period_offsets(freq)[:-1]
    in_base = da.time.year[starts] <= 2001                       # the period starts inside the base
    ends_in_base = da.time.year[starts + lengths - 1] <= 2001
    inside, outside = in_base & ends_in_base, ~in_base
    assert np.count_nonzero(boot[inside] > plain[inside]) > np.count_nonzero(boot[inside] < plain[inside])
    np.testing.assert_array_almost_equal(boot[outside], plain[outside], 15)
    assert labels.size == boot.shape[0]


def test_precip_average(backend):                         # :2441-2465
    import pandas as pd
    from xclim_b200 import indices
    pr = np.zeros(100)
    pr[5:10] = 1
    np.testing.assert_allclose(indices.precip_average(series(pr, "mm/d"), freq="MS").values[0], 5 / 31, rtol=1e-6)
    t = pd.date_range("2000-01-01", "2010-12-31", freq="D")
    out = indices.precip_average(series(t.year.values, "mm d-1", start="2000-01-01"))
    np.testing.assert_allclose(out.values, np.arange(2000, 2011))
    assert out.attrs["units"] == "mm"


# ---- select_time indexers at the indicator level ------------------------------------------------------------
def test_missing_any_with_indexers(backend):              # tests/test_missing.py:100-124
    from xclim_b200 import missing
    ts = series(np.zeros(36))                              # 2000-07-01 .. 2000-08-05
    np.testing.assert_array_equal(missing.missing_any(ts, freq="YS", month=7).values, [False])
    np.testing.assert_array_equal(missing.missing_any(ts, freq="YS", month=8).values, [True])
    np.testing.assert_array_equal(missing.missing_any(ts, freq="YS", month=[7, 8]).values, [True])
    ts = series(np.zeros(76))
    np.testing.assert_array_equal(missing.missing_any(ts, freq="YS", month=[7, 8]).values, [False])
    for cal, n in (("standard", 360), ("noleap", 359), ("360_day", 354)):
        ts = series(np.zeros(n), start="2000-01-01", calendar=cal)   # ends on 25 (24) December
        np.testing.assert_array_equal(missing.missing_any(ts, freq="YS", season="MAM").values, [False])
        np.testing.assert_array_equal(missing.missing_any(ts, freq="YS", season="DJF").values, [True])


def test_atmos_tx90p_seasonal_indexer(backend):           # tests/test_temperature.py:1183-1194
    from xclim_b200 import atmos, calendar as xcal
    arr = np.arange(366, dtype=np.float32)
    t90 = xcal.select_percentile(xcal.percentile_doy(series(arr, start="2000-01-01"), window=1, per=90), 90.0)
    x = arr.copy()
    x[175:180] = 1
    for fn in (atmos.tg90p, atmos.tx90p):                  # generic wrapper and the hand-fused entry point
        out = fn(series(x, start="2000-01-01"), t90, freq="YS", season="JJA")
        assert out.values[0] == 87  # 92 JJA days minus the five cold ones


def test_atmos_dry_spell_with_date_bounds(backend):       # tests/test_precip.py:645-674
    """The indexer masks the INPUT (NaN before 10 January), then the index runs: three 7-day windows
    (starting 10, 11, 12 January) total less than 3.1 mm and cover nine days."""
    from xclim_b200 import atmos
    pr = series([np.nan] + [1] * 4 + [0] * 10 + [1] * 350, "mm/d", start="1900-01-01")
    out = atmos.dry_spell_total_length(pr, window=7, op="sum", thresh="3.1 mm", freq="MS", date_bounds=("01-10", "12-31"))
    np.testing.assert_allclose(out.values, [9] + [0] * 11)
