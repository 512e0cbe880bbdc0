"""Known answers of the reference's own test-suite (VALUES only, re-expressed on the Field container),
asserted through the host layer twice: on the CPU with the oracle stand-ins of tests/fake_device.py, and
on the GPU through the real kernels (-m gpu)."""
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


def test_first_and_last_run(backend):
    """tests/test_run_length.py:313-331, 384-433: run at steps 30..39."""
    from xclim_b200 import run_length as rl
    t = np.zeros((60, 2), np.float32)
    t[30:40] = 2
    da = make_field(t, "2000-01-01", units="K")
    assert (rl.first_run(da, 1, freq="YS").values == 30).all()
    assert (rl.first_run(da, 1, freq="YS", coord="dayofyear").values == 31).all()
    assert (rl.last_run(da, 1, freq="YS").values == 39).all()
    assert (rl.last_run(da, 1, freq="YS", coord="dayofyear").values == 40).all()
    t[0] = 2                                   # :411-425 resample after: [30, 8] / doy [31, 40]
    da = make_field(t, "2000-01-01", units="K")
    np.testing.assert_array_equal(rl.last_run(da, 1, freq="MS").values[:, 0], [30, 8])
    np.testing.assert_array_equal(rl.last_run(da, 1, freq="MS", coord="dayofyear").values[:, 0], [31, 40])
    a = np.zeros((100, 1), np.float32)         # :302-306
    a[10:20] = 1
    assert rl.first_run(make_field(a, "2000-01-01", units=""), 5, freq="YS").values[0, 0] == 10


def test_windowed_max_run_sum(backend):
    """tests/test_run_length.py:374-381: 2 steps too short, 5 steps, 10 steps of 5 -> 50."""
    from xclim_b200 import run_length as rl
    a = np.zeros((50, 1), np.float32)
    a[4:6] = 5
    a[25:30] = 5
    a[35:45] = 5
    out = rl.windowed_max_run_sum(make_field(a, "2001-01-01", calendar="noleap", units=""), 3, freq="YS")
    assert out.values[0, 0] == 50


@pytest.mark.parametrize("op,expected", [("gt", [0, 5, 10, 0, 0]), (">=", [0, 5, 10, 0, 0]), ("<", [20, 0, 0, 7, 0])])
def test_cumulative_difference(backend, op, expected):
    """tests/test_generic.py:316-334 (per-step values: one period per step here, freq="D")."""
    from xclim_b200 import generic
    tas = make_field((np.array([-10, 15, 20, 3, 10]) + K2C).astype(np.float32), "2000-01-01", units="K")
    out = generic.cumulative_difference(tas, threshold="10 degC", op=op, freq="D")
    np.testing.assert_allclose(out.values, expected, rtol=1e-5, atol=1e-4)
    out_k = generic.cumulative_difference(tas, threshold="283.15 K", op=op, freq="D")
    np.testing.assert_allclose(out.values, out_k.values)
    with pytest.raises((NotImplementedError, ValueError)):
        generic.cumulative_difference(tas, threshold="10 degC", op="!=", freq="D")


def test_spell_length_with_holes(backend):
    """tests/test_run_length.py:150-162 through spell_length_statistics(min_gap=3)."""
    from xclim_b200 import generic
    v = np.zeros(365, np.float32)
    a = [0, 1, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0]
    v[:len(a)] = a
    da = make_field(v, "2000-01-01", calendar="noleap", units="")
    mx, total, n = generic.spell_length_statistics(da, 0.5, 1, None, ">", ["max", "sum", "count"], "YS", min_gap=3)
    assert mx.values[0] == 10 and total.values[0] == 15 and n.values[0] == 2


def test_atmos_tx90p_simple_and_nan_treatment(backend):
    """tests/test_temperature.py:1153-1181: leap year 0..365, window-1 percentiles (the table is the series
    itself re-mapped onto 366 days), cold spell in June; a NaN day turns its month into NaN (MissingAny)."""
    from xclim_b200 import atmos, calendar as xcal
    arr = np.arange(366, dtype=np.float32)
    tas = make_field(arr.copy(), "2000-01-01", units="K")
    t90 = xcal.select_percentile(xcal.percentile_doy(tas, window=1, per=90), 90.0)
    x = arr.copy()
    x[175:180] = 1
    out = atmos.tx90p(make_field(x, "2000-01-01", units="K"), t90, freq="MS").values
    assert out[0] == 30 and out[1] == 29 and out[5] == 25
    x[33] = np.nan
    out = atmos.tx90p(make_field(x, "2000-01-01", units="K"), t90, freq="MS").values
    assert out[0] == 30 and np.isnan(out[1]) and out[5] == 25


def test_atmos_cdd_vector(backend):
    """tests/test_precip.py:470-477: 31 wet days of 50 mm/day with five dry days -> 5."""
    from xclim_b200 import atmos
    x1 = np.full(31, 50.0, np.float32)
    x1[5:10] = 0
    out = atmos.maximum_consecutive_dry_days(make_field(x1, "2000-01-01", units="mm/day"), freq="MS")
    assert out.values[0] == 5
    x1[20] = np.nan
    out = atmos.maximum_consecutive_dry_days(make_field(x1, "2000-01-01", units="mm/day"), freq="MS")
    assert np.isnan(out.values[0])


@pytest.mark.parametrize("op_high,op_low,expected", [(">", "<", 1), (">", "<=", 2), (">=", "<", 3), (">=", "<=", 4)])
def test_count_level_crossings(backend, op_high, op_low, expected):
    """tests/test_generic.py:412-431."""
    from xclim_b200 import generic
    tasmin = make_field((np.array([-1, -3, 0, 5, 9, 1, 3]) + K2C).astype(np.float32), "2000-01-01", units="K")
    tasmax = make_field((np.array([5, 7, 3, 6, 13, 5, 4]) + K2C).astype(np.float32), "2000-01-01", units="K")
    out = generic.count_level_crossings(tasmin, tasmax, threshold="5 degC", freq="YS", op_high=op_high, op_low=op_low)
    np.testing.assert_array_equal(out.values, [expected])


@pytest.mark.parametrize("op_high,op_low", [("<=", "<="), (">=", ">="), ("<", ">"), ("==", "!=")])
def test_count_level_crossings_forbidden_ops(backend, op_high, op_low):
    """tests/test_generic.py:433-446."""
    from xclim_b200 import generic
    tasmin = make_field((np.zeros(7) + K2C).astype(np.float32), "2000-01-01", units="K")
    tasmax = make_field((np.ones(7) + K2C).astype(np.float32), "2000-01-01", units="K")
    with pytest.raises(ValueError):
        generic.count_level_crossings(tasmin, tasmax, threshold="0.5 degC", freq="YS", op_high=op_high, op_low=op_low)


@pytest.mark.parametrize("op,constrain,expected,should_fail", [
    ("<", ("!=", "<"), 4, False), (">", (">", "<="), 5, False), (">=", (">=", "=="), 6, False),
    ("==", ("==", "!="), 1, False), ("==", (">", ">="), 1, True), ("!=", ("!=", ">"), 9, False),
    ("!=", (">", "=="), 9, True), ("%", ("%", "$", "@"), 5.29e-11, True)])
def test_count_occurrences(backend, op, constrain, expected, should_fail):
    """tests/test_generic.py:448-469."""
    from xclim_b200 import generic
    tas = make_field((np.arange(10) + K2C).astype(np.float32), "2000-01-01", units="K")
    if should_fail:
        with pytest.raises(ValueError):
            generic.count_occurrences(tas, "4 degC", freq="YS", op=op, constrain=constrain)
    else:
        out = generic.count_occurrences(tas, "4 degC", freq="YS", op=op, constrain=constrain)
        np.testing.assert_array_equal(out.values, [expected])


@pytest.mark.parametrize("op,constrain,expected,should_fail", [
    ("<", None, np.nan, False), ("<=", None, 3, False), ("!=", ("!=",), 1, False), ("==", ("==", "!="), 3, False),
    ("==", (">=", ">", "<"), 3, True)])
def test_first_occurrence(backend, op, constrain, expected, should_fail):
    """tests/test_generic.py:469-487."""
    from xclim_b200 import generic
    tas = make_field((np.array([15, 12, 11, 12, 14, 13, 18, 11, 13]) + K2C).astype(np.float32), "2000-01-01", units="K")
    if should_fail:
        with pytest.raises(ValueError):
            generic.first_occurrence(tas, threshold="11 degC", freq="YS", op=op, constrain=constrain)
    else:
        out = generic.first_occurrence(tas, threshold="11 degC", freq="YS", op=op, constrain=constrain)
        np.testing.assert_array_equal(out.values, [expected])


@pytest.mark.parametrize("op,constrain,expected,should_fail", [
    ("<", None, np.nan, False), ("<=", None, 8, False), ("!=", ("!=",), 9, False), ("==", ("==", "!="), 8, False),
    ("==", (">=", ">", "<"), 5, True)])
def test_last_occurrence(backend, op, constrain, expected, should_fail):
    """tests/test_generic.py:489-507."""
    from xclim_b200 import generic
    tas = make_field((np.array([15, 12, 11, 12, 14, 13, 18, 11, 13]) + K2C).astype(np.float32), "2000-01-01", units="K")
    if should_fail:
        with pytest.raises(ValueError):
            generic.last_occurrence(tas, threshold="11 degC", freq="YS", op=op, constrain=constrain)
    else:
        out = generic.last_occurrence(tas, threshold="11 degC", freq="YS", op=op, constrain=constrain)
        np.testing.assert_array_equal(out.values, [expected])


@pytest.mark.parametrize("thresholds", [{}, {"thresh_tasmax": "0 degC", "thresh_tasmin": "0 degC"}])
def test_freezethaw_cycles(backend, thresholds):
    """tests/test_indices.py:1415-1439: five freeze-thaw days in January, one in February."""
    from xclim_b200 import indices
    mn = np.zeros(365, np.float32)
    mx = np.zeros(365, np.float32)
    mn[10:20] -= 1
    mx[10:15] += 1
    mn[40:44] += [1, 1, -1, -1]
    mx[40:44] += [1, -1, 1, -1]
    tn = make_field((mn + K2C).astype(np.float32), "2001-01-01", calendar="noleap", units="K")
    tx = make_field((mx + K2C).astype(np.float32), "2001-01-01", calendar="noleap", units="K")
    out = indices.multiday_temperature_swing(tn, tx, **thresholds, op="sum", window=1, freq="MS")
    np.testing.assert_array_equal(out.values[:2], [5, 1])
    np.testing.assert_array_equal(out.values[2:], 0)
    np.testing.assert_array_equal(indices.daily_freezethaw_cycles(tn, tx, freq="MS").values, out.values)
    # one event of five days and one of one day; the longest lasts five days
    assert indices.multiday_temperature_swing(tn, tx, op="count", freq="YS").values[0] == 2
    assert indices.multiday_temperature_swing(tn, tx, op="max", freq="YS").values[0] == 5
    assert indices.multiday_temperature_swing(tn, tx, op="count", window=2, freq="YS").values[0] == 1


def test_high_precip_low_temp(backend):
    """tests/test_indices.py:3727-3732."""
    from xclim_b200 import indices
    pr = make_field(np.array([0, 1, 2, 0], np.float32), "2000-01-01", units="kg m-2 s-1")
    tas = make_field((np.array([0, 0, 1, 1]) + K2C).astype(np.float32), "2000-01-01", units="K")
    out = indices.high_precip_low_temp(pr, tas, pr_thresh="1 kg m-2 s-1", tas_thresh="1 degC")
    np.testing.assert_array_equal(out.values, [1])


def test_snowfall_frequency_intensity_and_dates(backend):
    """tests/test_indices.py:4485-4521: [0, 2, .3, .2, 4] mm/day -> 40 % of the days, mean 3 mm/day."""
    from xclim_b200 import indices
    v = np.array([0, 2, 0.3, 0.2, 4], np.float32)
    prsnd = make_field(v, "2000-07-01", units="mm/day")
    np.testing.assert_allclose(indices.snowfall_frequency(prsnd).values, [40])
    np.testing.assert_allclose(indices.snowfall_intensity(prsnd).values, [3])
    prsn = make_field((v / 86400.0).astype(np.float32), "2000-07-01", units="kg m-2 s-1")
    np.testing.assert_allclose(indices.snowfall_frequency(prsn).values, [40])
    np.testing.assert_allclose(indices.snowfall_intensity(prsn).values, [3], rtol=1e-5)
    # first / last day with snowfall >= 1 mm/day: 2 July (doy 184 in a leap year) and 5 July
    assert indices.first_snowfall(prsnd).values[0] == 184
    assert indices.last_snowfall(prsnd).values[0] == 187
    none = make_field(np.zeros(5, np.float32), "2000-07-01", units="mm/day")
    assert indices.snowfall_intensity(none).values[0] == 0 and np.isnan(indices.first_snowfall(none).values[0])


def test_threshold_and_domain_count(backend):
    """tests/test_generic.py:69-81 (year-end labels)."""
    from xclim_b200 import generic
    ts = make_field(np.arange(365, dtype=np.float32), "2000-07-01", units="K")
    np.testing.assert_array_equal(generic.threshold_count(ts, "<", 50, "YE").values, [50, 0])
    np.testing.assert_array_equal(generic.domain_count(ts, low=10, high=20, freq="YE").values, [10, 0])


def test_rolling_resample_known_answers(backend):
    """tests/test_generic.py:35-67: q = 1 .. 1096 over 2000-2002."""
    from xclim_b200 import generic
    q = make_field(np.arange(1, 366 + 365 + 365 + 1, dtype=np.float32), "2000-01-01", units="m3 s-1")
    o = generic.select_rolling_resample_op(q, "max", window=14, window_center=False, window_op="mean")
    np.testing.assert_array_equal(o.values, [np.mean(np.arange(353, 367)), np.mean(np.arange(353 + 365, 367 + 365)),
                                             np.mean(np.arange(353 + 730, 367 + 730))])
    assert o.attrs["units"] == "m3 s-1"
    o = generic.select_rolling_resample_op(q, "max", window=3, window_center=True, window_op="sum", freq="MS")
    np.testing.assert_array_equal(o.values[:2], [30 + 31 + 32, 59 + 60 + 61])
