"""Host logic: period boundaries / labels vs pandas resample; calendars."""
import numpy as np
import pandas as pd
import pytest

from xclim_b200.timeaxis import TimeAxis, parse_offset


@pytest.mark.parametrize("start,n", [("2000-07-01", 365), ("1999-11-15", 1200), ("1981-01-01", 3000)])
@pytest.mark.parametrize("freq", ["YS", "YS-JUL", "MS", "ME", "QS-DEC", "QS", "YE", "QE", "YE-JUN", "2MS", "QS-NOV"])
def test_period_offsets_match_pandas(start, n, freq):
    ta = TimeAxis.daily(start, n, "standard")
    idx = pd.date_range(start, periods=n, freq="D")
    assert (ta.year == idx.year).all() and (ta.month == idx.month).all() and (ta.doy == idx.dayofyear).all()
    cnt = pd.Series(np.arange(n), index=idx).resample(freq).count()
    sizes = np.diff(ta.period_offsets(freq))
    keep = cnt.values > 0
    np.testing.assert_array_equal(sizes, cnt.values[keep])
    assert ta.period_labels(freq) == [str(d.date()) for d in cnt.index[keep]]


def test_noleap_and_360():
    ta = TimeAxis.daily("1981-01-01", 10950, "noleap")
    assert ta.year[-1] == 2010 and ta.doy[-1] == 365 and ta.doy.max() == 365
    np.testing.assert_array_equal(ta.period_offsets("YS"), np.arange(31) * 365)
    assert len(ta.period_offsets("MS")) == 361
    t360 = TimeAxis.daily("2000-01-01", 720, "360_day")
    assert t360.doy.max() == 360 and (np.diff(t360.period_offsets("MS")) == 30).all()
    assert ta.max_doy == 365 and t360.max_doy == 360


def test_bootstrap_groups():
    ta = TimeAxis.daily("1999-11-15", 1200)
    np.testing.assert_array_equal(ta.bootstrap_group_ids("MS"), ta.group_ids("YS"))
    np.testing.assert_array_equal(ta.bootstrap_group_ids("YS-JUL"), ta.group_ids("YS-JUL"))
    assert parse_offset("QS-DEC") == (1, "Q", True, "DEC")


@pytest.mark.parametrize("start,n", [("2000-01-01", 731), ("1999-11-17", 900), ("2003-02-28", 400)])
@pytest.mark.parametrize("freq", ["YS", "YS-JUL", "QS-DEC", "MS", "YE", "ME"])
def test_expected_period_lengths_match_pandas(start, n, freq):
    """core/missing.py:64-160 (`expected_count`, daily source): the days between consecutive period labels."""
    import pandas as pd
    from xclim_b200 import TimeAxis
    ta = TimeAxis.daily(start, n)
    idx = pd.date_range(start, periods=n, freq="D")
    res = pd.Series(1, index=idx).resample(freq).count().index
    if freq.endswith("S") or "S-" in freq:
        nxt = res.shift(1, freq=freq)
        exp = (nxt - res).days
    else:
        prev = res.shift(-1, freq=freq)
        exp = (res - prev).days
    np.testing.assert_array_equal(ta.expected_period_lengths(freq), np.asarray(exp))
    assert (ta.expected_period_lengths(freq) >= np.diff(ta.period_offsets(freq))).all()
