"""GPU parity: seasons and date-bounded runs vs the per-group restatement + reference known answers."""
import numpy as np
import pytest

from oracle import xclim_oracle as O
from xb_helpers import make_field

pytestmark = pytest.mark.gpu


def _tas(rng, T, shape):
    t = np.arange(T)
    x = 278 + 12 * np.sin(2 * np.pi * (t % 365 - 110) / 365)[(slice(None),) + (None,) * len(shape)]
    x = (x + 5 * rng.standard_normal((T,) + shape)).astype(np.float32)
    x[rng.random(x.shape) < 0.004] = np.nan
    return x


@pytest.mark.parametrize("freq,mid", [("YS", "07-01"), ("YS-JUL", "01-01"), ("YS", None)])
@pytest.mark.parametrize("window", [1, 5])
def test_season_parity(cuda, freq, mid, window):
    from xclim_b200 import generic
    rng = np.random.default_rng(71)
    x = _tas(rng, 365 * 4 + 100, (5, 7))
    x[:, 0, 0] = 300.0   # always in season
    x[:, 0, 1] = 200.0   # never
    da = make_field(x, "2001-01-01", calendar="noleap", units="K")
    ta = da.time
    poff = ta.period_offsets(freq)
    thr = 278.15
    cond = O.compare(x, ">=", thr)
    mids = [None] * (len(poff) - 1)
    if mid is not None:
        mids = [int(m) - int(s) if m >= 0 else None for m, s in zip(ta.date_index_in_periods(freq, mid), poff[:-1])]
    for stat in ("start", "end", "length"):
        got = generic.season(da, thr, window, ">=", stat, freq, mid_date=mid).values
        exp = O.season(cond, window, mids, poff, stat, ta.doy, has_date=mid is not None)
        np.testing.assert_array_equal(got, exp, err_msg=f"{stat} {freq} {mid} w={window}")


def test_date_bounded_runs_parity(cuda):
    from xclim_b200 import generic, indices, seasons, _lib
    from xclim_b200.generic import _unwrap
    rng = np.random.default_rng(72)
    x = _tas(rng, 365 * 3, (4, 6))
    da = make_field(x, "2001-01-01", calendar="noleap", units="K")
    ta = da.time
    poff = ta.period_offsets("YS")
    mids = [int(m) - int(s) for m, s in zip(ta.date_index_in_periods("YS", "07-01"), poff[:-1])]
    cond_lt = O.compare(x, "<", 273.15)
    x2d, cs, other, _ = _unwrap(da)

    def per_group(fn, cond, w):
        outs = []
        for p, (s, e) in enumerate(zip(poff[:-1], poff[1:])):
            v = fn(cond[s:e], w, mids[p])
            t = np.where(np.isnan(v), 0, v).astype(int) + s
            outs.append(np.where(np.isnan(v), np.nan, ta.doy[t]))
        return np.stack(outs)
    for w in (1, 3):
        got = indices.first_day_temperature_below(da, window=w).values
        np.testing.assert_array_equal(got, per_group(O.first_run_after_date, cond_lt, w))
        got = indices.last_spring_frost(da, window=w).values
        np.testing.assert_array_equal(got, per_group(O.last_run_before_date, cond_lt, w))
        got = seasons.run_end_after_date(x2d, ta, "YS", _lib.OPS["<"], 273.15, w, "07-01").cpu().numpy().reshape((3, 4, 6))
        np.testing.assert_array_equal(got, per_group(O.run_end_after_date, cond_lt, w))


@pytest.mark.parametrize("d1,d2,mid_date,expected", [
    ("1950-01-01", "1951-01-01", "07-01", np.nan), ("2000-01-01", "2000-12-31", "07-01", 365),
    ("2000-07-10", "2001-01-01", "07-01", np.nan), ("2000-06-15", "2000-07-15", "07-01", 198),
    ("2000-06-15", "2000-07-25", "07-15", 208), ("2000-06-15", "2000-07-15", "10-01", 275),
    ("2000-06-15", "2000-07-15", "01-10", np.nan), ("2000-06-15", "2000-07-15", "06-15", np.nan)])
def test_growing_season_reference_known_answers(cuda, d1, d2, mid_date, expected):
    """tests/test_indices.py:1654-1700."""
    import pandas as pd
    from xclim_b200 import indices
    idx = pd.date_range("2000-01-01", periods=365, freq="D")
    tas = np.zeros(365, np.float32)
    tas[(idx >= d1) & (idx <= d2)] = 280
    da = make_field(tas, "2000-01-01", units="K")
    out = indices.growing_season_end(da, mid_date=mid_date)
    np.testing.assert_array_equal(out.values[0], expected)
    assert out.attrs["is_dayofyear"] == 1 and out.attrs["units"] == ""
    if mid_date == "07-01":
        exp_len = {"1950-01-01": 0, "2000-01-01": 365, "2000-07-10": 0, "2000-06-15": 31}[d1]
        assert indices.growing_season_length(da).values[0] == exp_len
