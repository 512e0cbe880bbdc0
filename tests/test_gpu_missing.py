"""GPU parity: missing-value masks (any, pct, at_least_n, wmo) from the fused kernels."""
import numpy as np
import pytest

from oracle import xclim_oracle as O
from xb_helpers import make_field

pytestmark = pytest.mark.gpu


def test_missing_masks(cuda):
    from xclim_b200 import missing
    rng = np.random.default_rng(81)
    x = (280 + rng.standard_normal((365 * 3, 6, 8))).astype(np.float32)
    x[rng.random(x.shape) < 0.02] = np.nan
    x[40:47, 0, 0] = np.nan      # 7 consecutive days in February of year 1
    x[100:112, 0, 1] = np.nan    # 12 days spanning April (>= nm in one month? 11 in April)
    x[:, 0, 2] = np.nan
    da = make_field(x, "2001-01-01", calendar="noleap", units="K")
    ta = da.time
    for freq in ("YS", "MS", "QS-DEC"):
        poff = ta.period_offsets(freq)
        n_exp = ta.expected_period_lengths(freq)      # QS-DEC: the first and last quarters are incomplete
        np.testing.assert_array_equal(missing.missing_any(da, freq).values, O.missing_any(x, poff, n_exp))
        np.testing.assert_array_equal(missing.missing_pct(da, freq, 0.05).values, O.missing_pct(x, poff, 0.05, n_exp))
        if freq == "QS-DEC":
            assert missing.missing_any(da, freq).values[[0, -1]].all()
        np.testing.assert_array_equal(missing.at_least_n_valid(da, freq, 25).values, O.at_least_n_valid(x, poff, 25))
        pm = ta.period_offsets("MS")
        parent = np.searchsorted(poff, pm[:-1], side="right") - 1
        exp = O.missing_wmo(x, pm, parent, len(poff) - 1, expected_month=ta.expected_period_lengths("MS"),
                            months_per_period={"YS": 12, "MS": 1, "QS-DEC": 3}[freq])
        np.testing.assert_array_equal(missing.missing_wmo(da, freq).values, exp, err_msg=freq)
    with pytest.raises(ValueError):
        missing.missing_pct(da, "YS", 1.5)


def test_indicator_level_entry_points_apply_missing_any(cuda):
    """atmos.* = index + MissingAny mask from the fused valid counts (core/indicator.py:1522-1549)."""
    from xclim_b200 import atmos, calendar as xcal
    rng = np.random.default_rng(82)
    T, shape = 365 * 4, (4, 6)
    pr = rng.gamma(0.4, 6.0, size=(T,) + shape).astype(np.float32)
    pr[rng.random(pr.shape) < 0.45] = 0
    pr[400:403, 0, 0] = np.nan
    dpr = make_field(pr, "1981-01-01", calendar="noleap", units="mm/d")
    poff = dpr.time.period_offsets("YS")
    out = atmos.maximum_consecutive_dry_days(dpr)
    exp = np.where(O.missing_any(pr, poff), np.nan, O.maximum_consecutive_dry_days(pr, 1.0, poff))
    np.testing.assert_array_equal(out.values, exp.astype(np.float32))
    assert out.attrs["units"] == "days" and out.name == "cdd" and np.isnan(out.values[1, 0, 0])
    tas = (280 + 5 * rng.standard_normal((T,) + shape)).astype(np.float32)
    tas[10, 1, 1] = np.nan
    dt = make_field(tas, "1981-01-01", calendar="noleap", units="K")
    pm = dt.time.period_offsets("MS")
    got = atmos.tg_mean(dt, freq="MS").values
    exp = np.where(O.missing_any(tas, pm), np.nan, O.tg_mean(tas.astype(np.float64), pm))
    np.testing.assert_allclose(got, exp, rtol=1e-5, equal_nan=True)
    per = xcal.select_percentile(xcal.percentile_doy(dt, per=90.0), 90.0)
    got = atmos.tx90p(dt, per).values
    tab = O.percentile_doy(tas, dt.time.year, dt.time.doy, 5, 90.0)[:, 0]
    exp = np.where(O.missing_any(tas, poff), np.nan, O.doy_threshold_count(tas, tab, dt.time.doy, poff, ">").astype(float))
    np.testing.assert_array_equal(got, exp)
