"""GPU parity: percentile bootstrap (core/bootstrapping.py) vs the literal CPU restatement."""
import numpy as np
import pytest

from oracle import xclim_oracle as O
from xb_helpers import make_field

pytestmark = pytest.mark.gpu


def _tas(rng, T, shape, nan_frac=0.0):
    t = np.arange(T)
    x = 288 + 12 * np.sin(2 * np.pi * (t % 365 - 110) / 365)[(slice(None),) + (None,) * len(shape)]
    x = (x + 3 * rng.standard_normal((T,) + shape)).astype(np.float32)
    if nan_frac:
        x[rng.random(x.shape) < nan_frac] = np.nan
    return x


@pytest.mark.parametrize("nbase,nyears,per,op,freq", [
    (5, 8, 90.0, ">", "YS"),
    (6, 7, 10.0, "<", "YS"),
    (4, 6, 90.0, ">=", "MS"),
    (15, 17, 90.0, ">", "YS"),
])
def test_bootstrap_matches_literal_restatement(cuda, nbase, nyears, per, op, freq):
    from xclim_b200 import calendar as xcal, indices
    rng = np.random.default_rng(31)
    shape = (3, 4) if nbase < 10 else (2, 2)
    x = _tas(rng, 365 * nyears, shape, nan_frac=0.004)
    x[:, 0, 0] = np.round(x[:, 0, 0])          # many exact ties
    da = make_field(x, "1981-01-01", calendar="noleap", units="K")
    base = da.isel_time(da.time.sel_years(1982, 1982 + nbase - 1))  # base does not start at the series start
    pdoy = xcal.select_percentile(xcal.percentile_doy(base, window=5, per=per), per)
    fn = indices.tx90p if op in (">", ">=") else indices.tx10p
    got = fn(da, pdoy, freq=freq, bootstrap=True, op=op)
    exp = O.bootstrap_doy_count(x, da.time.year, da.time.doy, da.time.period_offsets(freq),
                                (1982, 1982 + nbase - 1), window=5, per=per, op=op)
    assert got.values.dtype == np.float64 and got.attrs["units"] == "d"
    np.testing.assert_array_equal(got.values, exp)   # integer sums / (N-1): bit-exact
    # the property the reference pins (tests/test_bootstrapping.py:65-71): out-of-base years equal the plain index
    plain = fn(da, pdoy, freq=freq, bootstrap=False, op=op).values.astype(np.float64)
    yrs = np.array([int(s[:4]) for s in da.time.period_labels(freq)])
    outside = (yrs < 1982) | (yrs > 1982 + nbase - 1)
    np.testing.assert_array_equal(got.values[outside], plain[outside])
    assert (got.values[~outside] != plain[~outside]).any()


def test_bootstrap_error_behaviour(cuda):
    """core/bootstrapping.py:131-168."""
    from xclim_b200 import calendar as xcal, indices, Field
    rng = np.random.default_rng(32)
    x = _tas(rng, 365 * 4, (2, 2))
    da = make_field(x, "1981-01-01", calendar="noleap", units="K")
    p_all = xcal.select_percentile(xcal.percentile_doy(da, per=90.0), 90.0)
    with pytest.raises(KeyError, match="all years are overlapping"):
        indices.tx90p(da, p_all, bootstrap=True)
    other = make_field(x, "1991-01-01", calendar="noleap", units="K")
    with pytest.raises(KeyError, match="no year overlap"):
        indices.tx90p(other, p_all, bootstrap=True)
    nohist = Field(p_all.values, p_all.dims, None, dict(p_all.coords), {"units": "K"})
    with pytest.raises(KeyError, match="percentile_doy"):
        indices.tx90p(da, nohist, bootstrap=True)
