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
        np.testing.assert_array_equal(missing.missing_any(da, freq).values, O.missing_any(x, poff))
        np.testing.assert_array_equal(missing.missing_pct(da, freq, 0.05).values, O.missing_pct(x, poff, 0.05))
        np.testing.assert_array_equal(missing.at_least_n_valid(da, freq, 25).values, O.at_least_n_valid(x, poff, 25))
        pm = ta.period_offsets("MS")
        parent = np.searchsorted(poff, pm[:-1], side="right") - 1
        exp = O.missing_wmo(x, pm, parent, len(poff) - 1)
        np.testing.assert_array_equal(missing.missing_wmo(da, freq).values, exp, err_msg=freq)
    with pytest.raises(ValueError):
        missing.missing_pct(da, "YS", 1.5)
