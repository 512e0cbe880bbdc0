"""GPU parity: first_run / last_run per period vs the whole-array restatement."""
import numpy as np
import pytest

from oracle import xclim_oracle as O
from xb_helpers import make_field

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("window", [1, 2, 5])
@pytest.mark.parametrize("freq", ["YS", "MS"])
def test_first_last_run(cuda, window, freq):
    from xclim_b200 import run_length as rl
    rng = np.random.default_rng(51)
    m = rng.random((800, 5, 7)) < 0.55
    m[:, 0, 0] = True       # all True: NaN for window == 1 (argmax == argmin quirk)
    m[:, 0, 1] = False      # all False: NaN
    da = make_field(m, "2000-01-01", units="")
    poff = da.time.period_offsets(freq)
    for fn, ofn in ((rl.first_run, O.first_run), (rl.last_run, O.last_run)):
        got = fn(da, window, freq=freq).values
        exp = ofn(m, window, poff=poff)
        np.testing.assert_array_equal(got, exp, err_msg=f"{fn.__name__} w={window} {freq}")
    doy = rl.first_run(da, window, freq=freq, coord="dayofyear").values
    idx = O.first_run(m, window, poff=poff)
    t = np.where(np.isnan(idx), 0, idx).astype(int) + poff[:-1, None, None]
    np.testing.assert_array_equal(doy, np.where(np.isnan(idx), np.nan, da.time.doy[t]))
