"""CPU-only exercise of the HOST layer (units, thresholds, operators, frequencies, wrapping, options)
with the device functions replaced by oracle stand-ins (tests/fake_device.py).  Kernel parity is the job
of the -m gpu tests; this file pins the glue above the C ABI without a GPU."""
import inspect

import numpy as np
import pytest

import fake_device
from oracle import xclim_oracle as O
from xb_helpers import make_field

import test_gpu_batch as batch   # reuse the oracle compositions of the 50 indicators


@pytest.fixture
def host(monkeypatch):
    fake_device.install(monkeypatch)


def test_batch_of_50_host_glue(host):
    from xclim_b200 import calendar as xcal, indices
    data = batch._inputs()
    units = {"tas": "K", "tasmax": "K", "tasmin": "K", "pr": "mm/d"}
    fields = {k: make_field(v, "1981-01-01", calendar="noleap", units=units[k]) for k, v in data.items()}
    for name, var in indices.BATCH_INDICATORS:
        da, x = fields[var], data[var]
        fn = getattr(indices, name)
        freq = inspect.signature(fn).parameters["freq"].default
        poff = da.time.period_offsets(freq)
        if name in ("tx90p", "tx10p", "tn90p"):
            per = 10.0 if name == "tx10p" else 90.0
            pdoy = xcal.select_percentile(xcal.percentile_doy(da, window=5, per=per), per)
            got = fn(da, pdoy)
            tab = O.percentile_doy(x, da.time.year, da.time.doy, 5, per)[:, 0]
            exp = O.doy_threshold_count(x, tab, da.time.doy, poff, "<" if name == "tx10p" else ">")
        else:
            got = fn(da)
            exp = batch._oracle(name, x, poff, da.time, data)
        exp = np.asarray(exp)
        assert got.values.shape == exp.shape, name
        assert "units" in got.attrs, name
        np.testing.assert_allclose(got.values, exp, rtol=1e-5, atol=1e-6, equal_nan=True, err_msg=name)


def test_unit_strings_and_operator_validation(host):
    from xclim_b200 import generic, indices
    rng = np.random.default_rng(3)
    pr = rng.gamma(0.5, 4.0, (365, 2, 3)).astype(np.float32)
    da = make_field(pr, "2001-01-01", calendar="noleap", units="mm/d")
    a = indices.wetdays(da, thresh="1 mm/day")
    b = indices.wetdays(make_field(pr / 86400.0, "2001-01-01", calendar="noleap", units="kg m-2 s-1"), thresh="1 mm/day")
    assert a.attrs["units"] == "d" and a.values.dtype == np.int64
    # the same physical threshold expressed in the data's units counts (almost) the same days
    assert np.abs(a.values - b.values).max() <= 1
    with pytest.raises(ValueError, match="not recognized"):
        generic.threshold_count(da, "=>", 1.0, "YS")
    with pytest.raises(ValueError, match="not permitted"):
        generic.threshold_count(da, "==", 1.0, "YS", constrain=(">", ">="))
    with pytest.raises(NotImplementedError):
        generic.spell_length_statistics(da, 1.0, 3, "sum", "<", "max", "YS", min_gap=2)
    out = generic.spell_length_statistics(da, 1.0, 1, None, "<", "max", "MS", min_gap=2)
    exp = O.spell_length_statistics(pr, 1.0, 1, None, "<", "max", da.time.period_offsets("MS"), min_gap=2)
    np.testing.assert_array_equal(out.values, exp)
    assert len(out.coords["time"]) == 12


def test_bootstrap_unequal_blocks_host_loop(host):
    """The N(N-1) virtual-row loop of bootstrapping.py on a standard calendar: same numbers as the oracle's
    literal block replacement (365 <-> 366 conversion)."""
    from xclim_b200 import calendar as xcal, indices
    rng = np.random.default_rng(4)
    T = 366 + 365 * 3 + 366                      # 2000 .. 2004
    t = np.arange(T)
    x = (288 + 10 * np.sin(2 * np.pi * t / 365.25)[:, None] + 3 * rng.standard_normal((T, 3))).astype(np.float32)
    da = make_field(x, "2000-01-01", calendar="standard", units="K")
    base = da.isel_time(da.time.sel_years(2001, 2004))
    pdoy = xcal.select_percentile(xcal.percentile_doy(base, window=5, per=90.0), 90.0)
    got = indices.tx90p(da, pdoy, freq="YS", bootstrap=True)
    exp = O.bootstrap_doy_count(x, da.time.year, da.time.doy, da.time.period_offsets("YS"), (2001, 2004), window=5,
                                per=90.0, op=">", cal_max_doy=366)
    np.testing.assert_array_equal(got.values, exp)
    assert got.values.dtype == np.float64 and got.attrs["units"] == "d"
