"""GPU parity for the 'batch of 50 atmos indicators' configuration (BASELINE.json configs[4]): every
index entry point of xclim_b200.indices.BATCH_INDICATORS against an oracle composition."""
import numpy as np
import pytest

from oracle import xclim_oracle as O
from xb_helpers import make_field

pytestmark = pytest.mark.gpu

K0 = 273.15


def _inputs():
    rng = np.random.default_rng(61)
    T, shape = 365 * 3, (4, 8)
    t = np.arange(T)
    season = 14 * np.sin(2 * np.pi * (t % 365 - 110) / 365)[:, None, None]
    tas = (278 + season + 4 * rng.standard_normal((T,) + shape)).astype(np.float32)
    spread = np.abs(4 + rng.standard_normal((T,) + shape)).astype(np.float32)
    tasmax = (tas + spread).astype(np.float32)
    tasmin = (tas - spread).astype(np.float32)
    pr = rng.gamma(0.4, 6.0, size=(T,) + shape).astype(np.float32)
    pr[rng.random(pr.shape) < 0.45] = 0
    for a in (tas, tasmax, tasmin, pr):
        a[rng.random(a.shape) < 0.003] = np.nan
    return {"tas": tas, "tasmax": tasmax, "tasmin": tasmin, "pr": pr}


def _spell(x, op, thr, red, window, poff, before, clip=None):
    out = O.resample_and_rl(O.compare(x, op, thr), before, O.rle_statistics, poff=poff, reducer=red, window=window)
    if clip is not None:
        out = np.where(out >= clip, out, 0)
    return out


def _oracle(name, x, poff, ta, data):
    c = lambda d: float(d) + K0  # noqa: E731  degC -> K as a Python float
    R = lambda op: O.select_resample_op(x.astype(np.float64), op, poff)  # noqa: E731
    table = {
        "tg_mean": lambda: R("mean"), "tg_max": lambda: R("max"), "tg_min": lambda: R("min"),
        "tn_mean": lambda: R("mean"), "tn_max": lambda: R("max"), "tn_min": lambda: R("min"),
        "tx_mean": lambda: R("mean"), "tx_max": lambda: R("max"), "tx_min": lambda: R("min"),
        "max_1day_precipitation_amount": lambda: R("max"), "precip_accumulation": lambda: R("sum"),
        "max_n_day_precipitation_amount": lambda: O.select_rolling_resample_op(
            x.astype(np.float64), "max", 1, poff, window_center=False, window_op="sum"),
        "frost_days": lambda: O.threshold_count(x, "<", c(0), poff),
        "ice_days": lambda: O.threshold_count(x, "<", c(0), poff),
        "hot_days": lambda: O.threshold_count(x, ">", c(25), poff),
        "tx_days_above": lambda: O.threshold_count(x, ">", c(25), poff),
        "tx_days_below": lambda: O.threshold_count(x, "<", c(25), poff),
        "tn_days_above": lambda: O.threshold_count(x, ">", c(20), poff),
        "tn_days_below": lambda: O.threshold_count(x, "<", c(-10), poff),
        "tg_days_above": lambda: O.threshold_count(x, ">", c(10), poff),
        "tg_days_below": lambda: O.threshold_count(x, "<", c(10), poff),
        "wetdays": lambda: O.threshold_count(x, ">=", 1.0, poff),
        "dry_days": lambda: O.threshold_count(x, "<", 0.2, poff),
        "wetdays_prop": lambda: O.threshold_count(x, ">=", 1.0, poff) / np.diff(poff)[:, None, None],
        "growing_degree_days": lambda: O.cumulative_difference(x, c(4), ">", poff),
        "cooling_degree_days": lambda: O.cumulative_difference(x, c(18), ">", poff),
        "heating_degree_days": lambda: O.cumulative_difference(x, c(17), "<", poff),
        "daily_pr_intensity": lambda: O.resample_reduce(np.where(O.compare(x, ">=", 1.0), x, 0).astype(np.float64),
                                                        poff, "sum") / O.threshold_count(x, ">=", 1.0, poff),
        "cold_spell_days": lambda: _spell(x, "<", c(-10), "sum", 5, poff, True),
        "cold_spell_frequency": lambda: _spell(x, "<", c(-10), "count", 5, poff, True),
        "cold_spell_max_length": lambda: _spell(x, "<", c(-10), "max", 1, poff, True, clip=1),
        "cold_spell_total_length": lambda: _spell(x, "<", c(-10), "sum", 3, poff, True),
        "hot_spell_frequency": lambda: _spell(x, ">", c(30), "count", 3, poff, True),
        "hot_spell_max_length": lambda: _spell(x, ">", c(30), "max", 1, poff, True, clip=1),
        "hot_spell_total_length": lambda: _spell(x, ">", c(30), "sum", 3, poff, True),
        "hot_spell_max_magnitude": lambda: O.resample_and_rl(
            np.where(np.isnan(x), 0, np.clip(x - np.float32(c(25)), 0, None)).astype(np.float64), True,
            O.windowed_max_run_sum, 3, poff=poff),
        "heat_wave_index": lambda: _spell(x, ">", c(25), "sum", 5, poff, True),
        "frost_free_spell_max_length": lambda: _spell(x, ">=", c(0), "max", 1, poff, True, clip=1),
        "maximum_consecutive_frost_days": lambda: _spell(x, "<", c(0), "max", 1, poff, True),
        "maximum_consecutive_frost_free_days": lambda: _spell(x, ">=", c(0), "max", 1, poff, True),
        "maximum_consecutive_tx_days": lambda: _spell(x, ">", c(25), "max", 1, poff, True),
        "maximum_consecutive_dry_days": lambda: O.maximum_consecutive_dry_days(x, 1.0, poff),
        "maximum_consecutive_wet_days": lambda: O.spell_length_statistics(x, 1.0, 1, None, ">=", "max", poff),
        "dry_spell_frequency": lambda: O.spell_length_statistics(x, 1.0, 3, "sum", "<", "count", poff),
        "dry_spell_total_length": lambda: O.spell_length_statistics(x, 1.0, 3, "sum", "<", "sum", poff),
        "dry_spell_max_length": lambda: O.spell_length_statistics(x, 1.0, 1, "sum", "<", "max", poff),
        "wet_spell_frequency": lambda: O.spell_length_statistics(x, 1.0, 3, "sum", ">=", "count", poff),
    }
    return table[name]()


def test_batch_of_50_indicators(cuda):
    from xclim_b200 import calendar as xcal, indices
    data = _inputs()
    units = {"tas": "K", "tasmax": "K", "tasmin": "K", "pr": "mm/d"}
    fields = {k: make_field(v, "1981-01-01", calendar="noleap", units=units[k]) for k, v in data.items()}
    assert len(indices.BATCH_INDICATORS) == 50
    checked = 0
    for name, var in indices.BATCH_INDICATORS:
        da, x = fields[var], data[var]
        fn = getattr(indices, name)
        defaults = fn.__defaults__ or ()
        import inspect
        freq = inspect.signature(fn).parameters["freq"].default
        poff = da.time.period_offsets(freq)
        if name in ("tx90p", "tx10p", "tn90p"):
            per = 10.0 if name == "tx10p" else 90.0
            pdoy = xcal.select_percentile(xcal.percentile_doy(da, window=5, per=per), per)
            got = fn(da, pdoy).values
            tab = O.percentile_doy(x, da.time.year, da.time.doy, 5, per)[:, 0]
            exp = O.doy_threshold_count(x, tab, da.time.doy, poff, "<" if name == "tx10p" else ">")
        else:
            got = fn(da).values
            exp = _oracle(name, x, poff, da.time, data)
        exp = np.asarray(exp)
        assert got.shape == exp.shape, name
        if np.issubdtype(got.dtype, np.integer) or name.startswith(("cold_spell", "hot_spell_f", "hot_spell_m", "hot_spell_t",
                                                                   "heat_wave", "frost_free", "maximum_consecutive",
                                                                   "dry_spell", "wet_spell")) and name != "hot_spell_max_magnitude":
            np.testing.assert_array_equal(got, exp.astype(got.dtype), err_msg=name)       # integer-valued: exact
        else:
            np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-6, equal_nan=True, err_msg=name)  # float: 1e-5
        checked += 1
    assert checked == 50


def test_device_outputs_option_matches_host_outputs(cuda):
    """set_options(device_outputs=True): device-resident Fields in, device-resident Fields out, the
    same values and dtypes as the host-materialised results."""
    import torch
    import xclim_b200
    from xclim_b200 import Field, indices, run_length
    data = _inputs()
    units = {"tas": "K", "tasmax": "K", "tasmin": "K", "pr": "mm/d"}
    host = {k: make_field(v, "1981-01-01", calendar="noleap", units=units[k]) for k, v in data.items()}
    dev = {k: Field(torch.from_numpy(v).cuda(), f.dims, f.time, dict(f.coords), dict(f.attrs)) for (k, v), f in
           zip(data.items(), host.values())}
    names = ["tg_mean", "wetdays", "maximum_consecutive_dry_days", "dry_spell_frequency", "growing_degree_days",
             "max_n_day_precipitation_amount", "hot_spell_max_magnitude", "cold_spell_days"]
    var_of = dict(indices.BATCH_INDICATORS)
    for name in names:
        fn = getattr(indices, name)
        ref = fn(host[var_of[name]])
        with xclim_b200.set_options(device_outputs=True):
            got = fn(dev[var_of[name]])
        assert isinstance(got, Field) and got.values.is_cuda, name
        assert got.numpy().dtype == ref.values.dtype, name
        np.testing.assert_array_equal(got.numpy(), ref.values, err_msg=name)
        assert got.attrs == ref.attrs
    # first_run with coord="dayofyear" does its index -> doy lookup on the device
    mask = host["pr"].assign_attrs()
    ref = run_length.first_run(Field((data["pr"] > 5).astype(np.float32), mask.dims, mask.time), 2, freq="YS",
                               coord="dayofyear")
    with xclim_b200.set_options(device_outputs=True):
        got = run_length.first_run(Field(torch.from_numpy((data["pr"] > 5).astype(np.float32)).cuda(), mask.dims,
                                         mask.time), 2, freq="YS", coord="dayofyear")
    np.testing.assert_array_equal(got.numpy(), ref.values)
    assert not xclim_b200.options.OPTIONS["device_outputs"]
