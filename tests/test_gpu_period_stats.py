"""GPU parity: streaming period kernels (through the C ABI) vs the CPU oracle.  Bit-exact for
integer-valued outputs; 1e-5 relative for float reductions (tolerance stated per test)."""
import numpy as np
import pytest

from oracle import xclim_oracle as O
from xb_helpers import make_field

pytestmark = pytest.mark.gpu

OPS = [">", "<", ">=", "<=", "==", "!="]


def _data(rng, T, shape, nan_frac=0.02, quant=True):
    x = rng.gamma(0.4, 6.0, size=(T,) + shape).astype(np.float32)
    x[rng.random(x.shape) < 0.45] = 0.0
    if quant:
        x = np.round(x * 2) / 2  # many exact ties with the thresholds
    x[rng.random(x.shape) < nan_frac] = np.nan
    return x.astype(np.float32)


@pytest.mark.parametrize("shape", [(8, 16), (5, 7), (3,)])
@pytest.mark.parametrize("freq", ["YS", "MS", "QS-DEC"])
def test_threshold_count_all_ops(cuda, shape, freq):
    from xclim_b200 import generic
    rng = np.random.default_rng(1)
    x = _data(rng, 365 * 3 + 1, shape)
    da = make_field(x, "2000-01-01", units="mm/d")
    poff = da.time.period_offsets(freq)
    for op in OPS[:4]:
        got = generic.threshold_count(da, op, 1.0, freq)
        exp = O.threshold_count(x, op, 1.0, poff)
        assert got.values.dtype == np.int64
        np.testing.assert_array_equal(got.values, exp, err_msg=f"{op} {freq}")
    with pytest.raises(ValueError, match="not permitted"):
        generic.threshold_count(da, "==", 1.0, freq)


def test_threshold_f32_vs_f64_semantics(cuda):
    """numpy>=2: float32 data vs Python float compares in float32; vs np.float64 array in float64.
    Data equal to float32(thr) separates the two (SURVEY.md section 7 'exact comparison semantics')."""
    from xclim_b200 import device, _lib
    import torch
    thr = 1 / 86400  # not float32-representable
    t32 = np.float32(thr)
    vals = np.array([t32, np.nextafter(t32, np.float32(0)), np.nextafter(t32, np.float32(1)), 0.0, np.nan],
                    dtype=np.float32)
    x = np.tile(vals, 40)[:, None].repeat(4, axis=1)
    xd = torch.from_numpy(x).cuda()
    poff = np.array([0, 100, 200], np.int32)
    for op in OPS:
        code = _lib.OPS[op]
        c32, _ = device.period_count(xd, poff, code, thr, cmp_f64=False)
        c64, _ = device.period_count(xd, poff, code, thr, cmp_f64=True)
        e32 = np.stack([O.compare(x[s:e], op, float(thr)).sum(0) for s, e in ((0, 100), (100, 200))])
        e64 = np.stack([O.compare(x[s:e].astype(np.float64), op, thr).sum(0) for s, e in ((0, 100), (100, 200))])
        np.testing.assert_array_equal(c32.cpu().numpy(), e32, err_msg=f"f32 {op}")
        np.testing.assert_array_equal(c64.cpu().numpy(), e64, err_msg=f"f64 {op}")
    assert not np.array_equal(
        device.period_count(xd, poff, _lib.OPS["<"], thr, cmp_f64=False)[0].cpu().numpy(),
        device.period_count(xd, poff, _lib.OPS["<"], thr, cmp_f64=True)[0].cpu().numpy())


@pytest.mark.parametrize("shape", [(8, 16), (5, 7)])
@pytest.mark.parametrize("before", [True, False])
@pytest.mark.parametrize("freq", ["YS", "MS"])
def test_spell_length_statistics_window1(cuda, shape, before, freq):
    from xclim_b200 import generic
    rng = np.random.default_rng(2)
    x = _data(rng, 365 * 2 + 40, shape, nan_frac=0.01)
    da = make_field(x, "2001-01-01", units="mm/d")
    poff = da.time.period_offsets(freq)
    for op in ("<", ">="):
        for red in ("max", "sum", "count", "min", "mean", "std"):
            got = generic.spell_length_statistics(da, 1.0, 1, None, op, red, freq, resample_before_rl=before)
            exp = O.spell_length_statistics(x, 1.0, 1, None, op, red, poff, resample_before_rl=before)
            assert got.values.dtype == np.float32
            if red in ("mean", "std"):
                np.testing.assert_allclose(got.values, exp, rtol=1e-5, atol=0)  # float statistic: 1e-5 relative
            else:
                np.testing.assert_array_equal(got.values, exp, err_msg=f"{op} {red} before={before} {freq}")


@pytest.mark.parametrize("window", [2, 3, 6])
@pytest.mark.parametrize("before", [True, False])
def test_run_length_module_masks(cuda, window, before):
    """rl.* entry points on a precomputed boolean mask (runs of True, length >= window)."""
    from xclim_b200 import run_length as rl
    rng = np.random.default_rng(3)
    m = rng.random((800, 6, 8)) < 0.7
    da = make_field(m, "2000-01-01", units="")
    poff = da.time.period_offsets("MS")
    for fn, red in ((rl.windowed_run_count, "sum"), (rl.windowed_run_events, "count")):
        got = rl.resample_and_rl(da, before, fn, window=window, freq="MS")
        exp = O.resample_and_rl(m, before, O.rle_statistics, poff=poff, reducer=red, window=window)
        np.testing.assert_array_equal(got.values, exp.astype(np.float32))
    got = rl.rle_statistics(da, "max", window, freq="MS")
    np.testing.assert_array_equal(got.values, O.rle_statistics(m, "max", window, poff=poff).astype(np.float32))
    # the two dedicated whole-array formulations of the reference agree with the generic one
    np.testing.assert_array_equal(rl.windowed_run_events(da, window, freq="MS").values,
                                  O.windowed_run_events(m, window, poff=poff))
    np.testing.assert_array_equal(rl.windowed_run_count(da, window, freq="MS").values,
                                  O.windowed_run_count(m, window, poff=poff))
    with pytest.raises(ValueError, match="not implemented for 1d method"):
        rl.rle_statistics(da, "max", 1, freq="MS", ufunc_1dim=True)


def test_cdd_reference_known_answers(cuda):
    """tests/test_indices.py:2354-2381 through the public index function."""
    from xclim_b200 import indices
    a = np.zeros(365, np.float32) + 10; a[5:15] = 0
    pr = make_field(a / 86400, "2000-01-01", units="kg m-2 s-1")
    assert indices.maximum_consecutive_dry_days(pr, freq="ME").values[0] == 10
    a = np.zeros(365, np.float32) + 10; a[:10] = 0
    pr = make_field(a / 86400, "2000-01-01", units="kg m-2 s-1")
    assert indices.maximum_consecutive_dry_days(pr, freq="ME").values[0] == 10
    a = np.zeros(365, np.float32) + 10; a[5:35] = 0
    pr = make_field(a / 86400, "2000-01-01", units="kg m-2 s-1")
    out = indices.maximum_consecutive_dry_days(pr, freq="ME", resample_before_rl=True)
    assert out.values[0] == 26 and out.attrs["units"] == "d" and out.values.dtype == np.float32
    assert indices.maximum_consecutive_dry_days(pr, freq="ME", resample_before_rl=False).values[0] == 30


def test_run_length_reference_known_answers(cuda):
    """tests/test_run_length.py:166-241 (resample before / after)."""
    from xclim_b200 import run_length as rl
    v = np.ones(365); v[35] = 0
    da = make_field(v != 0, "2000-07-01", units="")
    before = rl.resample_and_rl(da, True, rl.rle_statistics, reducer="max", window=1, freq="ME").values
    assert before[0] == 31 and before[1] == 26
    after = rl.rle_statistics(da, "max", 1, freq="ME").values
    assert after[0] == 35 and after[1] == 365 - 35 - 1 and (after[2:] == 0).all()
    da = make_field(np.ones(365, bool), "2000-07-01", units="")
    exp = np.zeros(12); exp[0] = 365
    np.testing.assert_array_equal(rl.rle_statistics(da, "max", 1, freq="ME").values, exp)
    da = make_field(v != 0, "2000-01-01", units="")
    assert rl.resample_and_rl(da, True, rl.rle_statistics, reducer="min", window=1, freq="YS").values[0] == 35
    assert rl.resample_and_rl(da, True, rl.rle_statistics, reducer="mean", window=36, freq="YS").values[0] == 329
    assert rl.resample_and_rl(da, True, rl.rle_statistics, reducer="std", window=1, freq="YS").values[0] == 147


@pytest.mark.parametrize("shape", [(8, 16), (5, 7)])
def test_resample_reductions(cuda, shape):
    from xclim_b200 import generic, indices
    rng = np.random.default_rng(4)
    x = (280 + 8 * rng.standard_normal((365 * 2,) + shape)).astype(np.float32)
    x[rng.random(x.shape) < 0.01] = np.nan
    x[31:59, 0] = np.nan  # an all-NaN month
    da = make_field(x, "2001-01-01", units="K")
    for freq in ("MS", "YS"):
        poff = da.time.period_offsets(freq)
        for op in ("mean", "sum", "min", "max", "std", "var", "count"):
            got = generic.select_resample_op(da, op, freq).values
            exp = O.select_resample_op(x.astype(np.float64), op, poff)
            if op in ("min", "max", "count"):
                np.testing.assert_array_equal(got, exp.astype(got.dtype))
            else:
                # float reductions: 1e-5 relative (float64 accumulation here vs float32 in the reference)
                np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-4 if op in ("std", "var") else 0)
    got = indices.tg_mean(da, freq="MS")
    assert got.attrs["units"] == "K" and got.values.shape == (24,) + shape
    for op in (">", "<"):
        got = generic.cumulative_difference(da, 283.0, op, freq="MS").values
        exp = O.cumulative_difference(x, 283.0, op, da.time.period_offsets("MS"))
        np.testing.assert_allclose(got, exp, rtol=1e-5)


def test_strided_rows_and_large_cells(cuda):
    """ldx != C (a lat tile viewed inside a wider buffer) and C >> one block."""
    import torch
    from xclim_b200 import device, _lib
    rng = np.random.default_rng(5)
    x = _data(rng, 400, (3000,))
    big = torch.from_numpy(np.concatenate([x, x], axis=1)).cuda()
    view = big[:, 1000:2204]  # C=1204, ldx=6000, misaligned start -> VEC=1 path
    poff = np.array([0, 100, 250, 400], np.int32)
    got, valid = device.period_runstat(view, poff, _lib.OPS["<"], 1.0, _lib.RL_REDUCERS["max"], 1, want_valid=True)
    xs = x[:, 1000:2204]
    np.testing.assert_array_equal(got.cpu().numpy(), O.spell_length_statistics(xs, 1.0, 1, None, "<", "max", poff))
    np.testing.assert_array_equal(valid.cpu().numpy() != np.diff(poff)[:, None], O.missing_any(xs, poff))
    view4 = big[:, 1000:2200]  # aligned -> VEC=4 path with ldx != C
    got, _ = device.period_runstat(view4, poff, _lib.OPS["<"], 1.0, _lib.RL_REDUCERS["sum"], 2)
    np.testing.assert_array_equal(got.cpu().numpy(),
                                  O.resample_and_rl((x[:, 1000:2200] < 1.0), True, O.rle_statistics, poff=poff,
                                                    reducer="sum", window=2))


@pytest.mark.parametrize("window", [1, 3])
@pytest.mark.parametrize("before", [True, False])
def test_hot_spell_max_magnitude(cuda, window, before):
    """windowed_max_run_sum of (tasmax - thresh).clip(0): 1e-5 relative (float64 run sums here, float32
    cumulative sums in the reference)."""
    from xclim_b200 import indices, generic
    rng = np.random.default_rng(6)
    x = (295 + 6 * rng.standard_normal((365 * 2 + 30, 6, 8))).astype(np.float32)
    x[rng.random(x.shape) < 0.01] = np.nan
    da = make_field(x, "2001-01-01", units="K")
    for freq in ("YS", "MS"):
        poff = da.time.period_offsets(freq)
        got = indices.hot_spell_max_magnitude(da, "25 degC", window=window, freq=freq, resample_before_rl=before)
        over = np.clip(x - np.float32(298.15), 0, None)
        over = np.where(np.isnan(over), 0, over).astype(np.float64)
        exp = O.resample_and_rl(over, before, O.windowed_max_run_sum, window, poff=poff)
        np.testing.assert_allclose(got.values, exp, rtol=1e-5, atol=1e-6)
    # thresholded statistics / count_occurrences wrappers
    got = generic.thresholded_statistics(da, ">", 298.15, "mean", "YS").values
    exp = O.resample_reduce(np.where(x > np.float32(298.15), x, np.nan).astype(np.float64), da.time.period_offsets("YS"), "mean")
    np.testing.assert_allclose(got, exp, rtol=1e-5)
    np.testing.assert_array_equal(generic.count_occurrences(da, 298.15, "YS", ">").values,
                                  O.threshold_count(x, ">", 298.15, da.time.period_offsets("YS")))


@pytest.mark.parametrize("before", [True, False])
def test_bivariate_heat_waves(cuda, before):
    from xclim_b200 import indices, generic
    rng = np.random.default_rng(7)
    T, shape = 365 * 2 + 20, (5, 8)
    season = 10 * np.sin(2 * np.pi * (np.arange(T) - 100) / 365)[:, None, None]
    tn = (291 + season + 3 * rng.standard_normal((T,) + shape)).astype(np.float32)
    tx = (tn + 8 + 2 * rng.standard_normal((T,) + shape)).astype(np.float32)
    tn[rng.random(tn.shape) < 0.01] = np.nan
    tx[rng.random(tx.shape) < 0.01] = np.nan
    dtn, dtx = make_field(tn, "2001-01-01", units="K"), make_field(tx, "2001-01-01", units="K")
    poff = dtn.time.period_offsets("YS")
    cond = O.compare(tn, ">", 22 + 273.15) & O.compare(tx, ">", 30 + 273.15)
    for fn, red in ((indices.heat_wave_frequency, "count"), (indices.heat_wave_max_length, "max"),
                    (indices.heat_wave_total_length, "sum")):
        got = fn(dtn, dtx, window=3, resample_before_rl=before).values
        exp = O.resample_and_rl(cond, before, O.rle_statistics, poff=poff, reducer=red, window=3)
        np.testing.assert_array_equal(got, exp.astype(np.float32), err_msg=fn.__name__)
    np.testing.assert_array_equal(indices.tx_tn_days_above(dtn, dtx).values,
                                  np.stack([cond[s:e].sum(0) for s, e in zip(poff[:-1], poff[1:])]))
    got = generic.bivariate_count_occurrences(data_var1=dtn, data_var2=dtx, threshold_var1=295.15, threshold_var2=303.15,
                                              freq="MS", op_var1=">", op_var2="<", var_reducer="any").values
    cond2 = O.compare(tn, ">", 295.15) | O.compare(tx, "<", 303.15)
    pm = dtn.time.period_offsets("MS")
    np.testing.assert_array_equal(got, np.stack([cond2[s:e].sum(0) for s, e in zip(pm[:-1], pm[1:])]))
    with pytest.raises(ValueError, match="Unsupported value"):
        generic.bivariate_count_occurrences(data_var1=dtn, data_var2=dtx, threshold_var1=1, threshold_var2=1, freq="YS",
                                            op_var1=">", op_var2=">", var_reducer="most")


@pytest.mark.parametrize("indexer", [{"season": "JJA"}, {"month": [1, 12]}, {"doy_bounds": (300, 40)},
                                     {"date_bounds": ("02-25", "03-05")}, {"season": ["DJF", "MAM"]}])
@pytest.mark.parametrize("calendar", ["standard", "noleap"])
def test_select_time_indexers_on_resample_ops(cuda, indexer, calendar):
    """select_resample_op(da, op, freq, **indexer): select_time then reduce (indices/generic.py:110-114)."""
    from xclim_b200 import generic
    rng = np.random.default_rng(8)
    x = (280 + 5 * rng.standard_normal((365 * 3 + 1, 4, 6))).astype(np.float32)
    x[rng.random(x.shape) < 0.01] = np.nan
    da = make_field(x, "2003-01-01", calendar=calendar, units="K")
    ta = da.time
    kw = dict(indexer)
    if "month" in kw:
        kw = {"months": kw["month"]}
    keep = O.select_time_mask(ta.month, ta.day, ta.doy, ta.calendar, **kw)
    np.testing.assert_array_equal(ta.select_mask(**indexer), keep)
    xm = np.where(keep[:, None, None], x, np.nan)
    for op in ("mean", "max", "count", "sum"):
        got = generic.select_resample_op(da, op, "YS", **indexer).values
        exp = O.select_resample_op(xm.astype(np.float64), op, ta.period_offsets("YS"))
        np.testing.assert_allclose(got, exp.astype(got.dtype) if op == "count" else exp, rtol=1e-5, equal_nan=True)
    with pytest.raises(ValueError, match="Only one method"):
        generic.select_resample_op(da, "mean", "YS", season="JJA", month=[1])


def test_run_length_quantile_reducers(cuda):
    """rle_statistics(reducer="q90"/"q10") -- reference known answers 299.6 and 64.4
    (tests/test_run_length.py:264-278) and parity on random masks."""
    from xclim_b200 import run_length as rl
    v = np.ones(365); v[35] = 0
    da = make_field(v != 0, "2000-01-01", units="")
    assert rl.rle_statistics(da, "q90", 1, freq="YS").values[0] == pytest.approx(299.6, rel=1e-6)
    assert rl.rle_statistics(da, "q10", 1, freq="YS").values[0] == pytest.approx(64.4, rel=1e-6)
    rng = np.random.default_rng(9)
    m = rng.random((730, 5, 6)) < 0.6
    da = make_field(m, "2001-01-01", units="")
    for freq in ("YS", "MS"):
        poff = da.time.period_offsets(freq)
        for red in ("q50", "q90", "q25"):
            for w in (1, 2):
                got = rl.rle_statistics(da, red, w, freq=freq).values
                exp = O.rle_statistics(m, red, w, poff=poff)
                np.testing.assert_allclose(got, exp, rtol=1e-6, err_msg=f"{red} w={w} {freq}")
                got = rl.resample_and_rl(da, True, rl.rle_statistics, reducer=red, window=w, freq=freq).values
                exp = O.resample_and_rl(m, True, O.rle_statistics, poff=poff, reducer=red, window=w)
                np.testing.assert_allclose(got, exp, rtol=1e-6)
