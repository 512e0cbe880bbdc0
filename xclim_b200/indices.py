"""Index entry points (L3 of the reference) for the hot-path configurations.

Same names / arguments / units as ``xclim.indices`` so that they can be bound to the existing
``Indicator`` registry through ``compute=`` (core/indicator.py:471-518, 884-886).
"""
from __future__ import annotations

from . import generic
from .units import threshold_in_units_of


def maximum_consecutive_dry_days(pr, thresh="1 mm/day", op="<", freq="YS", resample_before_rl=True):
    """Longest spell with precipitation under a threshold -- indices/_threshold.py:2895-2937."""
    thr = threshold_in_units_of(thresh, pr)
    return generic.spell_length_statistics(pr, thr, 1, None, op, "max", freq, resample_before_rl=resample_before_rl)


def maximum_consecutive_wet_days(pr, thresh="1 mm/day", freq="YS", resample_before_rl=True):
    """indices/_threshold.py:799-841 (``op`` is fixed to ">=")."""
    thr = threshold_in_units_of(thresh, pr)
    return generic.spell_length_statistics(pr, thr, 1, None, ">=", "max", freq, resample_before_rl=resample_before_rl)


def tg_mean(tas, freq="YS"):
    """Mean of daily mean temperature -- indices/_simple.py:76-113."""
    return generic.select_resample_op(tas, op="mean", freq=freq)


def wetdays(pr, thresh="1.0 mm/day", freq="YS", op=">="):
    """indices/_threshold.py:2749-2789."""
    thr = threshold_in_units_of(thresh, pr)
    out = generic.threshold_count(pr, op, thr, freq, constrain=(">", ">="))
    return out.assign_attrs(units="d")


def dry_days(pr, thresh="0.2 mm/d", freq="YS", op="<"):
    """indices/_threshold.py:756-796."""
    thr = threshold_in_units_of(thresh, pr)
    out = generic.threshold_count(pr, op, thr, freq, constrain=("<", "<="))
    return out.assign_attrs(units="d")


# ------------------------------------------------------------------ percentile-threshold day counts
def _percentile_day_count(da, per, freq, bootstrap, op, constrain):
    """Shared body of tx90p & family -- indices/_multivariate.py:1583-1590:
    ``thresh = resample_doy(per, da); threshold_count(da, op, thresh, freq)``; the (lat, lon, time)
    float64 threshold array of the reference is never built: the kernel indexes the per-doy table."""
    import numpy as np

    from . import _lib, device
    from .calendar import adjust_table, table_on_device
    from .generic import _unwrap, _wrap_periods
    from .field import attrs_of

    code = _lib.op_code(op, constrain)
    if bootstrap:
        from .bootstrapping import bootstrap_doy_count
        return bootstrap_doy_count(da, per, freq, op, constrain)
    x2d, cell_shape, other, ta = _unwrap(da)
    table = table_on_device(per, cell_shape, other, x2d.device)
    table, doy_idx = adjust_table(table, ta)
    out, _ = device.doy_threshold_count(x2d, ta.period_offsets(freq), doy_idx, table, code)
    attrs = attrs_of(da)
    attrs["units"] = "d"
    return _wrap_periods(da, out, cell_shape, other, ta, freq, attrs, dtype=np.int64)


def tx90p(tasmax, tasmax_per, freq="YS", bootstrap=False, op=">"):
    """Days with daily maximum temperature over the 90th percentile -- indices/_multivariate.py:1534-1590."""
    return _percentile_day_count(tasmax, tasmax_per, freq, bootstrap, op, (">", ">="))


def tx10p(tasmax, tasmax_per, freq="YS", bootstrap=False, op="<"):
    """indices/_multivariate.py:1593-1650."""
    return _percentile_day_count(tasmax, tasmax_per, freq, bootstrap, op, ("<", "<="))


def tn90p(tasmin, tasmin_per, freq="YS", bootstrap=False, op=">"):
    """indices/_multivariate.py:1417-1473."""
    return _percentile_day_count(tasmin, tasmin_per, freq, bootstrap, op, (">", ">="))


def tn10p(tasmin, tasmin_per, freq="YS", bootstrap=False, op="<"):
    """indices/_multivariate.py:1476-1531."""
    return _percentile_day_count(tasmin, tasmin_per, freq, bootstrap, op, ("<", "<="))


def tg90p(tas, tas_per, freq="YS", bootstrap=False, op=">"):
    """indices/_multivariate.py:1300-1356."""
    return _percentile_day_count(tas, tas_per, freq, bootstrap, op, (">", ">="))


def tg10p(tas, tas_per, freq="YS", bootstrap=False, op="<"):
    """indices/_multivariate.py:1359-1414."""
    return _percentile_day_count(tas, tas_per, freq, bootstrap, op, ("<", "<="))


# ------------------------------------------------------------------ rolling / spell families
def max_n_day_precipitation_amount(pr, window=1, freq="YS"):
    """Highest precipitation amount cumulated over an n-day moving window -- indices/_simple.py:485-525:
    ``rate2amount(pr).rolling(time=window).sum(skipna=False).resample(time=freq).max()``.  The input
    is expected in mm/d (rate2amount is then the identity factor 1 d, core/units.py:853-937)."""
    from . import _lib, device
    from .field import attrs_of
    from .generic import _unwrap, _wrap_periods
    x2d, cell_shape, other, ta = _unwrap(pr)
    out = device.rolling_period_reduce(x2d, ta.period_offsets(freq), window, _lib.STATS["sum"], False,
                                       _lib.STATS["max"])
    attrs = attrs_of(pr)
    attrs["units"] = "mm"
    return _wrap_periods(pr, out, cell_shape, other, ta, freq, attrs)


def _dry_wet_spell(pr, thresh, window, op, win_reducer, spell_reducer, freq, resample_before_rl):
    thr = threshold_in_units_of(thresh, pr)
    return generic.spell_length_statistics(pr, thr, window, win_reducer, op, spell_reducer, freq,
                                           resample_before_rl=resample_before_rl)


def dry_spell_frequency(pr, thresh="1.0 mm", window=3, freq="YS", resample_before_rl=True, op="sum"):
    """indices/_threshold.py:3314-3382 (input in mm/d so that the daily amount equals the rate)."""
    return _dry_wet_spell(pr, thresh.replace(" mm", " mm/d") if isinstance(thresh, str) else thresh, window, "<", op,
                          "count", freq, resample_before_rl)


def dry_spell_total_length(pr, thresh="1.0 mm", window=3, op="sum", freq="YS", resample_before_rl=True):
    """indices/_threshold.py:3385-3454."""
    return _dry_wet_spell(pr, thresh.replace(" mm", " mm/d") if isinstance(thresh, str) else thresh, window, "<", op,
                          "sum", freq, resample_before_rl)


def dry_spell_max_length(pr, thresh="1.0 mm", window=1, op="sum", freq="YS", resample_before_rl=True):
    """indices/_threshold.py:3457-3522."""
    return _dry_wet_spell(pr, thresh.replace(" mm", " mm/d") if isinstance(thresh, str) else thresh, window, "<", op,
                          "max", freq, resample_before_rl)


def wet_spell_frequency(pr, thresh="1.0 mm", window=3, freq="YS", resample_before_rl=True, op="sum"):
    """indices/_threshold.py:3525-3592."""
    return _dry_wet_spell(pr, thresh.replace(" mm", " mm/d") if isinstance(thresh, str) else thresh, window, ">=", op,
                          "count", freq, resample_before_rl)


def hot_spell_max_magnitude(tasmax, thresh="25.0 degC", window=3, freq="YS", resample_before_rl=True):
    """Largest cumulated exceedance of a hot spell -- indices/_threshold.py:2019-2073; the
    ``(tasmax - thresh).clip(0)`` array of the reference is never built."""
    from . import _lib, device
    from .field import attrs_of
    from .generic import _unwrap, _wrap_periods
    thr = threshold_in_units_of(thresh, tasmax)
    x2d, cell_shape, other, ta = _unwrap(tasmax)
    out = device.period_run_maxsum(x2d, ta.period_offsets(freq), _lib.OPS[">"], thr, window, resample_before_rl)
    attrs = attrs_of(tasmax)
    attrs["units"] = "K d"
    return _wrap_periods(tasmax, out, cell_shape, other, ta, freq, attrs)
