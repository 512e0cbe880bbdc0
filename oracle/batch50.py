"""TEST INFRASTRUCTURE (like everything under oracle/): oracle compositions of the 50 indicators of
``xclim_b200.indices.BATCH_INDICATORS`` (BASELINE.json configs[4]).  Used by tests/test_gpu_batch.py and by
bench.py's sampled-cell check / CPU baseline leg; never by the product path."""
import numpy as np

from . import xclim_oracle as O

K0 = 273.15


def _spell(x, op, thr, red, window, poff, before, clip=None):
    out = O.resample_and_rl(O.compare(x, op, thr), before, O.rle_statistics, poff=poff, reducer=red, window=window)
    if clip is not None:
        out = np.where(out >= clip, out, 0)
    return out


def oracle_indicator(name, x, poff, ta=None, data=None):
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
        "wetdays_prop": lambda: O.threshold_count(x, ">=", 1.0, poff) / np.diff(poff).reshape((-1,) + (1,) * (x.ndim - 1)),
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


