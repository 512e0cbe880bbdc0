"""Date-bounded runs and seasons (indices/run_length.py:891-1331, indices/generic.py:769-853) composed
from the confined-range boundary kernel (``xc_period_boundary_run_range_f32``)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, device


def _ranges(ta, freq, date, mode, window):
    """(range_lo, range_hi) absolute step ranges per period for a date-bounded search."""
    poff = ta.period_offsets(freq)
    t0, t1 = poff[:-1].astype(np.int64), poff[1:].astype(np.int64)
    if date is None:
        return t0.astype(np.int32), t1.astype(np.int32)
    mid = ta.date_index_in_periods(freq, date).astype(np.int64)
    absent = mid < 0
    if mode == "after":          # time >= date            (run_length.py:1238-1243)
        lo, hi = mid, t1
    elif mode == "before_incl":  # time <= date            (:1281-1283)
        lo, hi = t0, mid + 1
    elif mode == "before_excl":  # time < date             (:1194)
        lo, hi = t0, mid
    else:                        # time < date + window - 1 (:1324-1326)
        lo, hi = t0, np.minimum(mid + window - 1, t1)
    lo = np.where(absent, -1, lo)
    return lo.astype(np.int32), hi.astype(np.int32)


def _to_coord(vals, poff, ta, coord):
    """index (relative to the period start, NaN = none) -> dayofyear (lazy_indexing, core/utils.py:202-276)."""
    if not coord:
        return vals
    if coord != "dayofyear":
        raise NotImplementedError("coord must be False or 'dayofyear'")
    from .run_length import index_to_doy
    return index_to_doy(vals.to(torch.float64), poff, ta)


def first_run_after_date(x2d, ta, freq, op_code, thr, window, date, coord="dayofyear"):
    poff = ta.period_offsets(freq)
    lo, hi = _ranges(ta, freq, date, "after", window)
    out = device.period_boundary_run_range(x2d, poff, lo, hi, op_code, thr, window)
    return _to_coord(out, poff, ta, coord)


def last_run_before_date(x2d, ta, freq, op_code, thr, window, date, coord="dayofyear"):
    poff = ta.period_offsets(freq)
    lo, hi = _ranges(ta, freq, date, "before_incl", window)
    out = device.period_boundary_run_range(x2d, poff, lo, hi, op_code, thr, window, last=True)
    return _to_coord(out, poff, ta, coord)


def first_run_before_date(x2d, ta, freq, op_code, thr, window, date, coord="dayofyear"):
    poff = ta.period_offsets(freq)
    lo, hi = _ranges(ta, freq, date, "start", window)
    out = device.period_boundary_run_range(x2d, poff, lo, hi, op_code, thr, window)
    return _to_coord(out, poff, ta, coord)


def run_end_after_date(x2d, ta, freq, op_code, thr, window, date, coord="dayofyear"):
    poff = ta.period_offsets(freq)
    lo_a, hi_a = _ranges(ta, freq, date, "after", window)
    lo_b, hi_b = _ranges(ta, freq, date, "before_excl", window)
    end = device.period_boundary_run_range(x2d, poff, lo_a, hi_a, op_code, thr, window, negate=True)
    beg = device.period_boundary_run_range(x2d, poff, lo_b, hi_b, op_code, thr, window)
    n = torch.from_numpy(np.diff(poff).astype(np.float32)).to(x2d.device)[:, None]
    end = torch.where(torch.isnan(end) & ~torch.isnan(beg), n - 1, end)
    end = torch.where(torch.isnan(beg), torch.full_like(end, float("nan")), end)
    return _to_coord(end, poff, ta, coord)


def season(x2d, ta, freq, op_code, thr, window, mid_date, stat):
    """rl.season per period: start = first run of `window` True (beginning before mid_date + window - 1),
    end = first run of `window` False at/after both the start and mid_date; length per :1072-1077."""
    poff = ta.period_offsets(freq)
    lo_s, hi_s = _ranges(ta, freq, mid_date, "start", window)
    beg = device.period_boundary_run_range(x2d, poff, lo_s, hi_s, op_code, thr, window)
    lo_e, hi_e = _ranges(ta, freq, mid_date, "after", window)
    end = device.period_boundary_run_range(x2d, poff, lo_e, hi_e, op_code, thr, window, negate=True, cell_lo=beg)
    n = torch.from_numpy(np.diff(poff).astype(np.float32)).to(x2d.device)[:, None]
    nan = torch.full_like(beg, float("nan"))
    if stat == "length":
        length = torch.where(torch.isnan(beg), torch.zeros_like(beg), torch.where(torch.isnan(end), n - beg, end - beg))
        if mid_date is not None:  # groups without the date: start is NaN -> 0
            pass
        return length
    end2 = torch.where(torch.isnan(end) & ~torch.isnan(beg), n - 1, end)
    end2 = torch.where(torch.isnan(beg), nan, end2)
    return _to_coord(beg if stat == "start" else end2, poff, ta, "dayofyear")
