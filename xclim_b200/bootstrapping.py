"""B200 implementation behind ``xclim.core.bootstrapping`` (percentile bootstrap, Zhang 2005)."""
from __future__ import annotations

import numpy as np

from . import _lib, device
from .calendar import adjust_table, table_on_device
from .field import attrs_of
from .generic import _unwrap, _wrap_periods


def bootstrap_doy_count(da, per, freq, op, constrain):
    """Body of ``percentile_bootstrap`` for the doy-percentile day counts (tx90p & family) --
    core/bootstrapping.py:81-211.  Error behaviour follows the reference: ``KeyError`` when the
    percentile array was not made by ``percentile_doy`` (:131-143) or when the base period covers
    all / none of the studied period (:159-168)."""
    code = _lib.op_code(op, constrain)
    pattrs = attrs_of(per)
    if "percentile_doy" not in pattrs.get("history", ""):
        raise KeyError("`bootstrap` can only be used with percentiles computed using `percentile_doy`")
    clim = pattrs["climatology_bounds"]
    window, alpha, beta = int(pattrs["window"]), float(pattrs["alpha"]), float(pattrs["beta"])
    pcs = np.atleast_1d(np.asarray(per.coords["percentiles"] if "percentiles" in getattr(per, "coords", {}) else
                                   pattrs.get("percentiles")))
    if pcs.size != 1:
        raise ValueError("select one percentile first: per.sel(percentiles=p)")
    percentile = float(pcs[0])
    x2d, cell_shape, other, ta = _unwrap(da)
    y0, y1 = int(str(clim[0])[:4]), int(str(clim[1])[:4])
    sl = ta.sel_years(y0, y1)                                 # da.sel(time=slice(*clim))  (:158)
    n_over = sl.stop - sl.start
    if n_over == len(ta):
        raise KeyError("`bootstrap` is unnecessary when all years are overlapping between reference "
                       "(percentiles period) and studied (index period) periods")
    if n_over == 0:
        raise KeyError("`bootstrap` is unnecessary when no year overlap between reference "
                       "(percentiles period) and studied (index period) periods.")
    gid = ta.bootstrap_group_ids(freq)                        # year groups (:175, 214-223)
    base_gid = gid[sl]
    groups, starts, lens = np.unique(base_gid, return_index=True, return_counts=True)
    if len(set(lens.tolist())) != 1:
        raise NotImplementedError("bootstrap on year blocks of unequal length (365 <-> 366 conversion, "
                                  "core/bootstrapping.py:266-269) is not supported by the B200 hot path")
    L, N = int(lens[0]), len(groups)
    if N < 2:
        raise KeyError("`bootstrap` needs at least two years in the reference period")
    poff = ta.period_offsets(freq)
    P = len(poff) - 1
    pidx = np.repeat(np.arange(P), np.diff(poff))             # period of every step
    step_period = pidx[sl].astype(np.int32)
    boot = device.bootstrap_doy_count(x2d, sl.start, N, L, step_period, P, window, percentile, alpha, beta, code)
    # periods outside the base: plain count against the original table (:205-207)
    table = table_on_device(per, cell_shape, other, x2d.device)
    table, doy_idx = adjust_table(table, ta)
    plain, _ = device.doy_threshold_count(x2d, poff, doy_idx, table, code)
    in_base = np.zeros(P, bool)
    in_base[np.unique(step_period)] = True
    import torch
    mask = torch.from_numpy(in_base).to(x2d.device)[:, None]
    out = torch.where(mask, boot, plain.to(torch.float64))
    attrs = attrs_of(da)
    attrs["units"] = "d"
    return _wrap_periods(da, out, cell_shape, other, ta, freq, attrs, dtype=np.float64)
