"""Spell statistics for rolling-window spells (window > 1): generic.spell_mask + run statistics."""
from __future__ import annotations

from . import _lib, device


def spell_runstat(x2d, poff, window, win_reducer, op_code, thr, reducer_code, resample_before_rl=True):
    """indices/generic.py:434-585 for ``window > 1``."""
    if win_reducer not in ("min", "max", "sum", "mean"):
        raise ValueError(f"win_reducer must be one of min, max, sum, mean; got {win_reducer!r}")
    return device.spell_runstat(x2d, poff, window, _lib.STATS[win_reducer], op_code, thr, reducer_code,
                                resample_before_rl)
