"""Generate golden fixtures from the REFERENCE's own numpy/numba cores (run in the authoring
container, where /root/reference exists; the GPU box only sees the committed .npz files).

    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/ref_run_length_1d.npz  -- statistics_run_1d / windowed_run_count_1d /
        windowed_run_events_1d / first_run_1d / _cumsum_reset_np on seeded boolean series
        (reference: indices/run_length.py:143-151, 1334-1477)
  tests/golden/ref_quantile.npz       -- calc_perc on seeded float32 samples with NaNs
        (reference: core/utils.py:279-557)
Nothing from the reference is copied into the repo: the functions are exec'ed where they lie
(tests/golden/_ref_extract.py) and only their numeric inputs/outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_extract as ref  # noqa: E402


def run_length_fixture(rng):
    rl = ref.load_run_length()
    n_series, T = 64, 120
    series = np.zeros((n_series, T), dtype=bool)
    for i in range(n_series):
        p = rng.uniform(0.05, 0.95)
        # persistent series: flip state with small probability so that long runs occur
        flips = rng.random(T) < rng.uniform(0.05, 0.5)
        state = rng.random() < p
        for t in range(T):
            if flips[t]:
                state = rng.random() < p
            series[i, t] = state
    series[0] = False
    series[1] = True
    series[2, :] = True
    series[2, 35] = False
    reducers = ["max", "min", "sum", "count", "mean", "std"]
    windows = [1, 2, 3, 5, 10]
    stats = np.zeros((len(reducers), len(windows), n_series), dtype=np.float64)
    wcount = np.zeros((len(windows), n_series), dtype=np.float64)
    wevents = np.zeros((len(windows), n_series), dtype=np.float64)
    first = np.zeros((len(windows), n_series), dtype=np.float64)
    for i in range(n_series):
        for wi, w in enumerate(windows):
            for ri, r in enumerate(reducers):
                stats[ri, wi, i] = rl["statistics_run_1d"](series[i], r, w)
            wcount[wi, i] = rl["windowed_run_count_1d"](series[i], w)
            wevents[wi, i] = rl["windowed_run_events_1d"](series[i], w)
            first[wi, i] = rl["first_run_1d"](series[i], w)
    cs_last = rl["_cumsum_reset_np"](series.astype(np.uint8).copy(), "last", np.uint8(1))
    cs_first = rl["_cumsum_reset_np"](series.astype(np.uint8).copy(), "first", np.uint8(1))
    np.savez_compressed(os.path.join(HERE, "ref_run_length_1d.npz"), series=series,
                        reducers=np.array(reducers), windows=np.array(windows), stats=stats,
                        wcount=wcount, wevents=wevents, first=first, cs_last=cs_last, cs_first=cs_first)


def quantile_fixture(rng):
    ut = ref.load_utils()
    cases = []
    for n in (1, 2, 3, 5, 7, 10, 30, 75, 148, 150):
        a = (280 + 10 * rng.standard_normal((40, n))).astype(np.float32)
        # NaNs: some rows partially, one row entirely
        for r in range(40):
            k = rng.integers(0, max(1, n // 3) + 1) if r % 3 == 0 else 0
            if k:
                a[r, rng.choice(n, size=k, replace=False)] = np.nan
        a[5, :] = np.nan
        # ties
        if n >= 5:
            a[7, : n // 2] = a[7, 0]
        cases.append(a)
    pers = np.array([10.0, 50.0, 90.0, 99.0, 0.0, 100.0])
    out = {}
    for i, a in enumerate(cases):
        out[f"x{i}"] = a
        for (al, be), tag in (((1 / 3, 1 / 3), "t8"), ((1.0, 1.0), "t7")):
            with np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    q = ut["calc_perc"](a.copy(), list(pers), al, be)
            out[f"q{i}_{tag}"] = q
    out["percentiles"] = pers
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "ref_quantile.npz"), **out)


if __name__ == "__main__":
    if not ref.available():
        raise SystemExit("reference sources not found under " + ref.REF_ROOT)
    rng = np.random.default_rng(20260923)
    run_length_fixture(rng)
    quantile_fixture(rng)
    print("golden fixtures written to", HERE)
