"""Generate golden fixtures from the REFERENCE's own numpy/numba cores (run in the authoring
container, where /root/reference exists; the GPU box only sees the committed .npz files).

    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/ref_run_length_1d.npz  -- statistics_run_1d / windowed_run_count_1d /
        windowed_run_events_1d / first_run_1d / _cumsum_reset_np on seeded boolean series
        (reference: indices/run_length.py:143-151, 1334-1477)
  tests/golden/ref_quantile.npz       -- calc_perc on seeded float32 samples with NaNs
        (reference: core/utils.py:279-557)
  tests/golden/ref_cffwis.npz         -- the fire-weather recurrences (_fire_weather_calc with its season,
        overwintering and dry-start modes) on seeded float32 weather series
        (reference: indices/fire/_cffwis.py:161-900)
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


def cffwis_inputs(seed=20260924, C=16, T=800):
    """Seeded float32 weather series (cells, T), time LAST as the reference's ufunc core takes them."""
    rng = np.random.default_rng(seed)
    lat = np.array([-60, -35, -29.9, -20, -15, -10, 0, 10, 14.9, 15, 25, 30, 40, 47.5, 60, 75], dtype=np.float64)[:C]
    t = np.arange(T)
    doy = t % 365
    mth = np.minimum(doy // 30.42, 11).astype(np.int64) + 1
    season = np.cos(2 * np.pi * (doy - 200) / 365)                      # northern summer
    amp = np.where(lat >= 0, 1.0, -1.0)[:, None] * (6 + 0.25 * np.abs(lat))[:, None]
    base = (22 - 0.35 * np.abs(lat))[:, None]
    tas = base + amp * season[None, :] + 3.0 * rng.standard_normal((C, T))
    wet = rng.random((C, T)) < 0.35
    pr = np.where(wet, rng.gamma(0.7, 5.0, size=(C, T)), 0.0)
    hurs = np.clip(55 + 20 * rng.standard_normal((C, T)) + 15 * wet, 5, 100)
    ws = np.abs(12 + 8 * rng.standard_normal((C, T)))
    snd = np.clip(0.25 * (-(tas - 1.0)) / 10 + 0.05 * rng.standard_normal((C, T)), 0, None) * (tas < 4)
    # a missing block and single missing values
    tas[3, 100:104] = np.nan
    pr[5, 300] = np.nan
    hurs[7, 50] = np.nan
    ws[9, 60] = np.nan
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    return dict(tas=f32(tas), pr=f32(pr), hurs=f32(hurs), ws=f32(ws), snd=f32(snd), mth=mth, lat=lat)


#: the modes of _fire_weather_calc the fixture covers (name -> keyword overrides)
CFFWIS_CASES = {
    "always_on": dict(season_method=None, outputs=["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "DSR"]),
    "wf93": dict(season_method="WF93", outputs=["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "season_mask"]),
    "wf93_end4": dict(season_method="WF93", temp_end_thresh=4.0, temp_condition_days=4, outputs=["DC", "season_mask"]),
    "la08_overwinter": dict(season_method="LA08", overwintering=True, outputs=["DC", "season_mask", "winter_pr"],
                            state="some"),
    "mask_cfs": dict(season_method="mask", dry_start="CFS", dmc_dry_factor=5, outputs=["DC", "DMC"], mask_from="wf93"),
    "gfwed_snow": dict(season_method="GFWED", dry_start="GFWED+SNOW", snow_cover_days=20, snow_min_mean_depth=0.05,
                       outputs=["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "season_mask"]),
    "gfwed_dry": dict(season_method="mask", dry_start="GFWED", outputs=["DC", "DMC", "FFMC"], mask_from="wf93"),
    "mask_ow_cfs_cont": dict(season_method="mask", overwintering=True, dry_start="CFS", initial_start_up=False,
                             outputs=["DC", "DMC", "FFMC", "winter_pr"], mask_from="la08_overwinter", state="all"),
    "dc_only_always_on": dict(season_method=None, outputs=["DC"], state="some"),
}


def cffwis_params(fw, **over):
    kw = {k: (v if not isinstance(v, tuple) else v[0]) for k, v in fw["default_params"].items()}
    kw.update(season_method=None, overwintering=False, dry_start=None, initial_start_up=True)
    kw.update({k: v for k, v in over.items() if k not in ("mask_from", "state")})
    return kw


def cffwis_state(inp, kind):
    C = inp["tas"].shape[0]
    rng = np.random.default_rng(77)
    nan = np.full(C, np.nan, np.float32)
    if kind is None:
        return nan.copy(), nan.copy(), nan.copy(), np.zeros(C, np.float32)
    dc0 = (50 + 300 * rng.random(C)).astype(np.float32)
    dmc0 = (5 + 40 * rng.random(C)).astype(np.float32)
    ffmc0 = (60 + 30 * rng.random(C)).astype(np.float32)
    wpr = (200 * rng.random(C)).astype(np.float32)
    if kind == "some":
        dc0[::3] = np.nan
        dmc0[1::3] = np.nan
        ffmc0[2::3] = np.nan
    return dc0, dmc0, ffmc0, wpr


def cffwis_fixture():
    fw = ref.load_cffwis()
    inp = cffwis_inputs()
    out = {k: v for k, v in inp.items()}
    masks = {}
    for name, over in CFFWIS_CASES.items():
        kw = cffwis_params(fw, **over)
        dc0, dmc0, ffmc0, wpr = cffwis_state(inp, over.get("state"))
        mask = masks[over["mask_from"]] if "mask_from" in over else None
        res = fw["_fire_weather_calc"](inp["tas"], inp["pr"], inp["hurs"], inp["ws"], inp["snd"], inp["mth"], inp["lat"],
                                       mask, dc0, dmc0, ffmc0, wpr, **kw)
        if len(kw["outputs"]) == 1:
            res = (res,)
        for oname, arr in zip(kw["outputs"], res):
            out[f"{name}__{oname}"] = np.asarray(arr)
            if oname == "season_mask":
                masks[name] = np.asarray(arr).astype(bool)
    np.savez_compressed(os.path.join(HERE, "ref_cffwis.npz"), **out)
    return out


if __name__ == "__main__":
    if not ref.available():
        raise SystemExit("reference sources not found under " + ref.REF_ROOT)
    rng = np.random.default_rng(20260923)
    run_length_fixture(rng)
    quantile_fixture(rng)
    cffwis_fixture()
    print("golden fixtures written to", HERE)
