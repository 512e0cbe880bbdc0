"""The DEVICE code of the fire-weather kernel (xclim_b200/csrc/fwi_core.cuh), compiled for the host, against
the reference fixtures and the oracle.  The CUDA kernel adds only the thread-to-cell mapping to this code
(xclim_b200/csrc/fwi.cu); its run on a GPU is tests/test_zzz_gpu_fire.py."""
import numpy as np
import pytest

import fwi_host_build as hb
from oracle import fire_oracle as FO
from test_fire_oracle import case_inputs, check_outputs, golden, mg  # noqa: F401


@pytest.mark.parametrize("name", list(mg.CFFWIS_CASES))
def test_host_build_of_the_kernel_matches_reference_fixture(golden, name):  # noqa: F811
    args, kw, exp = case_inputs(golden, name)
    if kw.get("season_method") is None:
        args = args[:4] + (None,) + args[5:]      # snd is not needed: the library accepts NULL
    got = hb.run(*args, **kw)
    check_outputs(got, exp, name, exact_frac=0.99)


def test_host_build_matches_oracle_on_other_parameters():
    rng = np.random.default_rng(11)
    inp = mg.cffwis_inputs(seed=5, C=16, T=500)
    tc = lambda a: np.ascontiguousarray(a.T)   # noqa: E731
    dc0, dmc0, ffmc0, wpr = mg.cffwis_state(inp, "some")
    base = (tc(inp["tas"]), tc(inp["pr"]), tc(inp["hurs"]), tc(inp["ws"]), tc(inp["snd"]), inp["mth"], inp["lat"])
    cases = [
        dict(season_method="GFWED", temp_condition_days=9, snow_condition_days=12, outputs=["FFMC", "ISI", "season_mask"]),
        dict(season_method="LA08", snow_condition_days=5, temp_condition_days=2, dry_start="CFS", outputs=["DC", "DMC", "BUI", "season_mask"]),
        dict(season_method="GFWED", dry_start="GFWED+SNOW", snow_cover_days=61, snow_min_cover_frac=0.3,
             snow_min_mean_depth=0.02, outputs=["DC", "DMC", "season_mask"]),
        dict(season_method="WF93", overwintering=True, carry_over_fraction=1.0, wetting_efficiency_fraction=0.5,
             dc_start=20, outputs=["DC", "FWI", "DSR", "winter_pr", "season_mask"]),
        dict(season_method=None, ffmc_start=70, dmc_start=10, dc_start=200, outputs=["FWI"]),
    ]
    for kw in cases:
        asked = kw.pop("outputs")
        extra = [o for o in asked if o in ("season_mask", "winter_pr")]
        exp = FO.fire_weather_calc(*base, None, dc0, dmc0, ffmc0, wpr, outputs=FO.complete_indexes(
            [o for o in asked if o not in extra]) + extra, **kw)
        got = hb.run(*base, None, dc0, dmc0, ffmc0, wpr, outputs=asked, **kw)   # the library adds what it needs
        assert set(got) == set(asked)
        check_outputs(got, {k: np.asarray(exp[k]) for k in asked}, str(kw), exact_frac=0.99)
    assert rng is not None


def test_argument_checks_of_the_library():
    inp = {k: (v[:4, :400] if getattr(v, "ndim", 0) == 2 else v[:400] if k == "mth" else v[:4]) for k, v in mg.cffwis_inputs().items()}
    tc = lambda a: np.ascontiguousarray(a.T)   # noqa: E731
    a = (tc(inp["tas"]), tc(inp["pr"]), None, None, None, inp["mth"], inp["lat"], None, None, None, None, None)
    assert set(hb.run(*a, outputs=["DC"])) == {"DC"}
    with pytest.raises(ValueError, match="hurs"):
        hb.run(*a, outputs=["DMC"])
    with pytest.raises(ValueError, match="overwintering"):
        hb.run(*a, outputs=["DC"], overwintering=True)
    with pytest.raises(ValueError, match="snd"):
        hb.run(*a, outputs=["DC"], season_method="LA08")
    with pytest.raises(ValueError, match="1..32"):
        hb.run(*a, outputs=["DC"], season_method="WF93", temp_condition_days=40)


def test_parameter_struct_layouts_agree():
    """The product's ctypes struct (xclim_b200/_lib.py) and the one the host build was driven with lay
    XcFwiParams out identically (144 bytes; the doubles 8-aligned after 17 4-byte fields)."""
    import ctypes

    from xclim_b200 import _lib
    assert ctypes.sizeof(_lib.FwiParams) == ctypes.sizeof(hb.XcFwiParams) == 144
    for (n1, _), (n2, _) in zip(_lib.FwiParams._fields_, hb.XcFwiParams._fields_):
        assert n1 == n2 and getattr(_lib.FwiParams, n1).offset == getattr(hb.XcFwiParams, n2).offset
    assert _lib.FwiParams.snow_min_cover_frac.offset == 72 and _lib.FwiParams.in_scale.offset == 104


def test_elementwise_functions_of_the_host_build(golden):  # noqa: F811
    """ISI / BUI / FWI / DSR of the stored codes reproduce the reference's own arrays; the overwintered DC its
    tests' known answers (tests/test_cffwis.py:126-145)."""
    from test_fire_oracle import assert_index_close
    g = golden
    ws = g["ws"].T
    ffmc, dmc, dc = (g[f"always_on__{k}"].T for k in ("FFMC", "DMC", "DC"))
    assert_index_close(hb.elementwise("ISI", ws, ffmc), g["always_on__ISI"].T, "ISI")
    assert_index_close(hb.elementwise("BUI", dmc, dc), g["always_on__BUI"].T, "BUI")
    assert_index_close(hb.elementwise("FWI", g["always_on__ISI"].T, g["always_on__BUI"].T), g["always_on__FWI"].T, "FWI")
    assert_index_close(hb.elementwise("DSR", g["always_on__FWI"].T), g["always_on__DSR"].T, "DSR")
    for (dcf, wpr, a, b, mn), exp in (((300, 110, 0.75, 0.75, 15), 109.4657), ((300, 110, 1.0, 0.9, 15), 16.35315),
                                      ((100, 50, 0.75, 0.75, 15), 105.176), ((1, 550, 0.75, 0.75, 10), 10)):
        got = hb.elementwise("OWDC", np.array([dcf], np.float32), np.array([wpr], np.float32), (a, b, mn))
        np.testing.assert_allclose(got, exp, rtol=1e-6)
    assert np.isnan(hb.elementwise("OWDC", np.array([np.nan], np.float32), np.array([5.0], np.float32), (0.75, 0.75, 15)))[0]
    assert hb.elementwise("BUI", np.zeros(1, np.float32), np.zeros(1, np.float32))[0] == 0


@pytest.mark.skipif(not __import__("_ref_extract").available(), reason="reference sources not present (GPU box)")
def test_every_mode_combination_against_the_live_reference():
    """All combinations of season method x overwintering x dry start x initial_start_up, each with random
    thresholds / window lengths / previous codes: the oracle AND the host build of the device code against the
    reference's `_fire_weather_calc` executed where it lies."""
    import itertools
    import warnings

    import _ref_extract as ref
    fw = ref.load_cffwis()
    rng = np.random.default_rng(123)
    T = 320
    combos = itertools.product([None, "mask", "WF93", "LA08", "GFWED"], [False, True], [None, "CFS", "GFWED", "GFWED+SNOW"],
                               [True, False])
    n = 0
    for trial, (season, ow, dry, isu) in enumerate(combos):
        if ow and season is None:
            continue
        inp = mg.cffwis_inputs(seed=1000 + trial, C=16, T=T)
        dc0, dmc0, ffmc0, wpr = mg.cffwis_state(inp, [None, "some", "all"][trial % 3])
        over = dict(season_method=season, overwintering=ow, dry_start=dry, initial_start_up=isu,
                    temp_condition_days=int(rng.integers(1, 6)), snow_condition_days=int(rng.integers(1, 6)),
                    snow_cover_days=int(rng.integers(5, 70)), temp_start_thresh=float(rng.uniform(8, 14)),
                    temp_end_thresh=float(rng.uniform(2, 7)), snow_thresh=float(rng.choice([0.01, 0.03])),
                    prec_thresh=float(rng.choice([1.0, 2.5])), dc_dry_factor=int(rng.integers(2, 7)),
                    dmc_dry_factor=int(rng.integers(1, 4)), snow_min_cover_frac=float(rng.uniform(0.3, 0.8)),
                    snow_min_mean_depth=float(rng.uniform(0.02, 0.1)), carry_over_fraction=float(rng.choice([0.5, 0.75, 1.0])),
                    wetting_efficiency_fraction=float(rng.choice([0.5, 0.75, 0.9])))
        outs = ["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "DSR"]
        mask = None
        if season == "mask":                      # a persistent random mask
            mask = np.zeros((16, T), bool)
            st = rng.random(16) < 0.5
            for t in range(T):
                st = np.where(rng.random(16) < 0.03, ~st, st)
                mask[:, t] = st
        elif season is not None:
            outs.append("season_mask")
        if ow:
            outs.append("winter_pr")
        kw = mg.cffwis_params(fw, outputs=outs, **over)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = fw["_fire_weather_calc"](inp["tas"], inp["pr"], inp["hurs"], inp["ws"], inp["snd"], inp["mth"], inp["lat"],
                                           mask, dc0.copy(), dmc0.copy(), ffmc0.copy(), wpr.copy(), **kw)
        exp = {o: (np.asarray(a).T if np.asarray(a).ndim == 2 else np.asarray(a)) for o, a in zip(outs, res)}
        tc = lambda a: np.ascontiguousarray(a.T)   # noqa: E731
        args = (tc(inp["tas"]), tc(inp["pr"]), tc(inp["hurs"]), tc(inp["ws"]), tc(inp["snd"]), inp["mth"], inp["lat"],
                None if mask is None else np.ascontiguousarray(mask.T), dc0, dmc0, ffmc0, wpr)
        label = f"{season},{ow},{dry},{isu}"
        check_outputs(FO.fire_weather_calc(*args, outputs=outs, **over), exp, "oracle:" + label, exact_frac=0.98)
        check_outputs(hb.run(*args, outputs=outs, **over), exp, "host build:" + label, exact_frac=0.98)
        n += 1
    assert n == 72


def test_ring_mean_rounds_as_numpy_mean():
    """The GFWED season and the snow-aware dry start compare float32 window means with thresholds
    (_cffwis.py:661-668, 746-756): np_mean_f32 must round exactly as np.mean does (pairwise summation)."""
    import ctypes
    lib = hb.load()
    lib.fwi_host_np_mean.restype = ctypes.c_float
    rng = np.random.default_rng(4)
    for n in list(range(1, 41)) + [59, 60, 61, 64, 100, 127, 128]:
        for _ in range(6):
            vals = (rng.random(n) * rng.choice([1e-3, 1.0, 1e3], size=n)).astype(np.float32)
            head = int(rng.integers(0, n))
            ring = np.empty(n, np.float32)
            ring[(head + 1 + np.arange(n)) % n] = vals            # oldest value right after the head
            got = lib.fwi_host_np_mean(ring.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(n), ctypes.c_int32(head))
            assert np.float32(got) == np.mean(vals), (n, head)


@pytest.mark.skipif(not __import__("_ref_extract").available(), reason="reference sources not present (GPU box)")
def test_branch_point_inputs_against_the_live_reference():
    """Inputs drawn from the branch points of the equations (rain thresholds 0.5 / 1.5 / 2.8 mm and just above,
    humidity 0 and 100 %, calm and storm winds, temperatures at -2.8 / -1.1 / 21.1 degC, snow depth at its
    threshold, latitudes on the band edges, previous codes at 0 / 33 / 65 / 101), with and without missing
    values: oracle and host build of the device code against the reference executed where it lies."""
    import warnings

    import _ref_extract as ref
    fw = ref.load_cffwis()
    rng = np.random.default_rng(7)
    C, T = 32, 240
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)   # noqa: E731
    for trial in range(6):
        tas = rng.choice([-45.0, -2.8, -1.1, 0.0, 21.1, 35.0, 50.0], size=(C, T)) + rng.normal(0, 0.5, (C, T))
        pr = rng.choice([0.0, 0.5, 0.50001, 1.5, 1.50001, 2.8, 2.80001, 50.0, 300.0], size=(C, T))
        hurs = rng.choice([0.0, 1e-3, 5.0, 50.0, 99.999, 100.0], size=(C, T))
        ws = rng.choice([0.0, 1e-3, 10.0, 100.0, 250.0], size=(C, T))
        snd = rng.choice([0.0, 0.01, 0.010001, 0.5], size=(C, T))
        if trial % 3 == 0:
            for a in (tas, pr, hurs, ws):
                a[rng.random((C, T)) < 0.002] = np.nan
        mth = rng.integers(1, 13, T).astype(np.int64)
        lat = rng.choice([-90, -30, -15, 15, 30, 90, -29.999, 14.999, 0], size=C).astype(np.float64)
        dc0, dmc0, ffmc0, wpr = mg.cffwis_state({"tas": f32(tas)}, [None, "some", "all"][trial % 3])
        if trial % 3:
            dc0[:4], dmc0[:4], ffmc0[:4] = [0.0, 1e-3, 800.0, 2000.0], [0.0, 33.0, 65.0, 500.0], [0.0, 101.0, 50.0, 99.9]
        for season, ow, dry in ((None, False, None), ("WF93", True, "CFS"), ("GFWED", False, "GFWED+SNOW"), ("LA08", True, None)):
            outs = ["DC", "DMC", "FFMC", "ISI", "BUI", "FWI", "DSR"] + (["season_mask"] if season else []) + \
                (["winter_pr"] if ow else [])
            over = dict(season_method=season, overwintering=ow, dry_start=dry, snow_cover_days=10)
            kw = mg.cffwis_params(fw, outputs=outs, **over)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = fw["_fire_weather_calc"](f32(tas), f32(pr), f32(hurs), f32(ws), f32(snd), mth, lat, None, dc0.copy(),
                                               dmc0.copy(), ffmc0.copy(), wpr.copy(), **kw)
            exp = {o: (np.asarray(a).T if np.asarray(a).ndim == 2 else np.asarray(a)) for o, a in zip(outs, res)}
            args = tuple(f32(a.T) for a in (tas, pr, hurs, ws, snd)) + (mth, lat, None, dc0, dmc0, ffmc0, wpr)
            check_outputs(FO.fire_weather_calc(*args, outputs=outs, **over), exp, f"oracle {trial} {season}", exact_frac=0.97)
            check_outputs(hb.run(*args, outputs=outs, **over), exp, f"host build {trial} {season}", exact_frac=0.97)
