"""CPU tests of the host layer that need no GPU: units, containers, period bookkeeping, error behaviour."""
import numpy as np
import pytest

from xclim_b200 import Field, TimeAxis
from xclim_b200 import units as U


def test_unit_conversions():
    # thresholds reach the kernels as Python floats (core/units.py:398-403)
    assert U.convert_units_to("1 mm/day", "mm/d") == 1.0
    assert U.convert_units_to("1 mm/day", "kg m-2 s-1") == pytest.approx(1 / 86400)
    assert U.convert_units_to("25 degC", "K") == pytest.approx(298.15)
    assert U.convert_units_to("0 °C", "K") == pytest.approx(273.15)
    assert U.convert_units_to("300 K", "degC") == pytest.approx(26.85)
    assert U.convert_units_to(3.5, "K") == 3.5
    assert isinstance(U.convert_units_to("25 degC", "K"), float)
    with pytest.raises(ValueError):
        U.convert_units_to("1 mm/day", "K")
    assert U.parse_quantity("10.8 m s-1") == (10.8, "m s-1")
    da = Field(np.zeros((3, 2), np.float32), ("time", "x"), TimeAxis.daily("2000-01-01", 3), attrs={"units": "K"})
    assert U.threshold_in_units_of("0 degC", da) == pytest.approx(273.15)
    assert U.to_agg_units_attrs(da, "count") == {"units": "d"} and U.to_agg_units_attrs(da, "mean") == {"units": "K"}


def test_field_container_and_wrapping():
    from xclim_b200.field import wrap_like, time_axis_of, dims_of
    ta = TimeAxis.daily("2001-01-01", 10, "noleap")
    f = Field(np.arange(40, dtype=np.float32).reshape(10, 2, 2), ("time", "lat", "lon"), ta,
              {"lat": np.array([0., 1.]), "lon": np.array([5., 6.])}, {"units": "K"})
    assert f.shape == (10, 2, 2) and dims_of(f) == ("time", "lat", "lon") and time_axis_of(f) is ta
    sub = f.isel_time(slice(2, 5))
    assert sub.shape == (3, 2, 2) and len(sub.time) == 3 and sub.time.doy[0] == 3
    out = wrap_like(f, np.zeros((1, 2, 2)), ("time", "lat", "lon"), time=np.array(["2001-01-01"]), attrs={"units": "d"})
    assert out.attrs["units"] == "d" and (out.coords["lat"] == f.coords["lat"]).all() and "time" in out.coords
    with pytest.raises(TypeError):
        time_axis_of(np.zeros(3))


def test_date_index_and_bootstrap_groups():
    ta = TimeAxis.daily("2000-01-01", 730)
    mids = ta.date_index_in_periods("YS", "07-01")
    assert ta.date_strings(mids) == ["2000-07-01", "2001-07-01"]
    assert (ta.date_index_in_periods("YS-JUL", "01-01") >= 0).tolist() == [True, True, False]
    with pytest.raises(ValueError, match="More than 1 instance"):
        TimeAxis.daily("2000-01-01", 800).date_index_in_periods("2YS", "07-01")
    assert ta.sel_years(2001, 2001) == slice(366, 730)


def test_lat_tiles_cover_grid():
    from xclim_b200.multigpu import lat_tiles, shard_lat
    x = np.arange(721 * 3).reshape(1, 721, 3)
    parts = [shard_lat(x, 1, r, 8) for r in range(8)]
    assert sum(p.shape[1] for p in parts) == 721
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), x)


def test_no_cuda_means_loud_failure():
    """No CPU fallback: without a CUDA device the device layer raises (never silently computes)."""
    import torch
    from xclim_b200 import device, _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.XclimB200Error, match="no CPU fallback"):
        device.to_time_cell(np.zeros((4, 2), np.float32), 0)
    from xclim_b200 import indices
    f = Field(np.zeros((365, 2), np.float32), ("time", "x"), TimeAxis.daily("2001-01-01", 365), attrs={"units": "mm/d"})
    with pytest.raises(_lib.XclimB200Error):
        indices.maximum_consecutive_dry_days(f)


def test_spell_sum_interval_matches_float32_compare():
    """xc_spell_sum_interval (host helper of xc_spell_runstat_f32): the float64-sum interval must
    reproduce `np.float32(s [/ w]) op np.float32(thr)` for every double s, including the doubles
    right next to the rounding boundaries."""
    import ctypes as C
    from xclim_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    ops = {">": np.greater, "<": np.less, ">=": np.greater_equal, "<=": np.less_equal, "==": np.equal,
           "!=": np.not_equal}
    for thr in (1.0, 0.0, 3.3, -5.5, 1e-30, float("inf"), 16777216.0):
        for w in (2, 3, 7):
            for stat in ("sum", "mean"):
                t32 = np.float32(thr)
                centre = float(t32) * (w if stat == "mean" else 1)
                near = np.nextafter(centre, np.inf) if np.isfinite(centre) else centre
                s = [centre, near]
                x = centre
                for _ in range(40):                      # 40 doubles either side of the centre
                    x = np.nextafter(x, -np.inf)
                    s.append(x)
                x = centre
                for _ in range(40):
                    x = np.nextafter(x, np.inf)
                    s.append(x)
                if np.isfinite(centre):                  # the float32 neighbours' midpoints
                    for nb in (np.nextafter(t32, np.float32(-np.inf)), np.nextafter(t32, np.float32(np.inf))):
                        mid = (float(nb) + float(t32)) / 2 * (w if stat == "mean" else 1)
                        s += [mid, np.nextafter(mid, np.inf), np.nextafter(mid, -np.inf)]
                s += list(rng.standard_normal(200) * 10.0) + [np.inf, -np.inf, 0.0, -0.0]
                s = np.asarray(s, dtype=np.float64)
                with np.errstate(all="ignore"):
                    r = (s / w if stat == "mean" else s).astype(np.float32)
                for name, fn in ops.items():
                    lo, hi = C.c_double(), C.c_double()
                    flags = (C.c_int32 * 4)()
                    _lib.check(lib.xc_spell_sum_interval(_lib.op_code(name), float(thr), w, _lib.STATS[stat],
                                                         C.byref(lo), C.byref(hi), flags))
                    with np.errstate(invalid="ignore"):
                        got = ((s >= lo.value) & (bool(flags[0]) | (s < hi.value))) != bool(flags[1])
                    np.testing.assert_array_equal(got, fn(r, t32), err_msg=f"{thr} {w} {stat} {name}")
                    assert bool(flags[2]) == (name == "!=")


def test_bootstrap_replacement_rows_calendar_conversion():
    """core/bootstrapping.py:255-279 as a row map: equal blocks copy, 365 <- 366 drops the source's
    Feb 29, 366 <- 365 leaves a hole (-1 = NaN) on the target's Feb 29."""
    from xb_helpers import make_field
    from xclim_b200.bootstrapping import replacement_rows
    T = 366 + 365 * 3 + 366
    da = make_field(np.zeros((T, 1), np.float32), "2000-01-01", calendar="standard", units="K")
    ta = da.time
    gid = ta.bootstrap_group_ids("YS")
    _, starts, lens = np.unique(gid, return_index=True, return_counts=True)
    assert lens.tolist() == [366, 365, 365, 365, 366]
    r = replacement_rows(ta, starts, lens, 1, 2)
    np.testing.assert_array_equal(r, starts[2] + np.arange(365))
    r = replacement_rows(ta, starts, lens, 1, 4)               # common year <- leap year
    assert len(r) == 365 and (starts[4] + 59) not in r and r[58] == starts[4] + 58 and r[59] == starts[4] + 60
    r = replacement_rows(ta, starts, lens, 0, 2)               # leap year <- common year
    assert len(r) == 366 and r[59] == -1 and r[58] == starts[2] + 58 and r[60] == starts[2] + 59
    assert ta.month[starts[0] + 59] == 2 and ta.day[starts[0] + 59] == 29
