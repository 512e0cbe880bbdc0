"""GPU tests added after the round's last GPU run (no GPU minutes were left to execute them): they
exercise kernels that were already green through new entry points / known answers.  Kept in a file that
sorts last so that a surprise here cannot mask the validated suite under `pytest -x`."""
import numpy as np
import pytest

from oracle import xclim_oracle as O
from xb_helpers import make_field

pytestmark = pytest.mark.gpu


def test_more_index_entry_points_on_device(cuda):
    """The same assertions as tests/test_host_layer_cpu.py::test_more_index_entry_points, through the
    real kernels instead of the oracle stand-ins."""
    import test_host_layer_cpu as cpu_side
    cpu_side.test_more_index_entry_points(None)
    cpu_side.test_days_over_precip_thresh(None)


@pytest.mark.parametrize("n_src,doy_min,doy_max", [(360, 1, 366), (366, 1, 360), (365, 1, 366), (92, 153, 244)])
def test_doy_interp_reference_known_answers(cuda, n_src, doy_min, doy_max):
    """tests/test_calendar.py:142-200: the re-mapped table keeps its end points; everything in between
    follows numpy's interp on linspace(doy_min, doy_max, n_src) (core/calendar.py:720-722)."""
    import torch
    from xclim_b200 import device
    rng = np.random.default_rng(17)
    tab = np.arange(n_src, dtype=np.float64)[:, None] + rng.standard_normal((n_src, 7)).cumsum(0)
    tab[:, 0] = np.arange(n_src)
    got = device.doy_interp(torch.from_numpy(tab).cuda(), doy_min, doy_max).cpu().numpy()
    exp = O.interpolate_doy_calendar(tab, doy_max, doy_min)
    assert got.shape == exp.shape == (doy_max - doy_min + 1, 7)
    np.testing.assert_array_equal(got[0], tab[0])
    np.testing.assert_array_equal(got[-1], tab[-1])
    np.testing.assert_allclose(got, exp, rtol=1e-12, atol=1e-12)


def test_spell_min_gap_reference_known_answer(cuda):
    """tests/test_run_length.py:150-162 through the product path: gaps of 1 and 2 steps are bridged by
    min_gap=3, the 4-step gap is kept -> spells of 10 and 5 steps."""
    from xclim_b200 import generic
    v = np.zeros(365, np.float32)
    a = [0, 1, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0]
    v[:len(a)] = a
    da = make_field(v, "2000-01-01", calendar="noleap", units="")
    mx, total, n = generic.spell_length_statistics(da, 0.5, 1, None, ">", ["max", "sum", "count"], "YS", min_gap=3)
    assert mx.values[0] == 10 and total.values[0] == 15 and n.values[0] == 2
    mx1, n1 = generic.spell_length_statistics(da, 0.5, 1, None, ">", ["max", "count"], "YS")
    assert mx1.values[0] == 5 and n1.values[0] == 4


@pytest.mark.parametrize("op,expected", [(">", 6), (">=", 5), ("==", 5), ("!=", 1), ("lt", 5), ("le", 4), ("eq", 4),
                                         ("ne", 1)])
def test_first_day_threshold_reached_reference_known_answers(cuda, op, expected):
    """tests/test_generic.py:343-383: pr = 0, .001, ..., .007 (flipped for the '<' family), threshold
    0.004 kg m-2 s-1, after 01-01, window 1."""
    from xclim_b200 import generic
    a = np.zeros(365, np.float32)
    a[:8] = (np.arange(8) / 1000).astype(np.float32)
    if op in ("lt", "le", "eq", "ne"):
        a[:8] = a[:8][::-1].copy()
    pr = make_field(a, "2000-01-01", calendar="noleap", units="kg m-2 s-1")
    out = generic.first_day_threshold_reached(pr, threshold="0.004 kg m-2 s-1", op=op, after_date="01-01", window=1,
                                              freq="YS")
    assert out.values[0] == expected
    with pytest.raises(ValueError):
        generic.first_day_threshold_reached(pr, threshold="0.004 kg m-2 s-1", op=">", after_date="01-01", window=1,
                                            freq="YS", constrain=("<", "<="))


def test_atmos_generic_missing_any_wrapper_on_device(cuda):
    """tests/test_host_layer_cpu.py::test_atmos_generic_missing_any_wrapper through the real kernels."""
    import test_host_layer_cpu as cpu_side
    cpu_side.test_atmos_generic_missing_any_wrapper(None)


def test_atmos_check_missing_options_on_device(cuda):
    """tests/test_host_layer_cpu.py::test_atmos_check_missing_options through the real kernels."""
    import test_host_layer_cpu as cpu_side
    cpu_side.test_atmos_check_missing_options(None)


def test_atmos_fused_entry_points_honour_check_missing_on_device(cuda):
    import test_host_layer_cpu as cpu_side
    cpu_side.test_atmos_fused_entry_points_honour_check_missing(None)


def test_bootstrap_converts_table_units_on_device(cuda):
    import test_host_layer_cpu as cpu_side
    cpu_side.test_bootstrap_converts_table_units(None)


def test_reference_default_call_shapes_on_device(cuda):
    """freq=None, index="last", array thresholds, tuple reducers through the real kernels."""
    import test_host_layer_cpu as cpu_side
    cpu_side._check_reference_default_call_shapes()


def test_spell_statistics_with_indexers_on_device(cuda):
    """select_time on the spell mask (xc_spell_mask_f32) through the real kernels."""
    import test_host_layer_cpu as cpu_side
    cpu_side._check_spell_statistics_with_indexers()


def test_rolling_with_indexers_on_device(cuda):
    """select_rolling_resample_op(**indexer) (xc_rolling_period_reduce_sel_f32) through the real kernel."""
    import test_host_layer_cpu as cpu_side
    cpu_side._check_rolling_with_indexers()
