"""GPU: parity at the FULL headline size (10950, 721, 1440): sampled cells -- including the cells of the last
CTAs -- of cdd, the tx90p percentile table + counts and the bootstrap against the oracle (all
bit-exact tests ran on grids <= 8 x 16 cells).  Needs ~50 GB of HBM: skipped on smaller devices."""
import numpy as np
import pytest

from oracle import xclim_oracle as O

pytestmark = pytest.mark.gpu

T, Y, X, YEAR = 10950, 721, 1440, 365


def _sample(C, n, seed):
    rng = np.random.default_rng(seed)
    return np.unique(np.concatenate([rng.integers(0, C, size=n), np.arange(C - 256, C), [0, 1, 127, 128]]))


@pytest.fixture(scope="module")
def big(cuda):
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs 60 GB of free HBM")
    return cuda


def test_fullsize_cdd_sampled_cells(big):
    import torch
    from xclim_b200 import _lib, device
    C = Y * X
    assert T * C > 2 ** 31                       # index arithmetic beyond 32 bits is exercised
    poff = np.arange(T // YEAR + 1, dtype=np.int32) * YEAR
    pr = device.synth(T, C, kind=0, seed=2)
    out, valid = device.period_runstat(pr, poff, _lib.OPS["<"], 1.0, _lib.RL_REDUCERS["max"], 1, True, want_valid=True)
    out_a, _ = device.period_runstat(pr, poff, _lib.OPS["<"], 1.0, _lib.RL_REDUCERS["max"], 1, False)
    sel = _sample(C, 4096, 5)
    idx = torch.from_numpy(sel).cuda()
    xs = pr[:, idx].cpu().numpy()
    np.testing.assert_array_equal(out[:, idx].cpu().numpy(), O.maximum_consecutive_dry_days(xs, 1.0, poff))
    np.testing.assert_array_equal(valid[:, idx].cpu().numpy() != YEAR, O.missing_any(xs, poff))
    exp_after = O.resample_and_rl(O.compare(xs, "<", 1.0), False, O.rle_statistics, poff=poff, reducer="max", window=1)
    np.testing.assert_array_equal(out_a[:, idx].cpu().numpy(), exp_after)
    del pr, out, valid, out_a
    torch.cuda.empty_cache()


def test_fullsize_tx90p_and_bootstrap_sampled_cells(big):
    import torch
    from xclim_b200 import _lib, device
    C = Y * X
    N = T // YEAR
    poff = np.arange(N + 1, dtype=np.int32) * YEAR
    doy = (np.arange(T) % YEAR + 1).astype(np.int16)
    yidx = (np.arange(T) // YEAR).astype(np.int16)
    tasmax = device.synth(T, C, kind=1, seed=3)
    table = device.percentile_doy(tasmax, doy, yidx, YEAR, N, 5, [90.0, 10.0], 1 / 3, 1 / 3)
    cnt, valid = device.doy_threshold_count(tasmax, poff, doy, table[0], _lib.OPS[">"], want_valid=True)
    sel = _sample(C, 768, 6)
    idx = torch.from_numpy(sel).cuda()
    xs = tasmax[:, idx].cpu().numpy()
    tab_o = O.percentile_doy(xs, yidx.astype(np.int64), doy.astype(np.int64), 5, [90.0, 10.0])
    np.testing.assert_array_equal(table[:, :, idx].cpu().numpy(), np.moveaxis(tab_o, 1, 0))
    cnt_o = O.doy_threshold_count(xs, tab_o[:, 0], doy.astype(np.int64), poff, ">")
    np.testing.assert_array_equal(cnt[:, idx].cpu().numpy(), cnt_o)
    del table, cnt, valid
    # bootstrap (15-year base) on a handful of cells incl. the very last one
    nb = 15
    step_period = np.repeat(np.arange(nb), YEAR).astype(np.int32)
    boot = device.bootstrap_doy_count(tasmax, 0, nb, YEAR, step_period, N, 5, 90.0, 1 / 3, 1 / 3, _lib.OPS[">"])
    sb = np.array([0, 129, C // 2 + 3, C - 130, C - 1])
    ib = torch.from_numpy(sb).cuda()
    exp = O.bootstrap_doy_count(tasmax[:, ib].cpu().numpy(), (np.arange(T) // YEAR + 1981).astype(np.int64),
                                doy.astype(np.int64), poff, (1981, 1981 + nb - 1), window=5, per=90.0, op=">")
    np.testing.assert_array_equal(boot[:nb, ib].cpu().numpy(), exp[:nb])
    del tasmax, boot
    torch.cuda.empty_cache()
