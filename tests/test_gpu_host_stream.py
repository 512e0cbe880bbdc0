"""GPU: the host-buffer entry point equals the device-resident path (and the oracle)."""
import numpy as np
import pytest

from oracle import xclim_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("freq_len", [365, 31])
def test_host_stream_matches_device_and_oracle(cuda, freq_len):
    import torch
    from xclim_b200 import _lib, device
    rng = np.random.default_rng(21)
    T, C = freq_len * 7 + 5, 4 * 37
    x = rng.gamma(0.4, 6.0, size=(T, C)).astype(np.float32)
    x[rng.random(x.shape) < 0.4] = 0
    x[rng.random(x.shape) < 0.01] = np.nan
    poff = np.concatenate([np.arange(0, T, freq_len), [T]]).astype(np.int32)  # last period is short
    xh = torch.from_numpy(x).pin_memory()
    for red, w in (("max", 1), ("sum", 3), ("count", 2)):
        out_h, valid_h, _ = device.period_runstat_host(xh, poff, _lib.OPS["<"], 1.0, _lib.RL_REDUCERS[red], w)
        out_d, valid_d = device.period_runstat(xh.cuda(), poff, _lib.OPS["<"], 1.0, _lib.RL_REDUCERS[red], w,
                                               want_valid=True)
        assert torch.equal(out_h, out_d.cpu()) and torch.equal(valid_h, valid_d.cpu())
        exp = O.resample_and_rl(x < 1.0, True, O.rle_statistics, poff=poff, reducer=red, window=w)
        np.testing.assert_array_equal(out_h.numpy(), exp.astype(np.float32))
    with pytest.raises(ValueError, match="workspace too small"):
        device.load().xc_period_runstat_f32_host  # symbol exists
        ws = torch.empty(16, dtype=torch.uint8, device="cuda")
        out = torch.empty((len(poff) - 1, C), dtype=torch.float32)
        _lib.check(_lib.load().xc_period_runstat_f32_host(xh.data_ptr(), T, C, poff.ctypes.data, len(poff) - 1, 1,
                                                         1.0, 0, 0, 1, out.data_ptr(), None, ws.data_ptr(), 16))
