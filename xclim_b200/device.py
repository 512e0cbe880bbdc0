"""Thin device layer: torch tensors for HBM buffers + ctypes calls into ``libxclim_b200.so``.

PyTorch is plumbing here (device memory, streams, pinned host staging); every computation is one
of the hand-written kernels behind the C ABI.  All functions take/return CUDA tensors laid out as
``(time, cell)`` float32 row-major -- the flattened view of the reference's ``(time, lat, lon)``.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from . import _lib
from ._lib import check, load


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.XclimB200Error("xclim_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def current_stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------- staging
def to_time_cell(values, time_axis: int, device=None):
    """Unwrap an array to a contiguous ``(T, C)`` float32 CUDA tensor.

    Returns ``(x2d, cell_shape)``.  numpy inputs are staged through pinned host memory; CUDA
    tensors already in ``(time, ...)`` float32 contiguous layout are used in place (zero copy).
    """
    _require_cuda()
    if isinstance(values, torch.Tensor):
        t = values
        if time_axis != 0:
            t = t.movedim(time_axis, 0)
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        if not t.is_cuda:
            t = t.contiguous().pin_memory().to(device or "cuda", non_blocking=True)
        t = t.contiguous()
    else:
        a = np.asarray(values)
        if time_axis != 0:
            a = np.moveaxis(a, time_axis, 0)
        if a.dtype != np.float32:
            if a.dtype.kind == "f":
                warnings.warn(f"xclim_b200 computes in float32: casting input from {a.dtype}", stacklevel=3)
            a = a.astype(np.float32)
        a = np.ascontiguousarray(a)
        h = torch.from_numpy(a)
        try:
            h = h.pin_memory()
        except RuntimeError:  # pragma: no cover - pinning can fail for huge buffers
            pass
        t = h.to(device or "cuda", non_blocking=True)
    cell_shape = tuple(t.shape[1:])
    return t.reshape(t.shape[0], -1), cell_shape


_small_cache: dict = {}


def dev_ints(arr, dtype, device) -> torch.Tensor:
    """Small integer index arrays (period offsets, doy index) cached per device."""
    a = np.ascontiguousarray(np.asarray(arr, dtype=dtype))
    key = (str(device), a.dtype.str, a.tobytes())
    t = _small_cache.get(key)
    if t is None:
        if len(_small_cache) > 256:
            _small_cache.clear()
        t = torch.from_numpy(a).to(device)
        _small_cache[key] = t
    return t


# ---------------------------------------------------------------------------------------- kernels
def period_count(x2d, poff, op_code, thr, cmp_f64=False, want_valid=False):
    T, C = x2d.shape
    P = len(poff) - 1
    poff_d = dev_ints(poff, np.int32, x2d.device)
    out = torch.empty((P, C), dtype=torch.int32, device=x2d.device)
    valid = torch.empty((P, C), dtype=torch.int32, device=x2d.device) if want_valid else None
    check(load().xc_period_count_f32(x2d.data_ptr(), T, C, x2d.stride(0), poff_d.data_ptr(), P, op_code,
                                     float(thr), int(bool(cmp_f64)), out.data_ptr(), _ptr(valid),
                                     current_stream_ptr()))
    return out, valid


def period_runstat(x2d, poff, op_code, thr, reducer_code, window, resample_before_rl=True, cmp_f64=False,
                   want_valid=False):
    T, C = x2d.shape
    P = len(poff) - 1
    poff_d = dev_ints(poff, np.int32, x2d.device)
    out = torch.empty((P, C), dtype=torch.float32, device=x2d.device)
    valid = torch.empty((P, C), dtype=torch.int32, device=x2d.device) if want_valid else None
    check(load().xc_period_runstat_f32(x2d.data_ptr(), T, C, x2d.stride(0), poff_d.data_ptr(), P, op_code,
                                       float(thr), int(bool(cmp_f64)), reducer_code, int(window),
                                       int(bool(resample_before_rl)), out.data_ptr(), _ptr(valid),
                                       current_stream_ptr()))
    return out, valid


def period_reduce(x2d, poff, stat_code, transform=0, op_code=0, thr=0.0, want_valid=False):
    T, C = x2d.shape
    P = len(poff) - 1
    poff_d = dev_ints(poff, np.int32, x2d.device)
    out = torch.empty((P, C), dtype=torch.float32, device=x2d.device)
    valid = torch.empty((P, C), dtype=torch.int32, device=x2d.device) if want_valid else None
    check(load().xc_period_reduce_f32(x2d.data_ptr(), T, C, x2d.stride(0), poff_d.data_ptr(), P, stat_code,
                                      transform, op_code, float(thr), out.data_ptr(), _ptr(valid),
                                      current_stream_ptr()))
    return out, valid


def synth(T, C, kind, seed, cell_offset=0, cells_per_lat=1440, n_lat_global=721, year_len=365, device="cuda"):
    """Synthetic (T, C) float32 input generated in HBM (kind 0 = pr mm/d, 1 = tasmax K)."""
    _require_cuda()
    out = torch.empty((T, C), dtype=torch.float32, device=device)
    check(load().xc_synth_f32(out.data_ptr(), T, C, C, cell_offset, cells_per_lat, n_lat_global, year_len,
                              kind, seed, current_stream_ptr()))
    return out
