"""TEST INFRASTRUCTURE: run ``xclim_b200.streaming.run_streamed`` on the CPU.

``install(monkeypatch)`` puts the oracle-backed device layer (tests/fake_device.py) in place and replaces
what the slab streamer takes from CUDA -- streams, events, page-locked / device allocations and the two
C-ABI copy entry points -- by synchronous CPU stand-ins, so that the streamer's HOST logic (slab plan,
chunk alignment, reader thread, staging buffers, result assembly) is exercised by the ``-m "not gpu"``
suite.  It proves nothing about copies overlapping kernels; the ``-m gpu`` tests do.
"""
import ctypes

import torch

import fake_device


class _Stream:
    cuda_stream = 0

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class _Event:
    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


class _Lib:
    """The C-ABI entry points the streamer calls, on host memory."""

    def __init__(self):
        self.boxes = []

    def xc_host_pinned(self, ptr):
        return 0

    def xc_copy_box_async(self, dst, dst_pitch, src, src_pitch, width, rows, kind, stream):
        self.boxes.append((int(width), int(rows), int(kind)))
        for r in range(int(rows)):
            ctypes.memmove(int(dst) + r * int(dst_pitch), int(src) + r * int(src_pitch), int(width))
        return 0


def install(monkeypatch):
    from xclim_b200 import _lib, device
    fake_device.install(monkeypatch)
    lib = _Lib()
    real_empty = torch.empty

    def empty(*size, **kw):
        kw.pop("pin_memory", None)
        if str(kw.get("device", "cpu")).startswith("cuda"):
            kw["device"] = "cpu"
        return real_empty(*size, **kw)

    monkeypatch.setattr(torch, "empty", empty)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(device, "_require_cuda", lambda: None)
    monkeypatch.setattr(_lib, "load", lambda: lib)
    monkeypatch.setattr(_lib, "check", lambda status: None)
    # results that are torch tensors count as device results: they go back through xc_copy_box_async
    # (the strided-box arithmetic of the D2H leg), numpy results take the host branch
    from xclim_b200 import streaming
    monkeypatch.setattr(streaming, "_on_device", lambda v: isinstance(v, torch.Tensor))
    return lib
