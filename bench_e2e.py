"""bench.py section: the hot path end to end through the host-buffer C-ABI call
(``xc_period_runstat_f32_host``): pinned host input, H2D + kernels + D2H inside the timed region."""
from __future__ import annotations

import time

import numpy as np


def _bind_to_gpu_numa_node(local_index: int) -> str:
    """Pin this process (and therefore the first-touch placement of its pinned host buffer) to the CPUs
    of the NUMA node the GPU hangs off, so that H2D DMA reads local memory when 8 ranks stream at once.
    Best effort: any failure leaves the affinity untouched.  Disable with XCLIM_B200_NO_NUMA=1."""
    import os
    if os.environ.get("XCLIM_B200_NO_NUMA"):
        return "disabled"
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "no numa info"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "no allowed cpu on node"
        os.sched_setaffinity(0, allowed)
        return f"node {node} ({len(allowed)} cpus)"
    except Exception as e:  # pragma: no cover - depends on the host
        return f"unavailable ({type(e).__name__})"


def e2e_section(args, dev, rank, world, pr_dev, poff, barrier, ref_out=None, ref_valid=None):
    import torch
    import torch.distributed as dist

    from xclim_b200 import _lib, device

    T, C = pr_dev.shape
    P = len(poff) - 1
    numa = _bind_to_gpu_numa_node(dev.index if dev.index is not None else 0)
    # pinned host copy of this rank's tile (filled from the device generator: same bits)
    try:
        x_host = torch.empty((T, C), dtype=torch.float32, pin_memory=True)
    except RuntimeError as e:  # not enough lockable host memory
        return {"value": None, "unit": "grid-cells/s", "error": f"cannot pin {T * C * 4 / 1e9:.1f} GB: {e}"}
    step_rows = 365
    for s in range(0, T, step_rows):
        x_host[s:s + step_rows].copy_(pr_dev[s:s + step_rows], non_blocking=True)
    torch.cuda.synchronize()
    op, red = _lib.OPS["<"], _lib.RL_REDUCERS["max"]
    out_h, valid_h, ws = device.period_runstat_host(x_host, poff, op, 1.0, red, 1)  # warm-up 1 (allocations)
    for _ in range(2):
        device.period_runstat_host(x_host, poff, op, 1.0, red, 1, workspace=ws, out_host=out_h, valid_host=valid_h)
    if ref_out is not None:
        assert torch.equal(out_h, ref_out.cpu()), "e2e result differs from the device-resident result"
        assert torch.equal(valid_h, ref_valid.cpu())
    steps = max(1, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        device.period_runstat_host(x_host, poff, op, 1.0, red, 1, workspace=ws, out_host=out_h, valid_host=valid_h)
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item()) / steps
    h2d = T * C * 4
    d2h = 2 * P * C * 4
    del x_host
    return {"value": C * world / dt, "unit": "grid-cells/s", "ms_per_step": dt * 1e3, "steps": steps,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "h2d_gbs": h2d / dt / 1e9,
            "api": "xc_period_runstat_f32_host (pinned host (T,C) f32 in, (P,C) f32 + i32 out, "
                   "year slabs double-buffered on two streams)",
            "numa_binding": numa,
            "timer": "host perf_counter around the synchronous call, barrier + cuda sync both sides, max over ranks"}
