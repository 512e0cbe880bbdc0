"""Lat-tile sharding of the grid across the GPUs of one box.

Every output value depends on ONE grid cell's time series, so the ``(time, lat, lon)`` grid splits
into contiguous latitude tiles with **no collective on the data path**; one process per GPU computes
its tile, and the small ``(periods, lat_tile, lon)`` outputs are (optionally) gathered with
``torch.distributed.all_gather`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def lat_tiles(n_lat: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [start, stop) latitude ranges, sizes differing by at most one row
    (721 rows on 8 ranks -> 91, 90, ..., 90)."""
    base, extra = divmod(n_lat, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((s, s + n))
        s += n
    return out


def shard_lat(values, lat_axis: int, rank: int, world: int):
    """This rank's contiguous latitude tile of an array (numpy or torch): a strided view of
    ``(time, lat, lon)`` which the unwrap step copies to a contiguous per-GPU buffer."""
    s, e = lat_tiles(values.shape[lat_axis], world)[rank]
    idx = [slice(None)] * values.ndim
    idx[lat_axis] = slice(s, e)
    return values[tuple(idx)]


def gather_lat(local, lat_axis: int, n_lat: int, group=None):
    """Reassemble per-rank tiles (torch tensors, any backend) along ``lat_axis`` on every rank.
    Tiles may differ by one row, so they are padded to the widest tile for ``all_gather``."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    tiles = lat_tiles(n_lat, world)
    widest = max(e - s for s, e in tiles)
    loc = local.movedim(lat_axis, 0).contiguous()
    pad = widest - loc.shape[0]
    if pad:
        loc = torch.cat([loc, loc.new_zeros((pad,) + tuple(loc.shape[1:]))], dim=0)
    bufs = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(bufs, loc, group=group)
    parts = [b[: e - s] for b, (s, e) in zip(bufs, tiles)]
    return torch.cat(parts, dim=0).movedim(0, lat_axis)


def run_sharded(fn, values, lat_axis: int, rank: int, world: int, out_lat_axis: int | None = None, group=None):
    """``fn(tile) -> torch tensor`` on this rank's tile, then all-gather along the lat axis."""
    tile = shard_lat(values, lat_axis, rank, world)
    local = fn(tile)
    if world == 1:
        return local
    return gather_lat(local, lat_axis if out_lat_axis is None else out_lat_axis, values.shape[lat_axis], group)
