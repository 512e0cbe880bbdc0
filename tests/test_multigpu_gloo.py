"""The N > 1 path on CPU: lat-tile partition + all_gather over gloo (world_size 2 and 3).  The
per-tile computation is stood in by the CPU oracle (no CUDA here); what is tested is that sharding
by latitude and reassembling gives exactly the unsharded result, with uneven tiles."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_lat, q):
    sys.path.insert(0, ROOT)
    from oracle import xclim_oracle as O
    from xclim_b200 import multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)          # same global grid on every rank
    x = rng.gamma(0.4, 6.0, size=(365 * 2, n_lat, 6)).astype(np.float32)
    x[rng.random(x.shape) < 0.4] = 0
    poff = np.array([0, 365, 730])

    def fn(tile):
        return torch.from_numpy(O.maximum_consecutive_dry_days(np.ascontiguousarray(tile), 1.0, poff))

    out = multigpu.run_sharded(fn, x, 1, rank, world)
    full = O.maximum_consecutive_dry_days(x, 1.0, poff)
    ok = bool(np.array_equal(out.numpy(), full))
    tiles = multigpu.lat_tiles(n_lat, world)
    q.put((rank, ok, tiles))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_lat", [(2, 7), (3, 10)])
def test_lat_tile_sharding_over_gloo(world, n_lat):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_lat, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    tiles = res[0][2]
    assert tiles[0][0] == 0 and tiles[-1][1] == n_lat and all(a[1] == b[0] for a, b in zip(tiles, tiles[1:]))


def test_lat_tiles_721_on_8():
    from xclim_b200.multigpu import lat_tiles
    t = lat_tiles(721, 8)
    assert [e - s for s, e in t] == [91] + [90] * 7
