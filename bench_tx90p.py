"""bench.py section: tx90p = percentile_doy(window=5, per=90) over the 30-year base + doy-threshold
count (BASELINE.json configs[2], sub-case 3a of SURVEY.md section 8d)."""
from __future__ import annotations

import numpy as np

T_FULL, YEAR = 10950, 365


def tx90p_section(args, dev, rank, world, peak, barrier):
    import torch
    import torch.distributed as dist

    from oracle import xclim_oracle as O
    from xclim_b200 import _lib, device

    T, Y, X = T_FULL, args.lat, 1440
    C = Y * X
    N = T // YEAR
    poff = np.arange(N + 1, dtype=np.int32) * YEAR
    doy = (np.arange(T) % YEAR + 1).astype(np.int16)
    yidx = (np.arange(T) // YEAR).astype(np.int16)
    tasmax = device.synth(T, C, kind=1, seed=3, cell_offset=rank * C, cells_per_lat=X, n_lat_global=Y * world)
    steps = max(1, min(args.steps, 5))

    def run_percentile():
        return device.percentile_doy(tasmax, doy, yidx, YEAR, N, 5, [90.0], 1 / 3, 1 / 3)

    def run_count(tab):
        return device.doy_threshold_count(tasmax, poff, doy, tab, _lib.OPS[">"], want_valid=True)

    table = run_percentile()
    cnt, valid = run_count(table[0])
    for _ in range(2):
        table = run_percentile()
        cnt, valid = run_count(table[0])
    barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
    e[0].record()
    for i in range(steps):
        table = run_percentile()
        e[2 * i + 1].record()
        cnt, valid = run_count(table[0])
        e[2 * i + 2].record()
    barrier()
    t_per = float(np.mean([e[2 * i].elapsed_time(e[2 * i + 1]) for i in range(steps)]))
    t_cnt = float(np.mean([e[2 * i + 1].elapsed_time(e[2 * i + 2]) for i in range(steps)]))
    tt = torch.tensor([t_per, t_cnt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_per, t_cnt = (float(v) for v in tt.tolist())

    # ---- checks: (1) a sample of cells against the CPU oracle (bit-exact table, exact counts)
    sel = torch.tensor([0, 1, C // 3, C // 2 + 7, C - 1], device=dev)
    xs = tasmax[:, sel].cpu().numpy()
    tab_o = O.percentile_doy(xs, yidx.astype(np.int64), doy.astype(np.int64), 5, 90.0)[:, 0]
    tab_g = table[0][:, sel].cpu().numpy()
    table_equal = bool(np.array_equal(tab_o, tab_g, equal_nan=True))
    cnt_o = O.doy_threshold_count(xs, tab_o, doy.astype(np.int64), poff, ">")
    counts_equal = bool(np.array_equal(cnt_o, cnt[:, sel].cpu().numpy()))
    # (2) size-independent property: over the base period ~10 % of days exceed the 90th percentile
    frac = float(cnt.double().mean().item()) / YEAR
    assert 0.07 < frac < 0.13, frac
    assert table_equal and counts_equal, (table_equal, counts_equal)

    # ---- CPU side of the same path: the oracle port (rolling construct + unstack + sort-based quantile +
    # reindex-by-doy count, like the reference) on a bounded sample, one process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import time
        ncell = 4096
        xs_cpu = tasmax[:, :ncell].cpu().numpy()
        t0 = time.perf_counter()
        tab_c = O.percentile_doy(xs_cpu, yidx.astype(np.int64), doy.astype(np.int64), 5, 90.0)[:, 0]
        O.doy_threshold_count(xs_cpu, tab_c, doy.astype(np.int64), poff, ">")
        dt = time.perf_counter() - t0
        cpu = {"value": ncell / dt, "unit": "grid-cells/s", "cores": 1, "kind": "port",
               "sample": f"(10950, {ncell}) cells, percentile_doy + count, {dt:.1f} s"}

    alg_per = T * C * 4 + YEAR * C * 8
    alg_cnt = T * C * 4 + YEAR * C * 8 + N * C * 4 * 2
    ach_per = alg_per / (t_per * 1e-3) / 1e9
    ach_cnt = alg_cnt / (t_cnt * 1e-3) / 1e9
    return {
        "workload": f"percentile_doy(window=5, per=90, base=30 yr) + tx90p count on tasmax ({T},{Y},{X}) f32 per GPU",
        "value": C * world / ((t_per + t_cnt) * 1e-3), "unit": "grid-cells/s",
        "ms_percentile_doy": t_per, "ms_count": t_cnt, "steps": steps,
        "roofline_percentile_doy": {"bound": "hbm", "achieved": ach_per, "peak": peak, "unit": "GB/s",
                                    "frac": ach_per / peak, "algorithmic_bytes": alg_per,
                                    "kernel": "percentile_doy_w5p_kernel<16>"},
        "roofline_count": {"bound": "hbm", "achieved": ach_cnt, "peak": peak, "unit": "GB/s",
                           "frac": ach_cnt / peak, "algorithmic_bytes": alg_cnt, "kernel": "doy_count_years_kernel<GT,3,VALID>"},
        "gpu_launches_per_step": 2, "cpu_baseline": cpu,
        "check": {"oracle_cells": int(sel.numel()), "table_bit_exact": table_equal, "counts_exact": counts_equal,
                  "mean_exceedance_fraction": frac},
    }
