#!/usr/bin/env python
"""Summarise `ncu --set full` reports into the markdown tables kept under profiles/.

    python tools/ncu_summarise.py gpurun_out/prof_a.ncu-rep [more.ncu-rep ...] > profiles/ncu_xxx.md

Needs the `ncu` CLI (no GPU): reads `ncu -i <rep> --page raw --csv`."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    out = []
    for r in body:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        out.append(f"### `{name}`  grid={r[col['Grid Size']]} block={r[col['Block Size']]}\n")
        out.append("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in col and r[col[m]] != "":
                out.append(f"| {m} | {r[col[m]]} | {units[col[m]]} |")
        stalls = {}
        for h, i in col.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:7]
        out.append("\nWarp-state samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top if v > 0) + "\n")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"## {p}\n")
        print(summarise(p))
