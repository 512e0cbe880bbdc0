#!/usr/bin/env python
"""Build profiles/ncu_traffic.csv: per-kernel DRAM traffic of the committed `ncu --set full` captures.

    python tools/ncu_traffic.py gpurun_out/prof_a.ncu-rep:10950x721x1440 [more.ncu-rep:TxYxX ...]

Every argument is a report and the (time, lat, lon) grid the captured command ran on.  Rows are appended
(kernel, grid, dram_read_bytes, dram_write_bytes, duration_ms, report); bench.py reads `traffic` from this
table instead of carrying literals.  Needs the `ncu` CLI only (no GPU)."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_traffic.csv")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def rows_of(path, grid):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    out = []
    for r in body:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("xc::<unnamed>::", "").strip()
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_write.sum"]]]
        du = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * TIME[units[col["gpu__time_duration.sum"]]]
        out.append({"kernel": name, "grid": grid, "dram_read_bytes": f"{rd:.0f}", "dram_write_bytes": f"{wr:.0f}",
                    "duration_ms": f"{du:.4f}", "report": os.path.basename(path)})
    return out


def main():
    fields = ["kernel", "grid", "dram_read_bytes", "dram_write_bytes", "duration_ms", "report"]
    old = list(csv.DictReader(open(OUT))) if os.path.exists(OUT) else []
    new = []
    for arg in sys.argv[1:]:
        path, grid = arg.rsplit(":", 1)
        new += rows_of(path, grid)
    reports = {r["report"] for r in new}
    keep = [r for r in old if r["report"] not in reports]
    with open(OUT, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=fields)
        w.writeheader()
        for r in keep + new:
            w.writerow(r)
    print(f"{OUT}: {len(keep) + len(new)} rows")


if __name__ == "__main__":
    main()
