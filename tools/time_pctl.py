"""Experiment helper: time xc_percentile_doy_f32 (window 5, per 90, 30 noleap years) on the full grid
and print a table checksum, so that library variants (XCLIM_B200_LIB=...) can be compared.
    python tools/time_pctl.py [n_lat] [reps]"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xclim_b200 import device as dv  # noqa: E402


def main():
    n_lat = int(sys.argv[1]) if len(sys.argv) > 1 else 721
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    T, C = 10950, n_lat * 1440
    x = dv.synth(T, C, 1, 7)
    doy = (np.arange(T) % 365 + 1).astype(np.int16)
    yr = (np.arange(T) // 365).astype(np.int16)
    out = dv.percentile_doy(x, doy, yr, 365, 30, 5, [90.0], 1.0 / 3, 1.0 / 3)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        out = dv.percentile_doy(x, doy, yr, 365, 30, 5, [90.0], 1.0 / 3, 1.0 / 3)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    sample = out[0, :, :: max(1, C // 4096)].cpu().numpy()
    print(json.dumps({"lib": os.environ.get("XCLIM_B200_LIB", "default"), "n_lat": n_lat,
                      "ms": [round(m, 3) for m in ms], "min_ms": round(min(ms), 3),
                      "sha": hashlib.sha1(sample.tobytes()).hexdigest()[:16]}))


if __name__ == "__main__":
    main()
