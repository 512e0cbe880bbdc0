"""Experiment helper: time xc_eqm_train_f32 / xc_eqm_adjust_f32 on a lat band (XCLIM_B200_LIB selects a variant)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xclim_b200 import device as dv
n_lat = int(sys.argv[1]) if len(sys.argv) > 1 else 180
T, C = 10950, n_lat * 1440
ref = dv.synth(T, C, 1, 4, cells_per_lat=1440, n_lat_global=n_lat)
hist = dv.synth(T, C, 1, 5, cells_per_lat=1440, n_lat_global=n_lat)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
ms_train = t(lambda: dv.eqm_train(ref, hist, 20, 0))
af, hq = dv.eqm_train(ref, hist, 20, 0)
ms_adj = t(lambda: dv.eqm_adjust(hist, af, hq, 0, 1))
chk = float(af.double().nan_to_num().sum().item())
print(json.dumps({"lib": os.environ.get("XCLIM_B200_LIB", "default"), "train_ms": round(ms_train, 3),
                  "adjust_ms": round(ms_adj, 3), "af_sum": chk}))
