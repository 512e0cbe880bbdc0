"""One run of indices.run_batch on a lat band (for ncu launch lists): python tools/batch_once.py [n_lat]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xclim_b200
from xclim_b200 import Field, TimeAxis, calendar as xcal, device, indices
n_lat = int(sys.argv[1]) if len(sys.argv) > 1 else 90
T, X = 10950, 1440
C = n_lat * X
ta = TimeAxis.daily("1981-01-01", T, "noleap")
gen = lambda kind, seed: device.synth(T, C, kind=kind, seed=seed, cells_per_lat=X, n_lat_global=n_lat)
tasmax = gen(1, 3); tasmin = gen(1, 8); tasmin.sub_(8.0); tas = torch.add(tasmax, tasmin).mul_(0.5); pr = gen(0, 2)
mk = lambda t, u: Field(t.view(T, n_lat, X), ("time", "lat", "lon"), ta, {}, {"units": u})
fields = {"tas": mk(tas, "K"), "tasmax": mk(tasmax, "K"), "tasmin": mk(tasmin, "K"), "pr": mk(pr, "mm/d")}
with xclim_b200.set_options(device_outputs=True):
    pers = {(v, p): xcal.select_percentile(xcal.percentile_doy(fields[v], window=5, per=p), p)
            for v, p in (("tasmax", 90.0), ("tasmax", 10.0), ("tasmin", 90.0))}
    torch.cuda.synchronize()
    for _ in range(2):
        out = indices.run_batch(fields, pers)
    torch.cuda.synchronize()
print("ok", len(out))
