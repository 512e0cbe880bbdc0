"""Host-side rate of the slab streamer's staging copy (strided (time, lat-slab, lon) box -> contiguous buffer),
serial against xclim_b200.io.parallel_rows, for native and big-endian (NetCDF-3) sources.  CPU only."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xclim_b200 import io  # noqa: E402

T, Y, X = 2000, 64, 1440
r0, r1 = 8, 40
rng = np.random.default_rng(0)
native = rng.random((T, Y, X), dtype=np.float32)
big = native.astype(">f4")
dst = np.empty((T, r1 - r0, X), np.float32)


def rate(fn, n=5):
    fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return dst.nbytes * n / (time.perf_counter() - t) / 1e9


print(f"cpus {os.cpu_count()}, copy threads {io.copy_threads()}, box {dst.nbytes / 1e6:.0f} MB")
for name, src in (("native float32", native), ("big-endian float32", big)):
    serial = rate(lambda: np.copyto(dst, src[:, r0:r1], casting="unsafe"))
    par = rate(lambda: io.parallel_rows(lambda a, b: np.copyto(dst[a:b], src[a:b, r0:r1], casting="unsafe"), T, dst.nbytes))
    assert np.array_equal(dst, native[:, r0:r1])
    print(f"{name}: serial {serial:.1f} GB/s, parallel_rows {par:.1f} GB/s")
