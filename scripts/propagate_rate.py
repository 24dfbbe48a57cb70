#!/usr/bin/env python
"""scripts/propagate_rate.py -- device time of msckf_b200_propagate_n (10 readings per call) on a 20-clone window."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import capi, engine_filter, synth  # noqa: E402

for dtype in (np.float32, np.float64):
    wl = synth.make_window_workload(n_features=30, n_clones=20, seq=3)
    f = engine_filter(dtype)
    synth.drive(f, wl, marginalize_last=False)
    e = capi.Engine(dtype, borrowed=f.engineHandle())
    r = np.tile(np.array([0.01, -0.02, 0.015, 0.1, -0.05, 9.8, 0.005]), (10, 1))
    for k in (1, 10):
        rk = r[:k]
        e.propagate_n(rk); e.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            e.propagate_n(rk)
        e.synchronize()
        dt = (time.perf_counter() - t0) / 200
        print(np.dtype(dtype).name, f"propagate_n({k}): {1e6 * dt:.1f} us per call, {1e6 * dt / k:.1f} us per reading")
    import ctypes as C
    e.set_option(1, 1.0)
    e.propagate_n(r[:2]); e.synchronize()
    buf = (C.c_ulonglong * 80)()
    capi.lib().msckf_b200_tail_profile(e.h, buf, 80)
    st = [int(x) for x in list(buf)[:16]]
    print("  stamps (us; start, on chip, then per reading: calcF/RK, Pade, LU, observability, covariance):", [round((x - st[0]) / 1e3, 2) for x in st if x])
    e.set_option(1, 0.0)
