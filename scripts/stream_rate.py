#!/usr/bin/env python
"""scripts/stream_rate.py -- frames/s of the E-sim stream through the class (C view), fp32 and fp64, with the time split by call."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import engine_filter, synth  # noqa: E402

wl = synth.make_stream_workload(n_frames=200, seq=8, max_features=60, max_track_length=20, max_cam_states=20)
wl["noise"] = synth.euroc_noise(tuned=True)
for dtype in (np.float32, np.float64):
    for rep in range(3):
        g = engine_filter(dtype)
        acc = {}

        def timed(name, fn, *a):
            t0 = time.perf_counter()
            r = fn(*a)
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            return r

        g.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"])
        t0 = time.perf_counter()
        for fr in wl["frames"]:
            for (w, a, dT) in fr["imu"]:
                timed("propagate", g.propagate, w, a, dT)
            timed("augmentState", g.augmentState, fr["state_id"], fr["time"])
            timed("update", g.update, *fr["update"])
            timed("addFeatures", g.addFeatures, *fr["add"])
            timed("marginalize", g.marginalize)
            timed("pruneEmptyStates", g.pruneEmptyStates)
            timed("getImuState", g.getImuState)
        dt = time.perf_counter() - t0
    print(np.dtype(dtype).name, f"{200 / dt:.0f} frames/s;", "us per frame by call:", {k: round(1e6 * v / 200, 1) for k, v in acc.items()}, "updates", g.counters()["n_updates"])
