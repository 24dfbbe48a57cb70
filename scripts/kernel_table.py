#!/usr/bin/env python
"""scripts/kernel_table.py -- per-kernel device times (CUDA events between the kernels, engine option 1) and the
device-timed total of one marginalize() for a given window workload.  Usage: python scripts/kernel_table.py 2000 60 f64"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import capi, engine_filter, synth  # noqa: E402


def table(nf, nc, dtype, reps=6, options=()):
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=30)
    caps = dict(max_clones=nc + 8, max_tracks=max(512, nf + 48), max_obs=max(512, nf + 48) * nc)
    f = engine_filter(dtype, **caps)
    synth.drive(f, wl, marginalize_last=False)
    off, obs, idx = f.packQueued()
    batch = capi.TrackBatch(off, obs, idx, dtype)
    tmpl = capi.Engine(dtype, borrowed=f.engineHandle())
    work = capi.Engine(dtype, **caps)
    for key, val in options:
        work.set_option(key, val)
    tot = []
    for _ in range(reps):
        work.copy_state_from(tmpl)
        work.stage(capi.MARGINALIZE, batch)
        work.synchronize()
        tot.append(work.launch_timed())
        rep = work.fetch(batch.n_tracks)
    work.set_option(1, 1.0)
    per = {}
    for _ in range(reps):
        work.copy_state_from(tmpl)
        work.stage(capi.MARGINALIZE, batch)
        work.synchronize()
        work.launch_timed()
        work.fetch(batch.n_tracks)
        for name, ms in work.kernel_times():
            per.setdefault(name, []).append(ms)
    return {"workload": f"{nf}x{nc} {np.dtype(dtype).name}", "m": rep["m"], "rank": rep["rank"], "accepted": int(rep["accepted"].sum()),
            "ms_total_graph": float(np.median(tot[2:])), "kernel_us": {k: round(1e3 * float(np.mean(v[1:])), 1) for k, v in per.items()}}


if __name__ == "__main__":
    nf, nc = int(sys.argv[1]), int(sys.argv[2])
    dtype = np.float64 if sys.argv[3] == "f64" else np.float32
    opts = [(int(a.split("=")[0]), float(a.split("=")[1])) for a in sys.argv[4:]]
    out = table(nf, nc, dtype, options=opts)
    out["options"] = opts
    print(json.dumps(out))
