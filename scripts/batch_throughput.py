#!/usr/bin/env python
"""scripts/batch_throughput.py -- device-timed and end-to-end throughput of device batches of independent filters
(config B: 300 x 30 fp32 each).  Usage: python scripts/batch_throughput.py 1 8 16 32"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import capi, engine_filter, synth  # noqa: E402


def run(nf, nfeat=300, nclones=30, dtype=np.float32, reps=8):
    caps = dict(max_clones=nclones + 8, max_tracks=max(512, nfeat + 16), max_obs=max(512, nfeat + 16) * nclones)
    tmpl, batches, work = [], [], []
    for i in range(nf):
        f = engine_filter(dtype, **caps)
        synth.drive(f, synth.make_window_workload(n_features=nfeat, n_clones=nclones, seq=(i % 16)), marginalize_last=False)
        off, obs, idx = f.packQueued()
        batches.append(capi.TrackBatch(off, obs, idx, dtype))
        tmpl.append((f, capi.Engine(dtype, borrowed=f.engineHandle())))
        work.append(capi.Engine(dtype, **caps))
    b = capi.Batch(work)
    dev, e2e = [], []
    for r in range(reps):
        for w, (_, t) in zip(work, tmpl):
            w.copy_state_from(t)
        b.stage(capi.MARGINALIZE, batches, threads=4)
        work[0].synchronize()
        dev.append(b.launch_timed())
        reps_ = b.fetch(batches)
    assert all(r["m"] == nfeat * (2 * nclones - 3) for r in reps_), [r["m"] for r in reps_]
    for r in range(reps):
        for w, (_, t) in zip(work, tmpl):
            w.copy_state_from(t)
        work[0].synchronize()
        t0 = time.perf_counter()
        b.update(capi.MARGINALIZE, batches, threads=4)
        e2e.append(time.perf_counter() - t0)
    work[0].set_option(1, 1.0)
    per = {}
    for r in range(4):
        for w, (_, t) in zip(work, tmpl):
            w.copy_state_from(t)
        b.stage(capi.MARGINALIZE, batches, threads=4)
        b.launch_timed()
        b.fetch(batches)
        for name, ms in b.kernel_times():
            per.setdefault(name, []).append(ms)
    work[0].set_option(1, 0.0)
    d, e = float(np.median(dev[3:])), float(np.median(e2e[3:]))
    out = {"filters": nf, "workload": f"{nfeat}x{nclones} {np.dtype(dtype).name}", "device_ms_per_batch": d, "device_updates_per_s": nf / d * 1e3,
           "e2e_ms_per_batch": e * 1e3, "e2e_updates_per_s": nf / e, "kernel_us": {k: round(1e3 * float(np.mean(v[1:])), 1) for k, v in per.items()}}
    b.close()
    return out


if __name__ == "__main__":
    for a in sys.argv[1:]:
        print(json.dumps(run(int(a))), flush=True)
