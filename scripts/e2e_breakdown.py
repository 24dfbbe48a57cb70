#!/usr/bin/env python
"""scripts/e2e_breakdown.py -- where the host time of a batched update goes: stage (pack into the pinned arena) / launch
(H2D + graph launch, asynchronous) / fetch (wait + D2H + unpack), per batch of config-B filters.  Usage: e2e_breakdown.py [nf] [threads]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import capi, engine_filter, synth  # noqa: E402

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
caps = dict(max_clones=38, max_tracks=512, max_obs=512 * 30)
tmpl, batches, work = [], [], []
for i in range(nf):
    f = engine_filter(np.float32, **caps)
    synth.drive(f, synth.make_window_workload(n_features=300, n_clones=30, seq=i % 16), marginalize_last=False)
    off, obs, idx = f.packQueued()
    batches.append(capi.TrackBatch(off, obs, idx, np.float32))
    tmpl.append((f, capi.Engine(np.float32, borrowed=f.engineHandle())))
    work.append(capi.Engine(np.float32, **caps))
b = capi.Batch(work)
tr = b._tracks(batches)
acc = {"stage": [], "launch": [], "fetch": [], "update (one call)": []}
for r in range(30):
    for w, (_, t) in zip(work, tmpl):
        w.copy_state_from(t)
    work[0].synchronize()
    t0 = time.perf_counter()
    b.stage(capi.MARGINALIZE, batches, threads=threads)
    t1 = time.perf_counter()
    b.launch()
    t2 = time.perf_counter()
    b.fetch(batches)
    t3 = time.perf_counter()
    acc["stage"].append(t1 - t0); acc["launch"].append(t2 - t1); acc["fetch"].append(t3 - t2)
    for w, (_, t) in zip(work, tmpl):
        w.copy_state_from(t)
    work[0].synchronize()
    t0 = time.perf_counter()
    b.update(capi.MARGINALIZE, batches, threads=threads)
    acc["update (one call)"].append(time.perf_counter() - t0)
print(f"{nf} filters, {threads} staging threads: us per batch (median of 30):", {k: round(1e6 * float(np.median(v[5:])), 1) for k, v in acc.items()})
