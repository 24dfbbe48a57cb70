#!/usr/bin/env python
"""scripts/tail_profile.py -- %globaltimer stamps of the tail kernel's phases (CTA 0) for one config-B update.
Usage: python scripts/tail_profile.py [nf nc f32|f64]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import capi, engine_filter, synth  # noqa: E402

nf, nc = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (300, 30)
dtype = np.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else np.float32
wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=0)
f = engine_filter(dtype, max_clones=nc + 8, max_tracks=max(512, nf + 16), max_obs=max(512, nf + 16) * nc)
synth.drive(f, wl, marginalize_last=False)
off, obs, idx = f.packQueued()
batch = capi.TrackBatch(off, obs, idx, dtype)
tmpl = capi.Engine(dtype, borrowed=f.engineHandle())
work = capi.Engine(dtype, max_clones=nc + 8, max_tracks=max(512, nf + 16), max_obs=max(512, nf + 16) * nc)
work.set_option(1, 1.0)
for _ in range(4):
    work.copy_state_from(tmpl)
    work.stage(capi.MARGINALIZE, batch)
    work.synchronize()
    work.launch_timed()
    work.fetch(batch.n_tracks)
buf = (C.c_ulonglong * 80)()
capi.lib().msckf_b200_tail_profile(work.h, buf, 80)
st = [int(x) for x in list(buf)[:20] if x]
ws = [int(x) for x in list(buf)[20:40] if x]
rel = [round((x - st[0]) / 1e3, 1) for x in st]
print("tail stamps (us since start):", rel)
print("  deltas:", [round(b - a, 1) for a, b in zip(rel[:-1], rel[1:])])
if ws:
    print("  worker CTA 1 (us since kernel start; per block: B1 passed, panel, B2 passed, exchanged, trailing):", [round((x - st[0]) / 1e3, 1) for x in ws])
print("  block 0 factor (us):", round((int(buf[77]) - int(buf[76])) / 1e3, 2))
dd = [int(x) for x in list(buf)[60:74]]
if dd[0]:
    print("  block 0 fine stamps (us since entry; entry | per panel 0..2: micro-block + rows, barrier, trailing, barrier | end):",
          [round((x - dd[0]) / 1e3, 2) if x else None for x in dd])
sj = [int(x) for x in list(buf)[40:60] if x]
if sj:
    rj = [round((x - sj[0]) / 1e3, 1) for x in sj]
    print("jac stamps (us since start; bookkeeping, X/r, QR|Y, transform, Cholesky, outputs):", rj)
print("kernel times:", [(n, round(1e3 * ms, 1)) for n, ms in work.kernel_times()])
