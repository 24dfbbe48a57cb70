import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from tests.common import *
from msckf_mono_b200 import engine_filter
lib = ROOT / "oracle" / "libmsckf_oracle.so"

def stream(thr, drop, nfr=150):
    wl = synth.make_stream_workload(n_frames=nfr, seq=7, max_features=40, max_track_length=14, max_cam_states=12)
    g = engine_filter(np.float64); o = make_oracle(lib, np.float64, drop_null_rows=drop)
    g.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"]); o.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"])
    if thr is not None: g.setOption(100, thr)
    shown = 0
    for k, fr in enumerate(wl["frames"]):
        for f in (g, o):
            for (w, a, dT) in fr["imu"]: f.propagate(w, a, dT)
            f.augmentState(fr["state_id"], fr["time"]); f.update(*fr["update"]); f.addFeatures(*fr["add"])
        Ppre = rel(g.getCovariance(), o.getCovariance())
        g.marginalize(); o.marginalize()
        cg, co = g.counters(), o.counters()
        if len(g.lastReport()["valid"]) and co["m"] > 0:
            d = rel(g.lastDeltaX(), o.lastDeltaX())
            if d > 1e-7 and shown < 12:
                shown += 1
                L = [int(x) for x in g.queuedTracks()[1]]
                print(f"   frame {k}: dx rel {d:.2e} |dx| {np.linalg.norm(o.lastDeltaX()):.2e} m {co['m']} rank g {cg['rows_kept']} o {co['rows_kept']} M {g.getNumCamStates()} Ppre {Ppre:.1e} Ppost {rel(g.getCovariance(), o.getCovariance()):.1e} L {L}")
        g.pruneEmptyStates(); o.pruneEmptyStates()
    print(f"thr {thr} drop {drop}: final P rel {rel(g.getCovariance(), o.getCovariance()):.2e} imu p {np.abs(g.getImuState()['p_I_G'] - o.getImuState()['p_I_G']).max():.2e}")

if __name__ == "__main__":
    stream(None, True)
    stream(None, False)
    stream(1e-13, True)
    stream(1e-9, True)
