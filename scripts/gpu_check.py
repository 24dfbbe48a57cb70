"""Developer probe (GPU box): detailed engine-vs-oracle comparison on a few workloads."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from msckf_mono_b200 import synth, engine_filter
from msckf_mono_b200.cview import CFilter

def rel(x, y):
    x = np.asarray(x, float); y = np.asarray(y, float)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))

def oracle(dtype, drop=False):
    o = CFilter(ROOT / "oracle" / "libmsckf_oracle.so", "msckf_oracle_", dtype)
    o.setOption(1, 1.0 if drop else 0.0)
    return o

def compare(tag, wl, dtype, drop, upto=None, **ekw):
    g = engine_filter(dtype, **ekw)
    o = oracle(dtype, drop)
    # state before the last marginalize
    synth.drive(g, wl, upto=upto, marginalize_last=False)
    synth.drive(o, wl, upto=upto, marginalize_last=False)
    print(f"[{tag} {np.dtype(dtype).name} drop={drop}] pre: P rel {rel(g.getCovariance(), o.getCovariance()):.2e} "
          f"imu p {rel(g.getImuState()['p_I_G'], o.getImuState()['p_I_G']):.2e} q {rel(g.getImuState()['q_IG'], o.getImuState()['q_IG']):.2e} "
          f"cam p {rel(g.getCamStates()['p_C_G'], o.getCamStates()['p_C_G']):.2e}")
    t0 = time.time(); g.marginalize(); tg = time.time() - t0
    t0 = time.time(); o.marginalize(); to = time.time() - t0
    rg, ro = g.lastReport(), o.lastReport()
    print(f"   tracks {len(rg['valid'])} valid eq {np.array_equal(rg['valid'], ro['valid'])} acc eq {np.array_equal(rg['accepted'], ro['accepted'])} "
          f"acc {int(rg['accepted'].sum())} gamma rel {rel(rg['gamma'], ro['gamma']):.2e} pfg rel {rel(rg['p_f_G'], ro['p_f_G']):.2e}")
    print(f"   dx rel {rel(g.lastDeltaX(), o.lastDeltaX()):.3e} |dx| {np.linalg.norm(o.lastDeltaX()):.3e} P rel {rel(g.getCovariance(), o.getCovariance()):.3e} "
          f"maxabs {np.abs(g.getCovariance()-o.getCovariance()).max()/np.abs(o.getCovariance()).max():.3e} counters g {g.counters()} o {o.counters()}")
    print(f"   imu p {np.abs(g.getImuState()['p_I_G']-o.getImuState()['p_I_G']).max():.2e} cam p {np.abs(g.getCamStates()['p_C_G']-o.getCamStates()['p_C_G']).max():.2e} "
          f"time gpu {tg*1e3:.2f} ms cpu {to*1e3:.1f} ms")

if __name__ == "__main__":
    for dtype in (np.float64, np.float32):
        for drop in (True, False):
            compare("3x4", synth.make_window_workload(n_features=3, n_clones=4, seq=5), dtype, drop)
            compare("8x6", synth.make_window_workload(n_features=8, n_clones=6, seq=3), dtype, drop)
            compare("40x12", synth.make_window_workload(n_features=40, n_clones=12, seq=4), dtype, drop)
            compare("300x30", synth.make_window_workload(n_features=300, n_clones=30, seq=0), dtype, drop)
    wl = synth.make_stream_workload(n_frames=60, seq=6, max_features=30, max_track_length=12, max_cam_states=10)
    for dtype in (np.float64, np.float32):
        g = engine_filter(dtype); o = oracle(dtype, True)
        errs = []
        def onf(k, f):
            pass
        fr = wl["frames"]
        g.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"]); o.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"])
        mism = 0
        for k, f in enumerate(fr):
            for filt in (g, o):
                for (w, a, dT) in f["imu"]: filt.propagate(w, a, dT)
                filt.augmentState(f["state_id"], f["time"]); filt.update(*f["update"]); filt.addFeatures(*f["add"]); filt.marginalize(); filt.pruneEmptyStates()
            rg, ro = g.lastReport(), o.lastReport()
            if not (np.array_equal(rg["valid"], ro["valid"]) and np.array_equal(rg["accepted"], ro["accepted"])): mism += 1
        print(f"[stream {np.dtype(dtype).name}] flag mismatches {mism} M {g.getNumCamStates()}/{o.getNumCamStates()} P rel {rel(g.getCovariance(), o.getCovariance()):.2e} "
              f"imu p diff {np.abs(g.getImuState()['p_I_G']-o.getImuState()['p_I_G']).max():.2e} tracked eq {np.array_equal(g.getTrackedFeatureIds(), o.getTrackedFeatureIds())} "
              f"counters {g.counters()} {o.counters()}")
