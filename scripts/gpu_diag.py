import sys, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from tests.common import *
from msckf_mono_b200 import engine_filter
lib = ROOT / "oracle" / "libmsckf_oracle.so"

def fp32_case(nf, nc, seq):
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq)
    g = engine_filter(np.float32); o32 = make_oracle(lib, np.float32); o64 = make_oracle(lib, np.float64, drop_null_rows=True)
    for f in (g, o32, o64):
        f._round = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    for f in (g, o32, o64):
        synth.drive(f, wl, marginalize_last=False)
    print(f"fp32 {nf}x{nc}: pre P err g {rel(g.getCovariance(), o64.getCovariance()):.2e} o32 {rel(o32.getCovariance(), o64.getCovariance()):.2e}")
    for f in (g, o32, o64): f.marginalize()
    rg, r32, r64 = g.lastReport(), o32.lastReport(), o64.lastReport()
    print(f"   pfg err g {rel(rg['p_f_G'], r64['p_f_G']):.2e} o32 {rel(r32['p_f_G'], r64['p_f_G']):.2e} | gamma err g {rel(rg['gamma'], r64['gamma']):.2e} o32 {rel(r32['gamma'], r64['gamma']):.2e}")
    print(f"   dx err g {rel(g.lastDeltaX(), o64.lastDeltaX()):.2e} o32 {rel(o32.lastDeltaX(), o64.lastDeltaX()):.2e} | P err g {rel(g.getCovariance(), o64.getCovariance()):.2e} o32 {rel(o32.getCovariance(), o64.getCovariance()):.2e}")
    eg = np.linalg.norm(rg['p_f_G'] - r64['p_f_G'], axis=1); eo = np.linalg.norm(r32['p_f_G'] - r64['p_f_G'], axis=1)
    print(f"   per-feature |pfg err| median g {np.median(eg):.2e} o32 {np.median(eo):.2e} max g {eg.max():.2e} o32 {eo.max():.2e}")

def try_(name, fn):
    try:
        fn()
    except Exception:
        print(f"[{name}] EXC"); traceback.print_exc(limit=3)

def stream_case(dtype, prune_red, **kw):
    wl = synth.make_stream_workload(**kw)
    if prune_red:
        wl["params"]["redundancy_angle_thresh"] = 0.2; wl["params"]["redundancy_distance_thresh"] = 0.2
    g = engine_filter(dtype); o = make_oracle(lib, dtype, drop_null_rows=(dtype == np.float64))
    g.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"]); o.initialize(wl["camera"], wl["noise"], wl["params"], wl["imu_state"])
    bad = 0
    for k, fr in enumerate(wl["frames"]):
        for f in (g, o):
            for (w, a, dT) in fr["imu"]: f.propagate(w, a, dT)
            f.augmentState(fr["state_id"], fr["time"]); f.update(*fr["update"]); f.addFeatures(*fr["add"]); f.marginalize()
        rg, ro = g.lastReport(), o.lastReport()
        if not (np.array_equal(rg["valid"], ro["valid"]) and np.array_equal(rg["accepted"], ro["accepted"]) and np.array_equal(rg["cm_passed"], ro["cm_passed"])):
            bad += 1
            if bad <= 3:
                print(f"   frame {k}: flags differ valid {rg['valid']} vs {ro['valid']} acc {rg['accepted']} vs {ro['accepted']} cm {rg['cm_passed']} vs {ro['cm_passed']}")
                print(f"      gamma g {rg['gamma']} o {ro['gamma']}")
        if prune_red:
            g.pruneRedundantStates(); o.pruneRedundantStates()
        g.pruneEmptyStates(); o.pruneEmptyStates()
        if g.getNumCamStates() != o.getNumCamStates():
            print(f"   frame {k}: M differs {g.getNumCamStates()} {o.getNumCamStates()}"); break
        sg, so = g.getCamStates(), o.getCamStates()
        if not np.array_equal(sg["state_id"], so["state_id"]):
            print(f"   frame {k}: clone ids differ {sg['state_id']} {so['state_id']}"); break
        if not np.array_equal(sg["last_correlated_id"], so["last_correlated_id"]):
            print(f"   frame {k}: last_correlated differ {sg['last_correlated_id']} {so['last_correlated_id']}"); break
    print(f"stream {np.dtype(dtype).name} prune_red={prune_red}: bad frames {bad} M {g.getNumCamStates()} P rel {rel(g.getCovariance(), o.getCovariance()):.2e} "
          f"imu p {np.abs(g.getImuState()['p_I_G'] - o.getImuState()['p_I_G']).max():.2e} tracked eq {np.array_equal(g.getTrackedFeatureIds(), o.getTrackedFeatureIds())} "
          f"pruned eq {np.array_equal(g.getPrunedStates()['state_id'], o.getPrunedStates()['state_id'])} npruned {len(g.getPrunedStates()['state_id'])} counters {g.counters()} {o.counters()}")

def reject_case(dtype):
    wl = synth.make_window_workload(n_features=120, n_clones=10, seq=12)
    synth.corrupt_observations(wl, seed=3)
    wl["params"]["translation_threshold"] = 0.2
    g = engine_filter(dtype); o = make_oracle(lib, dtype, drop_null_rows=(dtype == np.float64))
    synth.drive(g, wl); synth.drive(o, wl)
    rg, ro = g.lastReport(), o.lastReport()
    print(f"reject {np.dtype(dtype).name}: cm sum {rg['cm_passed'].sum()}/{ro['cm_passed'].sum()} eq {np.array_equal(rg['cm_passed'], ro['cm_passed'])} valid {rg['valid'].sum()}/{ro['valid'].sum()} eq {np.array_equal(rg['valid'], ro['valid'])} "
          f"acc {rg['accepted'].sum()}/{ro['accepted'].sum()} eq {np.array_equal(rg['accepted'], ro['accepted'])} counters {g.counters()} {o.counters()}")
    d = np.flatnonzero((rg['valid'] != ro['valid']) | (rg['accepted'] != ro['accepted']) | (rg['cm_passed'] != ro['cm_passed']))
    print("   differing tracks", d[:10], "gamma g", rg['gamma'][d[:10]], "o", ro['gamma'][d[:10]])
    print(f"   dx rel {rel(g.lastDeltaX(), o.lastDeltaX()):.2e}")

if __name__ == "__main__":
    for c in [(8, 6, 3), (40, 12, 4), (300, 30, 0)]:
        try_("fp32", lambda: fp32_case(*c))
    for dt in (np.float64, np.float32):
        try_("stream", lambda: stream_case(dt, False, n_frames=150, seq=7, max_features=40, max_track_length=14, max_cam_states=12))
        try_("prune", lambda: stream_case(dt, True, n_frames=90, seq=9, max_features=40, max_track_length=40, max_cam_states=21))
        try_("reject", lambda: reject_case(dt))
