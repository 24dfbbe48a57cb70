#!/usr/bin/env python
"""scripts/measure_parity.py -- prints (and writes as JSON) every parity number the GPU tests assert, so that the bounds in
tests/ can be set to measured x 10.  Test infrastructure: drives the engine and the CPU oracle on the
same seeded inputs.  Usage (GPU box):  python scripts/measure_parity.py [--full-s] > gpurun_out/parity.json
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import engine_filter, synth  # noqa: E402
from tests.common import make_oracle, quat_err, rel, run_collect, state_of  # noqa: E402

ORACLE = ROOT / "oracle" / "libmsckf_oracle.so"
out = {}


def log(k, v):
    out[k] = v
    print(f"{k}: {json.dumps(v)}", file=sys.stderr, flush=True)


def pair_numbers(g, o, rg, ro):
    sg, so = state_of(g), state_of(o)
    rep_g, rep_o = g.lastReport(), o.lastReport()
    flips = int((np.asarray(rg["accepted"]) != np.asarray(ro["accepted"])).sum() + (np.asarray(rg["valid"]) != np.asarray(ro["valid"])).sum())
    return {"flips": flips, "dx": rel(g.lastDeltaX(), o.lastDeltaX()), "P": float(np.abs(sg["P"] - so["P"]).max() / np.abs(so["P"]).max()),
            "gamma": rel(rep_g["gamma"], rep_o["gamma"]), "pfg": rel(rep_g["p_f_G"], rep_o["p_f_G"]),
            "imu_p": float(np.abs(sg["imu_p"] - so["imu_p"]).max()), "cam_p": float(np.abs(sg["cam_p"] - so["cam_p"]).max()),
            "imu_q": quat_err(sg["imu_q"], so["imu_q"]), "cam_q": quat_err(sg["cam_q"], so["cam_q"]),
            "rank_g": g.counters()["rows_kept"], "rank_o": o.counters()["rows_kept"], "m": g.counters()["m"]}


def window(nf, nc, seq, dtype, oracle_kw, eng_kw=None, **wkw):
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq, **wkw)
    g = engine_filter(dtype, **(eng_kw or {}))
    o = make_oracle(ORACLE, dtype, **oracle_kw)
    if dtype == np.float32:
        for f in (g, o):
            f._round = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    t0 = time.time()
    rg = run_collect(g, wl)
    t1 = time.time()
    ro = run_collect(o, wl)
    t2 = time.time()
    r = pair_numbers(g, o, rg, ro)
    r["t_engine_s"], r["t_oracle_s"] = round(t1 - t0, 2), round(t2 - t1, 2)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-s", action="store_true", help="also the full 2000 x 60 fp64 case against the oracle (minutes of CPU)")
    args = ap.parse_args()
    W = [(3, 4, 5), (8, 6, 3), (40, 12, 4), (300, 30, 0)]
    for nf, nc, seq in W:
        log(f"f64_clean_{nf}x{nc}", window(nf, nc, seq, np.float64, dict(drop_null_rows=True)))
        log(f"f64_faithful_{nf}x{nc}", window(nf, nc, seq, np.float64, dict(faithful_max_rows=900)))
        log(f"f32_direct_{nf}x{nc}", window(nf, nc, seq, np.float32, dict()))
        log(f"f32_direct_clean_{nf}x{nc}", window(nf, nc, seq, np.float32, dict(drop_null_rows=True)))
    log("f64_iso_40x12", window(40, 12, 4, np.float64, dict(), isotropic=True))
    big = dict(max_clones=64, max_tracks=2048, max_obs=2048 * 60)
    # config-S shapes: the oracle needs minutes there -> engine vs the committed oracle outputs (tests/golden/stress_*.npz)
    from tests.golden.make_golden_stress import load_P
    for nf, nc, seq, dtype, tag in ((500, 60, 30, np.float64, "stress_f64_500x60"), (2000, 60, 30, np.float64, "stress_f64_2000x60"),
                                    (300, 30, 0, np.float32, "configB_f32_300x30")):
        wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq)
        g = engine_filter(dtype, **big)
        if dtype == np.float32:
            g._round = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
        t0 = time.time()
        rg = run_collect(g, wl)
        sg = state_of(g)
        rep = g.lastReport()
        np.savez_compressed(ROOT / "gpurun_out" / f"engine_{tag}.npz", P=sg["P"], dx=g.lastDeltaX(), gamma=rep["gamma"], pfg=rep["p_f_G"],
                            accepted=rg["accepted"], valid=rg["valid"], imu_p=sg["imu_p"], cam_p=sg["cam_p"], rank=g.counters()["rows_kept"])
        for mode in ("clean", "faithful"):
            gp = ROOT / "tests" / "golden" / f"{tag}_{mode}.npz"
            if not gp.exists():
                continue
            gold = np.load(gp)
            Po = load_P(gold)
            log(f"{tag}_{mode}", {"flips": int((rg["accepted"] != gold["all_accepted"]).sum() + (rg["valid"] != gold["all_valid"]).sum()),
                                  "dx": rel(g.lastDeltaX(), gold["dx"]), "P": float(np.abs(sg["P"] - Po).max() / np.abs(Po).max()),
                                  "gamma": rel(rep["gamma"], gold["gamma"]), "pfg": rel(rep["p_f_G"], gold["p_f_G"]),
                                  "imu_p": float(np.abs(sg["imu_p"] - gold["imu_p"]).max()), "rank_g": g.counters()["rows_kept"],
                                  "rank_o": int(gold["rows_kept"]), "t_engine_s": round(time.time() - t0, 2)})
    # rejections case
    for dtype in (np.float64, np.float32):
        wl = synth.make_window_workload(n_features=120, n_clones=10, seq=12)
        synth.corrupt_observations(wl, seed=3)
        wl["params"]["translation_threshold"] = 0.2
        g, o = engine_filter(dtype), make_oracle(ORACLE, dtype, drop_null_rows=(dtype == np.float64))
        rg, ro = run_collect(g, wl), run_collect(o, wl)
        log(f"rejections_{np.dtype(dtype).name}", pair_numbers(g, o, rg, ro))
    # streams
    for dtype in (np.float64, np.float32):
        wl = synth.make_stream_workload(n_frames=150, seq=7, max_features=40, max_track_length=14, max_cam_states=12)
        g, o = engine_filter(dtype), make_oracle(ORACLE, dtype, drop_null_rows=(dtype == np.float64))
        rg, ro = run_collect(g, wl), run_collect(o, wl)
        log(f"stream150_{np.dtype(dtype).name}", pair_numbers(g, o, rg, ro))
        wl = synth.make_stream_workload(n_frames=90, seq=9, max_features=40, max_track_length=40, max_cam_states=21)
        wl["params"]["redundancy_angle_thresh"] = 0.2
        wl["params"]["redundancy_distance_thresh"] = 0.2
        g, o = engine_filter(dtype), make_oracle(ORACLE, dtype, drop_null_rows=(dtype == np.float64))
        rg, ro = run_collect(g, wl, prune_redundant=True), run_collect(o, wl, prune_redundant=True)
        log(f"prune_redundant_{np.dtype(dtype).name}", pair_numbers(g, o, rg, ro))
    # trajectories (E-sim 200 frames): RMS position difference engine vs oracle, fp64 and fp32, plus frames/s
    for dtype in (np.float64, np.float32):
        wl = synth.make_stream_workload(n_frames=200, seq=8, max_features=60, max_track_length=20, max_cam_states=20)
        wl["noise"] = synth.euroc_noise(tuned=True)
        g, o = engine_filter(dtype), make_oracle(ORACLE, dtype)
        pg, po = [], []
        t0 = time.time()
        synth.drive(g, wl, on_frame=lambda k, f: pg.append(f.getImuState()["p_I_G"].copy()))
        t1 = time.time()
        synth.drive(o, wl, on_frame=lambda k, f: po.append(f.getImuState()["p_I_G"].copy()))
        t2 = time.time()
        d = np.array(pg) - np.array(po)
        log(f"traj200_{np.dtype(dtype).name}", {"rms_m": float(np.sqrt((d ** 2).sum(axis=1).mean())), "max_m": float(np.abs(d).max()),
                                                 "engine_fps": 200 / (t1 - t0), "oracle_fps": 200 / (t2 - t1), "n_updates": g.counters()["n_updates"]})
    # smoke case
    for dtype in (np.float64, np.float32):
        log(f"smoke_{np.dtype(dtype).name}", window(40, 12, 4, dtype, dict()))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
