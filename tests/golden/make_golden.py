"""Generates tests/golden/*.npz from the independent NumPy/LAPACK restatement
(oracle/msckf_numpy.py).  The reference itself cannot be executed here (no Eigen/Boost/ROS),
so these vectors pin the C++ oracle against a second, LAPACK-based restatement -- not against
an execution of the reference ("parity unpinned").

Run:  python tests/golden/make_golden.py
Inputs are regenerated deterministically at test time by msckf_mono_b200.synth (seeded), so
only outputs are stored.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.msckf_numpy import MSCKF  # noqa: E402
from msckf_mono_b200 import synth  # noqa: E402

CASES = {
    # name: (kind, kwargs, dtype, drop_null_rows)
    "win_f64_8x6": ("window", dict(n_features=8, n_clones=6, seq=3), "float64", False),
    "win_f64_8x6_clean": ("window", dict(n_features=8, n_clones=6, seq=3), "float64", True),
    "win_f64_40x12_clean": ("window", dict(n_features=40, n_clones=12, seq=4), "float64", True),
    "win_f64_iso_40x12": ("window", dict(n_features=40, n_clones=12, seq=4, isotropic=True), "float64", False),
    "win_f32_40x12": ("window", dict(n_features=40, n_clones=12, seq=4), "float32", False),
    "win_f64_3x4": ("window", dict(n_features=3, n_clones=4, seq=5), "float64", False),
    "stream_f64_60": ("stream", dict(n_frames=60, seq=6, max_features=30, max_track_length=12, max_cam_states=10), "float64", True),
    "stream_f32_60": ("stream", dict(n_frames=60, seq=6, max_features=30, max_track_length=12, max_cam_states=10), "float32", False),
}


def make_workload(kind, kw):
    return synth.make_window_workload(**kw) if kind == "window" else synth.make_stream_workload(**kw)


def snapshot(f):
    cs = f.cam_states
    out = {
        "imu_p": f.imu["p_I_G"], "imu_v": f.imu["v_I_G"], "imu_q": f.imu["q_IG"], "imu_bg": f.imu["b_g"],
        "imu_ba": f.imu["b_a"], "P": f.getCovariance(),
        "cam_p": np.array([c.p_C_G for c in cs]).reshape(-1, 3), "cam_q": np.array([c.q_CG for c in cs]).reshape(-1, 4),
        "cam_ids": np.array([c.state_id for c in cs], dtype=np.int64),
        "cam_last_corr": np.array([c.last_correlated_id for c in cs], dtype=np.int64),
        "tracked_ids": np.array(f.tracked_feature_ids, dtype=np.uint64),
        "pruned_ids": np.array([c.state_id for c in f.getPrunedStates()], dtype=np.int64),
        "num_residualized": np.int64(f.num_feature_tracks_residualized),
        "n_updates": np.int64(f.stats["updates"]),
    }
    return out


def run_case(kind, kw, dtype, drop):
    wl = make_workload(kind, kw)
    f = MSCKF(np.dtype(dtype), drop_null_rows=drop)
    per_frame = {"valid": [], "accepted": [], "ntracks": []}

    def on_frame(k, filt):
        rec = filt.last_marg
        if rec is None:
            per_frame["ntracks"].append(0)
            return
        per_frame["ntracks"].append(len(rec["valid"]))
        per_frame["valid"].extend(int(v) for v in rec["valid"])
        per_frame["accepted"].extend(int(v) for v in rec["accepted"])

    synth.drive(f, wl, on_frame=on_frame)
    out = snapshot(f)
    out["frame_ntracks"] = np.array(per_frame["ntracks"], dtype=np.int64)
    out["all_valid"] = np.array(per_frame["valid"], dtype=np.int8)
    out["all_accepted"] = np.array(per_frame["accepted"], dtype=np.int8)
    if f.last_update is not None:
        out["last_dx"] = f.last_update["deltaX"]
    if f.last_marg is not None:
        out["last_gamma"] = np.array([g if g is not None else np.nan for g in f.last_marg["gamma"]], dtype=np.float64)
        out["last_pfg"] = np.array([p if p is not None else [np.nan] * 3 for p in f.last_marg["p_f_G"]], dtype=np.float64)
    return out


if __name__ == "__main__":
    here = Path(__file__).resolve().parent
    for name, (kind, kw, dtype, drop) in CASES.items():
        out = run_case(kind, kw, dtype, drop)
        np.savez_compressed(here / f"{name}.npz", **{k: np.asarray(v) for k, v in out.items()})
        print(name, "updates", int(out["n_updates"]), "M", len(out["cam_ids"]), "accepted", int(out["all_accepted"].sum()),
              "/", len(out["all_accepted"]))
