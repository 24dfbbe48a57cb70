"""Shared helpers for the parity tests."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import synth  # noqa: E402
from msckf_mono_b200.cview import CFilter  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def rel(x, y):
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


def make_oracle(lib, dtype, faithful_max_rows=0, drop_null_rows=False):
    o = CFilter(lib, "msckf_oracle_", dtype)
    o.setOption(0, faithful_max_rows)
    o.setOption(1, 1.0 if drop_null_rows else 0.0)
    return o


def quat_err(q1, q2):
    """angle between unit quaternions (sign-insensitive)."""
    q1 = np.asarray(q1, float).reshape(-1, 4)
    q2 = np.asarray(q2, float).reshape(-1, 4)
    d = np.abs((q1 * q2).sum(axis=1)) / (np.linalg.norm(q1, axis=1) * np.linalg.norm(q2, axis=1))
    return float(np.max(2 * np.arccos(np.clip(d, -1, 1))))


def state_of(f):
    s = f.getImuState()
    cs = f.getCamStates()
    return {"imu_p": s["p_I_G"], "imu_v": s["v_I_G"], "imu_q": s["q_IG"], "imu_bg": s["b_g"], "imu_ba": s["b_a"],
            "P": f.getCovariance(), "cam_p": cs["p_C_G"], "cam_q": cs["q_CG"], "cam_ids": cs["state_id"],
            "cam_last_corr": cs["last_correlated_id"], "tracked_ids": f.getTrackedFeatureIds()}


def run_collect(filt, wl, prune_redundant=False):
    """drive a workload; collect per-frame accept/valid flags like tests/golden/make_golden.py"""
    rec = {"valid": [], "accepted": [], "ntracks": []}

    def on_frame(k, f):
        r = f.lastReport()
        rec["ntracks"].append(len(r["valid"]))
        rec["valid"].extend(int(v) for v in r["valid"])
        rec["accepted"].extend(int(v) for v in r["accepted"])

    synth.drive(filt, wl, on_frame=on_frame)
    return {k: np.array(v) for k, v in rec.items()}
