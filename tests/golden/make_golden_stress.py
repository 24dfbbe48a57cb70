"""Generates tests/golden/stress_*.npz: outputs of the C++ CPU oracle (oracle/msckf_oracle.hpp) at BASELINE config-S
shapes (500 x 60 and the full 2000 x 60, fp64; 300 x 30 fp32 = config B), in both oracle modes (reference-literal thin-Q
compression and the exact-subspace mode `drop_null_rows`).  The oracle needs 1-5 minutes of CPU per case at these sizes,
so the GPU parity tests compare the engine with these committed vectors instead of re-running it on the GPU box.

Run:  python tests/golden/make_golden_stress.py [case ...]
Inputs are regenerated deterministically at test time by msckf_mono_b200.synth (seeded); only outputs are stored
(the covariance as float64 upper triangle).
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import synth  # noqa: E402
from tests.common import make_oracle, run_collect, state_of  # noqa: E402

STRESS_CASES = {
    # name: (n_features, n_clones, seq, dtype, drop_null_rows)
    "stress_f64_500x60_clean": (500, 60, 30, "float64", True),
    "stress_f64_500x60_faithful": (500, 60, 30, "float64", False),
    "stress_f64_2000x60_clean": (2000, 60, 30, "float64", True),
    "stress_f64_2000x60_faithful": (2000, 60, 30, "float64", False),
    "configB_f32_300x30_faithful": (300, 30, 0, "float32", False),
    "configB_f32_300x30_clean": (300, 30, 0, "float32", True),
}


def run_case(nf, nc, seq, dtype, drop):
    dtype = np.dtype(dtype)
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq)
    o = make_oracle(ROOT / "oracle" / "libmsckf_oracle.so", dtype, drop_null_rows=drop)
    if dtype == np.float32:
        o._round = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    rec = run_collect(o, wl)
    st = state_of(o)
    rep = o.lastReport()
    P = st["P"]
    iu = np.triu_indices(P.shape[0])
    cnt = o.counters()
    return {"all_valid": rec["valid"].astype(np.int8), "all_accepted": rec["accepted"].astype(np.int8), "P_triu": P[iu], "n": np.int64(P.shape[0]),
            "dx": o.lastDeltaX(), "gamma": rep["gamma"], "p_f_G": rep["p_f_G"], "imu_p": st["imu_p"], "imu_v": st["imu_v"], "imu_q": st["imu_q"],
            "imu_bg": st["imu_bg"], "imu_ba": st["imu_ba"], "cam_p": st["cam_p"], "cam_q": st["cam_q"], "rows_kept": np.int64(cnt["rows_kept"]),
            "m": np.int64(cnt["m"])}


def load_P(gold):
    n = int(gold["n"])
    P = np.zeros((n, n))
    iu = np.triu_indices(n)
    P[iu] = gold["P_triu"]
    return P + np.triu(P, 1).T


if __name__ == "__main__":
    here = Path(__file__).resolve().parent
    names = sys.argv[1:] or list(STRESS_CASES)
    for name in names:
        t0 = time.time()
        out = run_case(*STRESS_CASES[name])
        np.savez_compressed(here / f"{name}.npz", **{k: np.asarray(v) for k, v in out.items()})
        print(name, "accepted", int(out["all_accepted"].sum()), "/", len(out["all_accepted"]), "m", int(out["m"]), "rank", int(out["rows_kept"]),
              f"{time.time() - t0:.0f} s", flush=True)
