"""Parity cases shared by the GPU tests and by scripts/measure_parity.py.

Each case drives the B200 engine and the CPU oracle (or a committed oracle fixture) on the same seeded inputs, asserts what
must be bit-exact (flags, ids, counters) and RETURNS the floating-point differences as a dict of numbers.
scripts/measure_parity.py runs every case on the GPU box and writes the numbers to tests/golden/parity_measured.json;
the tests assert `value <= 10 x the committed measurement` (tests/parity_cases.py::check), so every asserted bound is a
measured one with one decade of headroom (no loose bounds), and a regression of more than 10x fails.
Test infrastructure only.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from msckf_mono_b200 import synth  # noqa: E402
from tests.common import GOLDEN, make_oracle, quat_err, rel, run_collect, state_of  # noqa: E402

MEASURED = GOLDEN / "parity_measured.json"
ORACLE = ROOT / "oracle" / "libmsckf_oracle.so"
# resolution floors of the measurements themselves (a measured 0.0 must not become a bound of 0):
# quaternion angles come from 2*acos(|q1.q2|), whose resolution is sqrt(eps)
FLOOR = {"dx": 1e-12, "P": 1e-13, "gamma": 1e-13, "pfg": 1e-13, "imu_p": 1e-13, "cam_p": 1e-13, "imu_q": 5e-7, "cam_q": 5e-7,
         "rms_m": 1e-12, "max_m": 1e-12, "map": 1e-13}
HEADROOM = 10.0
WINDOWS = [(3, 4, 5), (8, 6, 3), (40, 12, 4), (300, 30, 0)]


def make_engine(dtype, **kw):
    from msckf_mono_b200 import engine_filter
    return engine_filter(dtype, **kw)


def f32_inputs(*filters):
    """identical (float32-rounded) inputs for filters of different precision"""
    for f in filters:
        f._round = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)


def assert_bookkeeping_equal(g, o, rg, ro):
    assert np.array_equal(rg["ntracks"], ro["ntracks"])
    assert np.array_equal(rg["valid"], ro["valid"])
    assert np.array_equal(rg["accepted"], ro["accepted"])
    sg, so = state_of(g), state_of(o)
    for k in ("cam_ids", "cam_last_corr", "tracked_ids"):
        assert np.array_equal(sg[k], so[k]), k
    assert np.array_equal(g.getPrunedStates()["state_id"], o.getPrunedStates()["state_id"])
    cg, co = g.counters(), o.counters()
    for k in ("num_residualized", "pfg_shifted", "pfg_oob", "n_updates"):
        assert cg[k] == co[k], (k, cg, co)
    return sg, so


def numbers(g, o, sg, so):
    rep_g, rep_o = g.lastReport(), o.lastReport()
    dxg, dxo = g.lastDeltaX(), o.lastDeltaX()
    if len(dxg) != len(dxo):  # a prune after the last update changed the window: compare the part both still have
        k = min(len(dxg), len(dxo))
        dxg, dxo = dxg[:15], dxo[:15]
    ok = np.asarray(rep_o["valid"]).astype(bool)  # (a rejected track's position is whatever its failed triangulation left)
    acc = np.asarray(rep_o["accepted"]).astype(bool)  # (gamma of a track that failed the Cholesky is a sentinel)
    return {"dx": rel(dxg, dxo), "P": float(np.abs(sg["P"] - so["P"]).max() / np.abs(so["P"]).max()),
            "gamma": rel(np.asarray(rep_g["gamma"])[acc], np.asarray(rep_o["gamma"])[acc]) if acc.any() else 0.0,
            "pfg": rel(np.asarray(rep_g["p_f_G"])[ok], np.asarray(rep_o["p_f_G"])[ok]) if ok.any() else 0.0,
            "imu_p": float(np.abs(sg["imu_p"] - so["imu_p"]).max()), "cam_p": float(np.abs(sg["cam_p"] - so["cam_p"]).max()),
            "imu_q": quat_err(sg["imu_q"], so["imu_q"]), "cam_q": quat_err(sg["cam_q"], so["cam_q"]), "flips": 0,
            "rank_engine": int(g.counters()["rows_kept"]), "rank_oracle": int(o.counters()["rows_kept"])}


def window_case(nf, nc, seq, dtype, oracle_kw, eng_kw=None, **wkw):
    """engine vs oracle on one window workload: bookkeeping bit-exact (asserted here), numbers returned"""
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq, **wkw)
    g, o = make_engine(dtype, **(eng_kw or {})), make_oracle(ORACLE, dtype, **oracle_kw)
    if np.dtype(dtype) == np.float32:
        f32_inputs(g, o)
    rg, ro = run_collect(g, wl), run_collect(o, wl)
    sg, so = assert_bookkeeping_equal(g, o, rg, ro)
    return numbers(g, o, sg, so)


def golden_case(tag, mode, eng_kw=None):
    """engine vs committed oracle outputs at sizes where the oracle needs minutes (tests/golden/make_golden_stress.py)"""
    from tests.golden.make_golden_stress import STRESS_CASES, load_P
    nf, nc, seq, dtype, _ = STRESS_CASES[f"{tag}_{mode}"]
    gold = np.load(GOLDEN / f"{tag}_{mode}.npz")
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq)
    g = make_engine(np.dtype(dtype), **(eng_kw or dict(max_clones=nc + 4, max_tracks=nf + 48, max_obs=(nf + 48) * nc)))
    if np.dtype(dtype) == np.float32:
        f32_inputs(g)
    rg = run_collect(g, wl)
    sg, rep = state_of(g), g.lastReport()
    flips = int((rg["accepted"] != gold["all_accepted"]).sum() + (rg["valid"] != gold["all_valid"]).sum())
    assert flips == 0, flips
    Po = load_P(gold)
    return {"dx": rel(g.lastDeltaX(), gold["dx"]), "P": float(np.abs(sg["P"] - Po).max() / np.abs(Po).max()),
            "gamma": rel(rep["gamma"], gold["gamma"]), "pfg": rel(rep["p_f_G"], gold["p_f_G"]),
            "imu_p": float(np.abs(sg["imu_p"] - gold["imu_p"]).max()), "cam_p": float(np.abs(sg["cam_p"] - gold["cam_p"]).max()),
            "flips": flips, "rank_engine": int(g.counters()["rows_kept"]), "rank_oracle": int(gold["rows_kept"]), "m": int(g.counters()["m"])}


def rejections_case(dtype):
    wl = synth.make_window_workload(n_features=120, n_clones=10, seq=12)
    synth.corrupt_observations(wl, seed=3)
    wl["params"]["translation_threshold"] = 0.2  # between the features' orthogonal translations
    g, o = make_engine(dtype), make_oracle(ORACLE, dtype, drop_null_rows=(np.dtype(dtype) == np.float64))
    rg, ro = run_collect(g, wl), run_collect(o, wl)
    rep = g.lastReport()
    assert 0 < rep["cm_passed"].sum() < len(rep["cm_passed"])           # some checkMotion rejections ...
    assert 0 < rep["accepted"].sum() < rep["valid"].sum()              # ... some gate rejections
    sg, so = assert_bookkeeping_equal(g, o, rg, ro)
    assert g.counters()["pfg_shifted"] > 0
    assert np.array_equal(rep["cm_passed"], o.lastReport()["cm_passed"])
    return numbers(g, o, sg, so)


def stream_case(dtype):
    """E-sim: 150 frames of propagate / augment / update / addFeatures / marginalize / pruneEmptyStates."""
    wl = synth.make_stream_workload(n_frames=150, seq=7, max_features=40, max_track_length=14, max_cam_states=12)
    g, o = make_engine(dtype), make_oracle(ORACLE, dtype, drop_null_rows=(np.dtype(dtype) == np.float64))
    rg, ro = run_collect(g, wl), run_collect(o, wl)
    sg, so = assert_bookkeeping_equal(g, o, rg, ro)
    assert g.counters()["n_updates"] > 60
    for cam in range(g.getNumCamStates()):
        assert np.array_equal(g.getCamTrackedIds(cam), o.getCamTrackedIds(cam))
    return numbers(g, o, sg, so)


def prune_redundant_case(dtype):
    wl = synth.make_stream_workload(n_frames=90, seq=9, max_features=40, max_track_length=40, max_cam_states=21)
    wl["params"]["redundancy_angle_thresh"] = 0.2   # slow the keyframe criterion down so clones are actually declared redundant
    wl["params"]["redundancy_distance_thresh"] = 0.2
    g, o = make_engine(dtype), make_oracle(ORACLE, dtype, drop_null_rows=(np.dtype(dtype) == np.float64))
    rg, ro = run_collect(g, wl, prune_redundant=True), run_collect(o, wl, prune_redundant=True)
    sg, so = assert_bookkeeping_equal(g, o, rg, ro)
    assert len(g.getPrunedStates()["state_id"]) > 10
    return numbers(g, o, sg, so)


def trajectory_case(dtype):
    """north-star trajectory bar: RMS position difference engine vs oracle over a 200-frame E-sim stream"""
    wl = synth.make_stream_workload(n_frames=200, seq=8, max_features=60, max_track_length=20, max_cam_states=20)
    wl["noise"] = synth.euroc_noise(tuned=True)
    g, o = make_engine(dtype), make_oracle(ORACLE, dtype)
    pg, po = [], []
    synth.drive(g, wl, on_frame=lambda k, f: pg.append(f.getImuState()["p_I_G"].copy()))
    synth.drive(o, wl, on_frame=lambda k, f: po.append(f.getImuState()["p_I_G"].copy()))
    d = np.array(pg) - np.array(po)
    truth = np.array([wl["traj"].pos(fr["time"]) for fr in wl["frames"]])
    return {"rms_m": float(np.sqrt((d ** 2).sum(axis=1).mean())), "max_m": float(np.abs(d).max()),
            "rms_vs_truth_m": float(np.sqrt(((np.array(pg) - truth) ** 2).sum(axis=1).mean())), "n_updates": int(g.counters()["n_updates"])}


def l98_case():
    wl = synth.make_window_workload(n_features=24, n_clones=98, seq=11, imu_per_frame=2)  # 10 ms frames: everything stays in view
    g = make_engine(np.float64, max_clones=104, max_tracks=64, max_obs=64 * 98)
    o = make_oracle(ORACLE, np.float64, drop_null_rows=True)
    rg, ro = run_collect(g, wl), run_collect(o, wl)
    sg, so = assert_bookkeeping_equal(g, o, rg, ro)
    assert g.lastReport()["rows"].max() == 2 * 98 - 3
    return numbers(g, o, sg, so)


def fused_vs_separate_case(nf, nc, seq):
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq)
    a, b = make_engine(np.float64), make_engine(np.float64)
    synth.drive(a, wl, marginalize_last=False)
    synth.drive(b, wl, marginalize_last=False)
    b.setOption(103, 0.0)
    a.marginalize(); b.marginalize()
    assert a.counters()["rows_kept"] == b.counters()["rows_kept"]
    Pa, Pb = a.getCovariance(), b.getCovariance()
    return {"dx": rel(a.lastDeltaX(), b.lastDeltaX()), "P": float(np.abs(Pa - Pb).max() / np.abs(Pa).max())}


def gram_mma_vs_simt_case(nf, nc, seq):
    """the Gram products of the compression on the FP64 tensor-core path (DMMA, default) vs SIMT DFMA tiles (option 5 = 0)"""
    wl = synth.make_window_workload(n_features=nf, n_clones=nc, seq=seq)
    a, b = make_engine(np.float64), make_engine(np.float64)
    synth.drive(a, wl, marginalize_last=False)
    synth.drive(b, wl, marginalize_last=False)
    b.setOption(105, 0.0)
    a.marginalize(); b.marginalize()
    assert a.counters()["rows_kept"] == b.counters()["rows_kept"]
    assert np.array_equal(a.lastReport()["accepted"], b.lastReport()["accepted"])
    Pa, Pb = a.getCovariance(), b.getCovariance()
    return {"dx": rel(a.lastDeltaX(), b.lastDeltaX()), "P": float(np.abs(Pa - Pb).max() / np.abs(Pa).max())}


def order_invariance_case():
    from tests.test_oracle import _permuted
    wl = synth.make_window_workload(n_features=40, n_clones=12, seq=22, isotropic=True)
    perm = np.random.default_rng(22).permutation(40)
    a, b = make_engine(np.float64), make_engine(np.float64)
    ra, rb = run_collect(a, wl), run_collect(b, _permuted(wl, perm))
    assert ra["valid"].all() and np.array_equal(np.asarray(ra["accepted"])[perm], rb["accepted"])
    Pa, Pb = a.getCovariance(), b.getCovariance()
    return {"dx": rel(a.lastDeltaX(), b.lastDeltaX()), "P": float(np.abs(Pa - Pb).max() / np.abs(Pa).max()),
            "gamma": rel(np.asarray(a.lastReport()["gamma"])[perm], b.lastReport()["gamma"])}


def finish_case():
    wl = synth.make_stream_workload(n_frames=25, seq=14, max_features=30, max_track_length=40, max_cam_states=30)
    g, o = make_engine(np.float64), make_oracle(ORACLE, np.float64, drop_null_rows=True)
    synth.drive(g, wl)
    synth.drive(o, wl)
    g.finish()
    o.finish()
    rg, ro = g.lastReport(), o.lastReport()
    assert len(rg["valid"]) > 10 and np.array_equal(rg["valid"], ro["valid"]) and np.array_equal(rg["accepted"], ro["accepted"])
    assert len(g.getMap()) == len(o.getMap())
    return {"P": rel(g.getCovariance(), o.getCovariance()), "map": rel(g.getMap(), o.getMap())}


def golden_numpy_case(name):
    """engine vs the committed fixtures of the independent NumPy/LAPACK restatement (no C++ oracle involved)"""
    from tests.golden.make_golden import CASES, make_workload
    kind, kw, dtype, drop = CASES[name]
    gold = np.load(GOLDEN / f"{name}.npz")
    g = make_engine(np.dtype(dtype))
    rec = run_collect(g, make_workload(kind, kw))
    st = state_of(g)
    assert np.array_equal(rec["valid"], gold["all_valid"]) and np.array_equal(rec["accepted"], gold["all_accepted"])
    assert np.array_equal(st["cam_ids"], gold["cam_ids"]) and np.array_equal(st["tracked_ids"], gold["tracked_ids"])
    return {"P": float(np.abs(st["P"] - gold["P"]).max() / np.abs(gold["P"]).max()), "imu_p": float(np.abs(st["imu_p"] - gold["imu_p"]).max()),
            "cam_q": quat_err(st["cam_q"], gold["cam_q"])}


BIG = dict(max_clones=64, max_tracks=2048, max_obs=2048 * 60)
CASES = {}
for _nf, _nc, _seq in WINDOWS:
    CASES[f"f64_clean_{_nf}x{_nc}"] = (window_case, (_nf, _nc, _seq, np.float64, dict(drop_null_rows=True)), {})
    CASES[f"f64_faithful_{_nf}x{_nc}"] = (window_case, (_nf, _nc, _seq, np.float64, dict(faithful_max_rows=900)), {})
    CASES[f"f32_direct_{_nf}x{_nc}"] = (window_case, (_nf, _nc, _seq, np.float32, dict()), {})
CASES["f64_iso_40x12"] = (window_case, (40, 12, 4, np.float64, dict()), dict(isotropic=True))
CASES["smoke_f64"] = (window_case, (40, 12, 4, np.float64, dict()), {})
CASES["smoke_f32"] = (window_case, (40, 12, 4, np.float32, dict()), {})
for _mode in ("clean", "faithful"):
    CASES[f"stress_f64_500x60_{_mode}"] = (golden_case, ("stress_f64_500x60", _mode), {})
    CASES[f"stress_f64_2000x60_{_mode}"] = (golden_case, ("stress_f64_2000x60", _mode), {})
    CASES[f"configB_f32_300x30_{_mode}"] = (golden_case, ("configB_f32_300x30", _mode), {})
for _dt in (np.float64, np.float32):
    _n = np.dtype(_dt).name
    CASES[f"rejections_{_n}"] = (rejections_case, (_dt,), {})
    CASES[f"stream150_{_n}"] = (stream_case, (_dt,), {})
    CASES[f"prune_redundant_{_n}"] = (prune_redundant_case, (_dt,), {})
    CASES[f"traj200_{_n}"] = (trajectory_case, (_dt,), {})
CASES["l98_f64"] = (l98_case, (), {})
CASES["fused_vs_separate_40x12"] = (fused_vs_separate_case, (40, 12, 4), {})
CASES["fused_vs_separate_300x30"] = (fused_vs_separate_case, (300, 30, 0), {})
CASES["gram_mma_vs_simt_300x30"] = (gram_mma_vs_simt_case, (300, 30, 0), {})
CASES["order_invariance"] = (order_invariance_case, (), {})
CASES["finish"] = (finish_case, (), {})
for _g in ("win_f64_8x6_clean", "win_f64_40x12_clean", "win_f64_iso_40x12", "win_f64_3x4", "stream_f64_60"):
    CASES[f"numpy_{_g}"] = (golden_numpy_case, (_g,), {})


def run_case(name):
    fn, args, kw = CASES[name]
    return fn(*args, **kw)


_measured = None


def measured():
    global _measured
    if _measured is None:
        _measured = json.loads(MEASURED.read_text())
    return _measured


# Cases that chain 60-170 updates: the filter's own dynamics amplify last-bit differences, so two valid implementations (or two
# kernel variants of this engine) differ by a run-dependent amount; their floors are the level such chains reach in fp64
CHAINED = ("stream150_", "prune_redundant_", "traj200_", "numpy_stream", "finish")
CHAIN_FLOOR = {"dx": 1e-8, "P": 1e-9, "gamma": 1e-9, "pfg": 1e-9, "imu_p": 1e-9, "cam_p": 1e-9, "rms_m": 1e-9, "max_m": 1e-9, "map": 1e-9}


def check(name, got, fields=None):
    """assert every number of `got` against HEADROOM x the committed measurement of the same case"""
    ref = measured()[name]
    chained = name.startswith(CHAINED)
    msgs = []
    for k, v in got.items():
        if k not in FLOOR or (fields is not None and k not in fields):
            continue
        floor = max(FLOOR[k], CHAIN_FLOOR.get(k, 0.0)) if chained else FLOOR[k]
        bound = HEADROOM * max(float(ref[k]), floor)
        if not (v <= bound):
            msgs.append(f"{name}.{k}: {v:.3e} > {bound:.3e} (10 x measured {float(ref[k]):.3e})")
    assert not msgs, "; ".join(msgs)
    print(f"{name}: " + ", ".join(f"{k} {v:.2e}" for k, v in got.items() if k in FLOOR))
    return got
