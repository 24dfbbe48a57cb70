"""ASL / EuRoC "mav0" folders for the ROS-free player (`msckf_mono_b200/asl/asl_player.cpp`).

`write_mav0` dumps a synthetic stream workload (`synth.make_stream_workload`) in the on-disk layout the reference's
dataset front end reads (/root/reference/datasets/asl_readers.cpp: `imu0/data.csv`, `cam0/sensor.yaml`,
`state_groundtruth_estimate0/data.csv`), plus two files the reference does not have because its front end extracts
them from images / hard-codes them: `cam0/tracks.csv` (pre-extracted feature tracks, normalised coordinates) and
`msckf.yaml` (noise and filter parameters of datasets/asl_msckf.cpp:73-125).  A real EuRoC `mav0/` folder with a
`cam0/tracks.csv` produced by any tracker is played the same way.
"""
import json
import os
import subprocess

import numpy as np

from . import synth

_HERE = os.path.dirname(os.path.abspath(__file__))


def player_path():
    p = os.path.join(_HERE, "asl_player")
    if not os.path.exists(p):
        raise RuntimeError("asl_player is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return p


def _quat_xyzw_of(R):
    return synth.rot_to_quat(R)


def write_mav0(wl, root, imu_per_frame=10, dT=0.005, feature_cov=7.0, max_gn_cost_px=11.0):
    """Write workload `wl` under `root`/mav0.  Returns the path of the mav0 folder.

    Timestamps are integer nanoseconds: frame k at t0 + k * imu_per_frame * dT, the IMU readings of frame k at
    T_{k-1} + (j + 1) * dT (they all precede or coincide with the image they lead to, like the reference's synchroniser
    delivers them)."""
    mav0 = os.path.join(root, "mav0")
    for d in ("imu0", "cam0", "state_groundtruth_estimate0"):
        os.makedirs(os.path.join(mav0, d), exist_ok=True)
    cam, noise, params, traj = wl["camera"], wl["noise"], wl["params"], wl["traj"]
    f_u, f_v, c_u, c_v = cam["f_u"], cam["f_v"], cam["c_u"], cam["c_v"]
    assert abs(noise["u_var_prime"] - (feature_cov / f_u) ** 2) < 1e-18, "workload was not built with this feature_cov"
    assert abs(noise["v_var_prime"] - (feature_cov / f_v) ** 2) < 1e-18, "isotropic-noise workloads are not representable"
    assert abs(params["max_gn_cost_norm"] - (max_gn_cost_px / f_u) ** 2) < 1e-18
    dT_ns = int(round(dT * 1e9))
    fdt_ns = imu_per_frame * dT_ns
    t0_ns = int(round(wl["t0"] * 1e9))
    frames = wl["frames"]
    # camera: T_BS = pose of the camera in the body (IMU) frame
    x, y, z, w = cam["q_CI"]
    R_CI = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])
    T_BS = np.eye(4)
    T_BS[:3, :3] = R_CI.T
    T_BS[:3, 3] = cam["p_C_I"]
    with open(os.path.join(mav0, "cam0", "sensor.yaml"), "w") as f:
        f.write("# camera of a synthetic sequence (EuRoC layout)\nsensor_type: camera\ncomment: synthetic\nT_BS:\n  cols: 4\n  rows: 4\n")
        f.write("  data: [" + ", ".join(repr(float(v)) for v in T_BS.reshape(-1)) + "]\n")
        f.write("rate_hz: %r\nresolution: [%d, %d]\ncamera_model: pinhole\n" % (1e9 / fdt_ns, synth.RESOLUTION[0], synth.RESOLUTION[1]))
        f.write("intrinsics: [%r, %r, %r, %r]\n" % (float(f_u), float(f_v), float(c_u), float(c_v)))
        f.write("distortion_model: radial-tangential\ndistortion_coefficients: [0.0, 0.0, 0.0, 0.0]\n")
    Q, P0 = noise["Q_imu"], noise["initial_imu_covar"]
    with open(os.path.join(mav0, "msckf.yaml"), "w") as f:
        f.write("# noise and filter parameters (reference: datasets/asl_msckf.cpp:73-125)\n")
        kv = {
            "feature_cov": feature_cov, "w_var": Q[0, 0], "dbg_var": Q[3, 3], "a_var": Q[6, 6], "dba_var": Q[9, 9],
            "q_var_init": P0[0, 0], "bg_var_init": P0[3, 3], "v_var_init": P0[6, 6], "ba_var_init": P0[9, 9], "p_var_init": P0[12, 12],
            "max_gn_cost_norm": max_gn_cost_px, "min_rcond": params["min_rcond"], "translation_threshold": params["translation_threshold"],
            "redundancy_angle_thresh": params["redundancy_angle_thresh"], "redundancy_distance_thresh": params["redundancy_distance_thresh"],
        }
        for k, v in kv.items():
            f.write("%s: %r\n" % (k, float(v)))
        for k in ("min_track_length", "max_track_length", "max_cam_states"):
            f.write("%s: %d\n" % (k, int(params[k])))
        f.write("gravity: [%r, %r, %r]\n" % tuple(float(v) for v in wl["imu_state"]["g"]))
    with open(os.path.join(mav0, "imu0", "data.csv"), "w") as f:
        f.write("#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]\n")
        for k, fr in enumerate(frames):
            assert len(fr["imu"]) in (0, imu_per_frame)
            for j, (omega, a, dt) in enumerate(fr["imu"]):
                assert dt == dT
                t = t0_ns + (k - 1) * fdt_ns + (j + 1) * dT_ns
                f.write("%d,%s\n" % (t, ",".join(repr(float(v)) for v in list(omega) + list(a))))
    with open(os.path.join(mav0, "cam0", "data.csv"), "w") as f, open(os.path.join(mav0, "cam0", "tracks.csv"), "w") as g:
        f.write("#timestamp [ns],filename\n")
        g.write("#timestamp [ns],feature_id,x_normalised,y_normalised\n")
        for k, fr in enumerate(frames):
            t = t0_ns + k * fdt_ns
            f.write("%d,%d.png\n" % (t, t))
            for key in ("update", "add"):  # still-tracked features first (front-end order), then the new detections
                if fr[key] is None:
                    continue
                obs, ids = fr[key]
                for (ox, oy), i in zip(np.asarray(obs, float).reshape(-1, 2), ids):
                    g.write("%d,%d,%r,%r\n" % (t, int(i), float(ox), float(oy)))
    with open(os.path.join(mav0, "state_groundtruth_estimate0", "data.csv"), "w") as f:
        f.write("#timestamp,p_RS_R_x [m],p_RS_R_y [m],p_RS_R_z [m],q_RS_w [],q_RS_x [],q_RS_y [],q_RS_z [],v_RS_R_x [m s^-1],v_RS_R_y [m s^-1],"
                "v_RS_R_z [m s^-1],b_w_RS_S_x [rad s^-1],b_w_RS_S_y [rad s^-1],b_w_RS_S_z [rad s^-1],b_a_RS_S_x [m s^-2],b_a_RS_S_y [m s^-2],b_a_RS_S_z [m s^-2]\n")
        for k in range(len(frames)):
            t = t0_ns + k * fdt_ns
            ts = wl["t0"] + k * imu_per_frame * dT
            qx, qy, qz, qw = _quat_xyzw_of(traj.R_GI(ts))
            row = list(traj.pos(ts)) + [qw, qx, qy, qz] + list(traj.vel(ts)) + [0.0] * 6
            f.write("%d,%s\n" % (t, ",".join(repr(float(v)) for v in row)))
    return mav0


def run_player(mav0, dtype="f32", out=None, max_frames=None, prune_redundant=True, state_id="imu", dry_run=False, timeout=600):
    """Run the C++ player; returns its JSON summary (dict)."""
    cmd = [player_path(), "--mav0", mav0, "--dtype", dtype, "--prune-redundant", "1" if prune_redundant else "0", "--state-id", state_id]
    if out:
        cmd += ["--out", out]
    if max_frames is not None:
        cmd += ["--max-frames", str(int(max_frames))]
    if dry_run:
        cmd += ["--dry-run"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("asl_player failed (%d): %s" % (r.returncode, r.stderr.strip()))
    return json.loads(r.stdout.strip().splitlines()[-1])


def read_trajectory(path):
    """-> dict of arrays: t_ns, p (N x 3), q_wxyz (N x 4), n_clones."""
    a = np.loadtxt(path, delimiter=",", comments="#", ndmin=2)
    return {"t_ns": a[:, 0].astype(np.int64), "p": a[:, 1:4], "q_wxyz": a[:, 4:8], "n_clones": a[:, 8].astype(int)}
