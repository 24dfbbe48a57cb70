"""NumPy/SciPy restatement of msckf_mono::MSCKF<_S>  --  TEST INFRASTRUCTURE ONLY.

This file is the *independent cross-check* of the C++ oracle (oracle/msckf_oracle.hpp).
It restates the reference algorithm (include/msckf_mono/msckf.h, all line numbers below
refer to that file unless another file is named) on top of LAPACK (scipy.linalg.qr for
Eigen's HouseholderQR / ColPivHouseholderQR, numpy.linalg for inverse/solve) so that the
two restatements do not share numerical kernels.  It is also the generator of the golden
fixtures in tests/golden/ (see tests/golden/make_golden.py).

PARITY UNPINNED: the reference ships no tests, golden vectors or recorded outputs, and it
cannot be compiled here (Eigen, Boost, ROS absent).  This restatement and the C++ oracle
pin each other; neither is pinned by an execution of the reference itself.

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import this.

Conventions
  * quaternions are stored as (x, y, z, w) (Eigen coeffs() order);
  * all arithmetic is carried out in `dtype` (np.float32 or np.float64) like the
    reference's template parameter _S.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
from scipy.stats import chi2


# ----------------------------------------------------------------------------- helpers
def skew(v):
    """matrix_utils.h:8-17 vectorToSkewSymmetric."""
    z = v.dtype.type(0)
    return np.array([[z, -v[2], v[1]], [v[2], z, -v[0]], [-v[1], v[0], z]], dtype=v.dtype)


def omega_mat(w):
    """matrix_utils.h:20-30 omegaMat."""
    O = np.zeros((4, 4), dtype=w.dtype)
    O[:3, :3] = -skew(w)
    O[:3, 3] = w
    O[3, :3] = -w
    return O


def quat_to_rot(q):
    """Eigen::Quaternion::toRotationMatrix (q = x,y,z,w)."""
    x, y, z, w = q
    one = q.dtype.type(1)
    two = q.dtype.type(2)
    tx, ty, tz = two * x, two * y, two * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array(
        [
            [one - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, one - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, one - (txx + tyy)],
        ],
        dtype=q.dtype,
    )


def quat_mul(a, b):
    """Eigen quaternion product a*b (Hamilton), storage x,y,z,w."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz,
        ],
        dtype=a.dtype,
    )


def quat_normalized(q):
    return (q / np.sqrt(np.dot(q, q))).astype(q.dtype)


def quat_inverse(q):
    n2 = np.dot(q, q)
    return np.array([-q[0], -q[1], -q[2], q[3]], dtype=q.dtype) / n2


def quat_rotate(q, v):
    """Eigen QuaternionBase::_transformVector: v + w*uv + vec x uv, uv = 2 vec x v."""
    vec = q[:3]
    uv = np.cross(vec, v)
    uv = uv + uv
    return (v + q[3] * uv + np.cross(vec, uv)).astype(q.dtype)


def quat_angular_distance(a, b):
    """Eigen 3.3 angularDistance: 2*atan2(|vec(a*conj(b))|, |w|)."""
    bc = np.array([-b[0], -b[1], -b[2], b[3]], dtype=b.dtype)
    d = quat_mul(a, bc)
    return a.dtype.type(2) * np.arctan2(np.sqrt(np.dot(d[:3], d[:3])), np.abs(d[3]))


def chi2_table(dtype):
    """msckf.h:91-95 : table[i-1] = quantile(chi_squared(i), 0.05), i = 1..99."""
    return chi2.ppf(0.05, np.arange(1, 100)).astype(dtype)


def expm_eigen(A):
    """unsupported/Eigen MatrixExponential (Pade + scaling/squaring), msckf.h:111.

    Degree selection thresholds follow Eigen's matrix_exp_computeUV<float/double>.
    """
    dt = A.dtype
    n = A.shape[0]
    I = np.eye(n, dtype=dt)
    l1 = np.abs(A).sum(axis=0).max()
    squarings = 0

    def pade(Am, b):
        deg = len(b) - 1
        A2 = Am @ Am
        pows = {0: I, 2: A2}
        for k in range(4, deg, 2):
            pows[k] = pows[k - 2] @ A2
        tmp = sum(dt.type(b[k + 1]) * pows[k] for k in range(0, deg, 2))
        U = Am @ tmp
        V = sum(dt.type(b[k]) * pows[k] for k in range(0, deg, 2))
        return U.astype(dt), V.astype(dt)

    b3 = [120.0, 60.0, 12.0, 1.0]
    b5 = [30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0]
    b7 = [17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0]
    b9 = [17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0,
          2162160.0, 110880.0, 3960.0, 90.0, 1.0]
    b13 = [64764752532480000.0, 32382376266240000.0, 7771770303897600.0,
           1187353796428800.0, 129060195264000.0, 10559470521600.0, 670442572800.0,
           33522128640.0, 1323241920.0, 40840800.0, 960960.0, 16380.0, 182.0, 1.0]
    if dt == np.float32:
        if l1 < 4.258730016922831e-001:
            U, V = pade(A, b3)
        elif l1 < 1.880152677804762e+000:
            U, V = pade(A, b5)
        else:
            maxnorm = 3.925724783138660
            _, squarings = np.frexp(l1 / maxnorm)
            squarings = max(int(squarings), 0)
            U, V = pade((A / dt.type(2.0 ** squarings)).astype(dt), b7)
    else:
        if l1 < 1.495585217958292e-002:
            U, V = pade(A, b3)
        elif l1 < 2.539398330063230e-001:
            U, V = pade(A, b5)
        elif l1 < 9.504178996162932e-001:
            U, V = pade(A, b7)
        elif l1 < 2.097847961257068e+000:
            U, V = pade(A, b9)
        else:
            maxnorm = 5.371920351148152
            _, squarings = np.frexp(l1 / maxnorm)
            squarings = max(int(squarings), 0)
            As = (A / dt.type(2.0 ** squarings)).astype(dt)
            # pade13 (Eigen evaluates it in a factored form; same polynomial)
            A2 = As @ As
            A4 = A2 @ A2
            A6 = A4 @ A2
            b = [dt.type(x) for x in b13]
            V = b[13] * A6 + b[11] * A4 + b[9] * A2
            tmp = A6 @ V
            tmp = tmp + b[7] * A6 + b[5] * A4 + b[3] * A2 + b[1] * I
            U = As @ tmp
            tmp = b[12] * A6 + b[10] * A4 + b[8] * A2
            V = A6 @ tmp
            V = V + b[6] * A6 + b[4] * A4 + b[2] * A2 + b[0] * I
    R = np.linalg.solve((-U + V).astype(dt), (U + V).astype(dt)).astype(dt)
    for _ in range(squarings):
        R = (R @ R).astype(dt)
    return R


# ----------------------------------------------------------------------------- types
class CamState:
    """types.h:58-68 camState."""

    def __init__(self, p, q, time, state_id):
        self.p_C_G = p
        self.q_CG = q
        self.time = time
        self.state_id = int(state_id)
        self.last_correlated_id = -1
        self.tracked_feature_ids = []

    def copy(self):
        c = CamState(self.p_C_G.copy(), self.q_CG.copy(), self.time, self.state_id)
        c.last_correlated_id = self.last_correlated_id
        c.tracked_feature_ids = list(self.tracked_feature_ids)
        return c


class FeatureTrack:
    """types.h:115-126 featureTrack."""

    def __init__(self, fid):
        self.feature_id = int(fid)
        self.observations = []
        self.cam_state_indices = []
        self.initialized = False
        self.p_f_G = None


class TrackToResidualize:
    """types.h:101-113 featureTrackToResidualize."""

    def __init__(self):
        self.feature_id = 0
        self.observations = []
        self.cam_states = []
        self.cam_state_indices = []
        self.initialized = False
        self.p_f_G = None


class MSCKF:
    """Restatement of msckf_mono::MSCKF<_S> (msckf.h:31-1512)."""

    def __init__(self, dtype=np.float64, faithful_max_rows=0, drop_null_rows=False, null_row_tol=1e-9):
        self.dt = np.dtype(dtype)
        self.S = self.dt.type
        # drop_null_rows: also drop rows of T_H (index >= 15) whose norm is <= null_row_tol * max row
        # norm -- the rounding-defined rows of numerically zero pivots.  OFF = reference.
        self.drop_null_rows = drop_null_rows
        self.null_row_tol = null_row_tol
        # when m <= faithful_max_rows, measurementUpdate materialises the dense m x m Q and
        # R_o exactly as msckf.h:406,1344,1366; otherwise the thin/block-diagonal form.
        self.faithful_max_rows = faithful_max_rows
        self.stats = {"pfg_shifted": 0, "pfg_oob": 0, "updates": 0}
        self.last_update = None  # debug record of the last measurementUpdate

    # ------------------------------------------------------------------ msckf.h:72-97
    def initialize(self, camera, noise, params, imu_state):
        dt = self.dt
        self.camera = {k: np.asarray(v, dtype=dt) for k, v in camera.items()}
        self.noise = {
            "u_var_prime": self.S(noise["u_var_prime"]),
            "v_var_prime": self.S(noise["v_var_prime"]),
            "Q_imu": np.asarray(noise["Q_imu"], dtype=dt).reshape(12, 12),
            "initial_imu_covar": np.asarray(noise["initial_imu_covar"], dtype=dt).reshape(15, 15),
        }
        self.params = dict(params)
        for k in ("max_gn_cost_norm", "min_rcond", "translation_threshold",
                  "redundancy_angle_thresh", "redundancy_distance_thresh"):
            self.params[k] = self.S(self.params[k])
        self.num_feature_tracks_residualized = 0
        st = {k: np.asarray(v, dtype=dt).copy() for k, v in imu_state.items()}
        st["p_I_G_null"] = st["p_I_G"].copy()
        st["v_I_G_null"] = st["v_I_G"].copy()
        st["q_IG_null"] = st["q_IG"].copy()
        self.imu = st
        self.imu_covar = self.noise["initial_imu_covar"].copy()
        self.cam_covar = np.zeros((0, 0), dtype=dt)
        self.imu_cam_covar = np.zeros((15, 0), dtype=dt)
        self.chi_table = chi2_table(dt)
        self.feature_tracks = []
        self.tracked_feature_ids = []
        self.feature_tracks_to_residualize = []
        self.tracks_to_remove = []
        self.cam_states = []
        self.pruned_states = []
        self.map = []

    # ------------------------------------------------------------------ msckf.h:101-145
    def propagate(self, omega, a, dT):
        dt, S = self.dt, self.S
        omega = np.asarray(omega, dtype=dt)
        a = np.asarray(a, dtype=dt)
        dT = S(dT)
        imu = self.imu
        # calcF :874-890
        F = np.zeros((15, 15), dtype=dt)
        omegaHat = omega - imu["b_g"]
        aHat = a - imu["b_a"]
        C_IG = quat_to_rot(imu["q_IG"])
        F[0:3, 0:3] = -skew(omegaHat)
        F[0:3, 3:6] = -np.eye(3, dtype=dt)
        F[6:9, 0:3] = -C_IG.T @ skew(aHat)
        F[6:9, 9:12] = -C_IG.T
        F[12:15, 6:9] = np.eye(3, dtype=dt)
        # calcG :892-903
        G = np.zeros((15, 12), dtype=dt)
        G[0:3, 0:3] = -np.eye(3, dtype=dt)
        G[3:6, 3:6] = np.eye(3, dtype=dt)
        G[6:9, 6:9] = -C_IG.T
        G[9:12, 9:12] = np.eye(3, dtype=dt)

        prop = self._propagate_imu_state_rk(omega, a, dT)
        F = (F * dT).astype(dt)
        Phi = expm_eigen(F)

        R_kk_1 = quat_to_rot(imu["q_IG_null"])
        Phi[0:3, 0:3] = quat_to_rot(prop["q_IG"]) @ R_kk_1.T
        u = R_kk_1 @ imu["g"]
        s = (u / np.dot(u, u)).astype(dt)  # row vector (u^T u)^-1 u^T
        A1 = Phi[6:9, 0:3].copy()
        tmp = imu["v_I_G_null"] - prop["v_I_G"]
        w1 = skew(tmp) @ imu["g"]
        Phi[6:9, 0:3] = A1 - np.outer(A1 @ u - w1, s)
        A2 = Phi[12:15, 0:3].copy()
        tmp = dT * imu["v_I_G_null"] + imu["p_I_G_null"] - prop["p_I_G"]
        w2 = skew(tmp) @ imu["g"]
        Phi[12:15, 0:3] = A2 - np.outer(A2 @ u - w2, s)
        Phi = Phi.astype(dt)

        imu_covar_prop = Phi @ (self.imu_covar + G @ self.noise["Q_imu"] @ G.T * dT) @ Phi.T
        for k in ("p_I_G", "v_I_G", "q_IG"):
            imu[k] = prop[k]
        imu["q_IG_null"] = imu["q_IG"].copy()
        imu["v_I_G_null"] = imu["v_I_G"].copy()
        imu["p_I_G_null"] = imu["p_I_G"].copy()
        self.imu_covar = ((imu_covar_prop + imu_covar_prop.T) / S(2.0)).astype(dt)
        self.imu_cam_covar = (Phi @ self.imu_cam_covar).astype(dt)

    def _propagate_imu_state_rk(self, omega, a, dT):
        """msckf.h:1425-1467."""
        dt, S = self.dt, self.S
        imu = self.imu
        omega_vec = omega - imu["b_g"]
        omega_psi = (S(0.5) * omega_mat(omega_vec)).astype(dt)
        q = imu["q_IG"]
        y0 = np.array([-q[0], -q[1], -q[2], q[3]], dtype=dt)
        k0 = omega_psi @ y0
        k1 = omega_psi @ (y0 + (k0 / S(4.0)) * dT)
        k2 = omega_psi @ (y0 + (k0 / S(8.0) + k1 / S(8.0)) * dT)
        k3 = omega_psi @ (y0 + (-k1 / S(2.0) + k2) * dT)
        k4 = omega_psi @ (y0 + (k0 * S(3.0) / S(16.0) + k3 * S(9.0) / S(16.0)) * dT)
        k5 = omega_psi @ (
            y0
            + (-k0 * S(3.0) / S(7.0) + k1 * S(2.0) / S(7.0) + k2 * S(12.0) / S(7.0)
               - k3 * S(12.0) / S(7.0) + k4 * S(8.0) / S(7.0)) * dT
        )
        y_t = y0 + (S(7.0) * k0 + S(32.0) * k2 + S(12.0) * k3 + S(32.0) * k4 + S(7.0) * k5) * dT / S(90.0)
        qn = quat_normalized(np.array([-y_t[0], -y_t[1], -y_t[2], y_t[3]], dtype=dt))
        delta_v = (quat_to_rot(q).T @ (a - imu["b_a"]) + imu["g"]) * dT
        return {
            "q_IG": qn,
            "v_I_G": (imu["v_I_G"] + delta_v).astype(dt),
            "p_I_G": (imu["p_I_G"] + imu["v_I_G"] * dT).astype(dt),
        }

    # ------------------------------------------------------------------ covariance helpers
    def _assemble_P(self):
        """msckf.h:166-174 / :1104-1110 / :1330-1336."""
        c = self.cam_covar.shape[0]
        P = np.zeros((15 + c, 15 + c), dtype=self.dt)
        P[:15, :15] = self.imu_covar
        if c:
            P[:15, 15:] = self.imu_cam_covar
            P[15:, :15] = self.imu_cam_covar.T
            P[15:, 15:] = self.cam_covar
        return P

    def _split_P(self, P):
        self.imu_covar = P[:15, :15].copy()
        self.cam_covar = P[15:, 15:].copy()
        self.imu_cam_covar = P[:15, 15:].copy()

    # ------------------------------------------------------------------ msckf.h:148-212
    def augmentState(self, state_id, time):
        dt, S = self.dt, self.S
        self.map = []
        imu, cam = self.imu, self.camera
        q_CG = quat_normalized(quat_mul(cam["q_CI"], imu["q_IG"]))
        p_C_G = imu["p_I_G"] + quat_rotate(quat_inverse(imu["q_IG"]), cam["p_C_I"])
        cs = CamState(p_C_G.astype(dt), q_CG, S(time), state_id)
        P = self._assemble_P()
        M = len(self.cam_states)
        n = 15 + 6 * M
        J = np.zeros((6, n), dtype=dt)
        J[0:3, 0:3] = quat_to_rot(cam["q_CI"])
        J[3:6, 0:3] = skew(quat_rotate(quat_inverse(imu["q_IG"]), cam["p_C_I"]))
        J[3:6, 12:15] = np.eye(3, dtype=dt)
        T = np.zeros((n + 6, n), dtype=dt)
        T[:n, :n] = np.eye(n, dtype=dt)
        T[n:, :] = J
        P_aug = T @ P @ T.T
        P_aug = ((P_aug + P_aug.T) / S(2.0)).astype(dt)
        self.cam_states.append(cs)
        self._split_P(P_aug)

    # ------------------------------------------------------------------ msckf.h:215-299
    def update(self, measurements, feature_ids):
        feature_ids = [int(i) for i in feature_ids]
        self.feature_tracks_to_residualize = []
        self.tracks_to_remove = []
        for id_iter, fid in enumerate(list(self.tracked_feature_ids)):
            is_valid = fid in feature_ids
            track = self.feature_tracks[id_iter]
            if is_valid:
                k = feature_ids.index(fid)
                track.observations.append(np.asarray(measurements[k], dtype=self.dt))
                last = self.cam_states[-1]
                last.tracked_feature_ids.append(fid)
                track.cam_state_indices.append(last.state_id)
            if (not is_valid) or len(track.observations) >= self.params["max_track_length"]:
                ttr = TrackToResidualize()
                ttr.cam_states, ttr.cam_state_indices = self._remove_tracked_feature(fid)
                if len(ttr.cam_states) >= self.params["min_track_length"]:
                    ttr.feature_id = track.feature_id
                    ttr.observations = [o.copy() for o in track.observations]
                    ttr.initialized = track.initialized
                    if track.initialized:
                        ttr.p_f_G = track.p_f_G.copy()
                    self.feature_tracks_to_residualize.append(ttr)
                self.tracks_to_remove.append(fid)
        for fid in self.tracks_to_remove:
            for ti, tr in enumerate(self.feature_tracks):
                if tr.feature_id == fid:
                    last_id = tr.cam_state_indices[-1]
                    for index in tr.cam_state_indices:
                        for cs in self.cam_states:
                            if (not cs.tracked_feature_ids) and cs.state_id == index:
                                cs.last_correlated_id = last_id
                    del self.feature_tracks[ti]
                    break
            if fid in self.tracked_feature_ids:
                self.tracked_feature_ids.remove(fid)

    def _remove_tracked_feature(self, fid):
        """msckf.h:1469-1485."""
        states, idx = [], []
        for c_i, cs in enumerate(self.cam_states):
            if fid in cs.tracked_feature_ids:
                cs.tracked_feature_ids.remove(fid)
                idx.append(c_i)
                states.append(cs.copy())
        return states, idx

    # ------------------------------------------------------------------ msckf.h:302-332
    def addFeatures(self, features, feature_ids):
        for i in range(len(features)):
            fid = int(feature_ids[i])
            if fid not in self.tracked_feature_ids:
                tr = FeatureTrack(fid)
                tr.observations.append(np.asarray(features[i], dtype=self.dt))
                last = self.cam_states[-1]
                last.tracked_feature_ids.append(fid)
                tr.cam_state_indices.append(last.state_id)
                self.feature_tracks.append(tr)
                self.tracked_feature_ids.append(fid)
            else:
                return  # :327-330 prints and drops the rest

    # ------------------------------------------------------------------ msckf.h:336-449
    def marginalize(self):
        dt, S = self.dt, self.S
        tracks = self.feature_tracks_to_residualize
        self.last_marg = None
        if not tracks:
            return
        valid_tracks, p_f_G_vec, own_pos = [], [], []
        total_nObs = 0
        num_passed = 0
        cm_flags, tri_flags = [], []
        for tr in tracks:
            if self.num_feature_tracks_residualized > 3 and not self._check_motion(
                tr.observations[0], tr.cam_states
            ):
                valid_tracks.append(False)
                own_pos.append(-1)
                cm_flags.append(False)
                tri_flags.append(False)
                continue
            cm_flags.append(True)
            isvalid, p_f_G = self._initialize_position(tr.cam_states, tr.observations)
            tri_flags.append(bool(isvalid))
            if isvalid:
                tr.initialized = True
                tr.p_f_G = p_f_G
                self.map.append(p_f_G)
            own_pos.append(len(p_f_G_vec))
            p_f_G_vec.append(p_f_G)
            if not isvalid:
                valid_tracks.append(False)
            else:
                num_passed += 1
                valid_tracks.append(True)
                total_nObs += len(tr.observations)
                self.num_feature_tracks_residualized += 1
        rec = {"valid": list(valid_tracks), "p_f_G": [None] * len(tracks), "accepted": [False] * len(tracks),
               "gamma": [None] * len(tracks), "rows": [0] * len(tracks)}
        for i, tr in enumerate(tracks):
            if own_pos[i] >= 0:
                rec["p_f_G"][i] = p_f_G_vec[own_pos[i]]
        self.last_marg = rec
        if not num_passed:
            return
        n = 15 + 6 * len(self.cam_states)
        H_blocks, r_blocks, Ro_blocks = [], [], []
        u_var, v_var = self.noise["u_var_prime"], self.noise["v_var_prime"]
        for it, tr in enumerate(tracks):
            if not valid_tracks[it]:
                continue
            # :419 indexes p_f_G_vec with the TRACK index (reference bug)
            if it < len(p_f_G_vec):
                if it != own_pos[it]:
                    self.stats["pfg_shifted"] += 1
                p_f_G = p_f_G_vec[it]
            else:
                self.stats["pfg_oob"] += 1
                p_f_G = p_f_G_vec[own_pos[it]]
            r_j = self._calc_residual(p_f_G, tr.cam_states, tr.observations)
            nObs = len(tr.observations)
            R_diag = np.tile(np.array([u_var, v_var], dtype=dt), nObs)
            H_o_j, A_j = self._calc_meas_jacobian(p_f_G, tr.cam_state_indices)
            r_o_j = (A_j.T @ r_j).astype(dt)
            R_o_j = (A_j.T @ (R_diag[:, None] * A_j)).astype(dt)
            ok, gamma = self._gating_test(H_o_j, r_o_j, len(tr.cam_states) - 1)
            rec["gamma"][it] = gamma
            if ok:
                rec["accepted"][it] = True
                rec["rows"][it] = H_o_j.shape[0]
                H_blocks.append(H_o_j)
                r_blocks.append(r_o_j)
                Ro_blocks.append(R_o_j)
        if H_blocks:
            H_o = np.vstack(H_blocks)
            r_o = np.concatenate(r_blocks)
        else:
            H_o = np.zeros((0, n), dtype=dt)
            r_o = np.zeros((0,), dtype=dt)
        self._measurement_update(H_o, r_o, Ro_blocks)

    # ------------------------------------------------------------------ msckf.h:905-958
    def _calc_meas_jacobian(self, p_f_G, cam_state_indices):
        dt, S = self.dt, self.S
        L = len(cam_state_indices)
        n = 15 + 6 * len(self.cam_states)
        H_f_j = np.zeros((2 * L, 3), dtype=dt)
        H_x_j = np.zeros((2 * L, n), dtype=dt)
        g = self.imu["g"]
        for c_i, index in enumerate(cam_state_indices):
            cs = self.cam_states[index]
            C = quat_to_rot(cs.q_CG)
            p_f_C = C @ (p_f_G - cs.p_C_G)
            X, Y, Z = p_f_C
            J_i = np.array([[1, 0, -X / Z], [0, 1, -Y / Z]], dtype=dt) * (S(1) / Z)
            A = np.hstack([J_i @ skew(p_f_C), -J_i @ C]).astype(dt)
            u = np.zeros(6, dtype=dt)
            u[:3] = C @ g
            tmp = p_f_G - cs.p_C_G
            u[3:] = skew(tmp) @ g
            H_x = A - np.outer(A @ u, u) * (S(1) / np.dot(u, u))
            H_x = H_x.astype(dt)
            H_f_j[2 * c_i:2 * c_i + 2, :] = -H_x[:, 3:6]
            H_x_j[2 * c_i:2 * c_i + 2, 15 + 6 * index:21 + 6 * index] = H_x
        A_j = self.left_nullspace(H_f_j)
        H_o_j = (A_j.T @ H_x_j).astype(dt)
        return H_o_j, A_j

    @staticmethod
    def left_nullspace(H_f_j):
        """msckf.h:954-955 : last 2L-3 columns of JacobiSVD full U.

        Eigen's JacobiSVD (default ColPivHouseholderQR preconditioner, rows > cols) forms full U
        as householderQ() of the column-pivoted QR times blockdiag(U_3x3, I): the trailing
        2L-3 columns are those of the Householder Q itself.  LAPACK geqp3 uses the same
        reflector convention (beta = -sign(alpha)*norm) and the same pivot rule (largest
        remaining column norm).
        """
        Q, _, _ = sla.qr(H_f_j, mode="full", pivoting=True)
        return Q[:, 3:].astype(H_f_j.dtype)

    # ------------------------------------------------------------------ msckf.h:960-978
    def _calc_residual(self, p_f_G, cam_states, observations):
        r = np.empty(2 * len(cam_states), dtype=self.dt)
        for i, cs in enumerate(cam_states):
            p_f_C = quat_to_rot(cs.q_CG) @ (p_f_G - cs.p_C_G)
            zhat = p_f_C[:2] / p_f_C[2]
            r[2 * i:2 * i + 2] = observations[i] - zhat
        return r

    # ------------------------------------------------------------------ msckf.h:980-1025
    def _check_motion(self, first_observation, cam_states):
        dt, S = self.dt, self.S
        if len(cam_states) < 2:
            return False
        first = cam_states[0]
        R0 = quat_to_rot(first.q_CG).T
        d = np.array([first_observation[0], first_observation[1], 1.0], dtype=dt)
        d = d / np.sqrt(np.dot(d, d))
        d = (R0 @ d).astype(dt)
        max_ortho = S(0)
        for cs in cam_states[1:]:
            t = cs.p_C_G - first.p_C_G
            par = np.dot(t, d)
            ortho = t - par * d
            nrm = np.sqrt(np.dot(ortho, ortho))
            if nrm > max_ortho:
                max_ortho = nrm
        return bool(max_ortho > self.params["translation_threshold"])

    # ------------------------------------------------------------------ msckf.h:1147-1285
    def _initialize_position(self, cam_states, measurements):
        dt, S = self.dt, self.S
        L = len(cam_states)
        # cam pose i : camera -> world ; then T_i <- T_i^-1 * T_0  (:1154-1168)
        R_w = [quat_to_rot(c.q_CG).T for c in cam_states]
        t_w = [c.p_C_G for c in cam_states]
        R0, t0 = R_w[0], t_w[0]
        Rs = np.empty((L, 3, 3), dtype=dt)
        ts = np.empty((L, 3), dtype=dt)
        for i in range(L):
            Rinv = R_w[i].T
            tinv = -(Rinv @ t_w[i])
            Rs[i] = Rinv @ R0
            ts[i] = Rinv @ t0 + tinv
        z = np.asarray(measurements, dtype=dt).reshape(L, 2)
        # generateInitialGuess :1126-1145 (first and last)
        Rl, tl = Rs[-1], ts[-1]
        z1, z2 = z[0], z[-1]
        m = Rl @ np.array([z1[0], z1[1], 1.0], dtype=dt)
        A = np.array([m[0] - z2[0] * m[2], m[1] - z2[1] * m[2]], dtype=dt)
        b = np.array([z2[0] * tl[2] - tl[0], z2[1] * tl[2] - tl[1]], dtype=dt)
        depth = (S(1) / np.dot(A, A)) * np.dot(A, b)
        init = np.array([z1[0] * depth, z1[1] * depth, depth], dtype=dt)
        sol = np.array([init[0] / init[2], init[1] / init[2], S(1.0) / init[2]], dtype=dt)

        def cost(x):
            h = Rs @ np.array([x[0], x[1], 1.0], dtype=dt) + x[2] * ts
            zh = h[:, :2] / h[:, 2:3]
            e = ((zh - z) ** 2).sum(axis=1)
            tot = S(0)
            for v in e:  # sequential accumulation like :1190-1194
                tot = tot + v
            return tot

        lam = S(1e-3)
        total_cost = cost(sol)
        outer = 0
        inner = 0
        is_cost_reduced = False
        delta_norm = S(0)
        I3 = np.eye(3, dtype=dt)
        while True:
            A3 = np.zeros((3, 3), dtype=dt)
            b3 = np.zeros(3, dtype=dt)
            for i in range(L):
                h = Rs[i] @ np.array([sol[0], sol[1], 1.0], dtype=dt) + sol[2] * ts[i]
                W = np.empty((3, 3), dtype=dt)
                W[:, :2] = Rs[i][:, :2]
                W[:, 2] = ts[i]
                J = np.empty((2, 3), dtype=dt)
                J[0] = S(1) / h[2] * W[0] - h[0] / (h[2] * h[2]) * W[2]
                J[1] = S(1) / h[2] * W[1] - h[1] / (h[2] * h[2]) * W[2]
                r = np.array([h[0] / h[2], h[1] / h[2]], dtype=dt) - z[i]
                e = np.sqrt(np.dot(r, r))
                w = S(1.0) if e <= S(0.01) else S(0.01) / (S(2) * e)
                if w == 1:
                    A3 += J.T @ J
                    b3 += J.T @ r
                else:
                    w2 = w * w
                    A3 += w2 * (J.T @ J)
                    b3 += w2 * (J.T @ r)
            while True:
                delta = np.linalg.solve((A3 + lam * I3).astype(dt), b3).astype(dt)
                new_sol = (sol - delta).astype(dt)
                delta_norm = np.sqrt(np.dot(delta, delta))
                new_cost = cost(new_sol)
                if new_cost < total_cost:
                    is_cost_reduced = True
                    sol = new_sol
                    total_cost = new_cost
                    lam = S(lam / 10) if float(lam / 10) > 1e-10 else S(1e-10)
                else:
                    is_cost_reduced = False
                    lam = S(lam * 10) if float(lam * 10) < 1e12 else S(1e12)
                cont = (inner < 10) and (not is_cost_reduced)
                inner += 1
                if not cont:
                    break
            inner = 0
            cont = (outer < 10) and (delta_norm > S(5e-7))
            outer += 1
            if not cont:
                break
        final = np.array([sol[0] / sol[2], sol[1] / sol[2], S(1.0) / sol[2]], dtype=dt)
        valid = True
        for i in range(L):
            pos = Rs[i] @ final + ts[i]
            if pos[2] <= 0:
                valid = False
                break
        normalized_cost = total_cost / S(2 * L * L)
        if normalized_cost > self.params["max_gn_cost_norm"]:
            valid = False
        p_f_G = (R0 @ final + t0).astype(dt)
        return valid, p_f_G

    # ------------------------------------------------------------------ msckf.h:1103-1124
    def _gating_test(self, H, r, dof):
        P = self._assemble_P()
        P1 = H @ P @ H.T
        P2 = self.noise["u_var_prime"] * np.eye(H.shape[0], dtype=self.dt)
        gamma = np.dot(r, np.linalg.solve((P1 + P2).astype(self.dt), r))
        return bool(gamma < self.chi_table[dof + 1]), float(gamma)

    # ------------------------------------------------------------------ msckf.h:1325-1423
    def _measurement_update(self, H_o, r_o, Ro_blocks):
        dt, S = self.dt, self.S
        m = r_o.shape[0]
        if m == 0:
            return
        P = self._assemble_P()
        n = P.shape[0]
        faithful = m <= self.faithful_max_rows
        if faithful:
            Q, R = sla.qr(H_o, mode="full")  # Q m x m, R m x n  (:1343-1345)
            R_o = sla.block_diag(*Ro_blocks).astype(dt)
        else:
            Q, R = sla.qr(H_o, mode="economic")  # rows >= n of the full R are zero => dropped
        R = np.triu(R)
        keep = np.any(R != 0, axis=1)  # :1347-1348
        if self.drop_null_rows:
            rn = np.sqrt((R.astype(np.float64) ** 2).sum(axis=1))
            keep &= (np.arange(R.shape[0]) < 15) | (rn > self.null_row_tol * rn.max())
        T_H = R[keep]
        Q_1 = Q[:, : R.shape[0]][:, keep]
        r_n = (Q_1.T @ r_o).astype(dt)
        if faithful:
            R_n = (Q_1.T @ R_o @ Q_1).astype(dt)
        else:
            R_n = np.zeros((Q_1.shape[1], Q_1.shape[1]), dtype=dt)
            off = 0
            for Rb in Ro_blocks:
                k = Rb.shape[0]
                Qb = Q_1[off:off + k]
                R_n += Qb.T @ Rb @ Qb
                off += k
            R_n = R_n.astype(dt)
        temp = (T_H @ P @ T_H.T + R_n).astype(dt)
        K = ((P @ T_H.T) @ np.linalg.inv(temp)).astype(dt)
        deltaX = (K @ r_n).astype(dt)
        self.last_update = {"T_H": T_H, "Q_1": Q_1, "r_n": r_n, "R_n": R_n, "S": temp, "K": K,
                            "deltaX": deltaX, "H_o": H_o, "r_o": r_o, "P": P}
        self.stats["updates"] += 1
        imu = self.imu
        imu["q_IG"] = quat_mul(self._build_update_quat(deltaX[0:3]), imu["q_IG"])  # not renormalised
        imu["b_g"] = imu["b_g"] + deltaX[3:6]
        imu["b_a"] = imu["b_a"] + deltaX[9:12]
        imu["v_I_G"] = imu["v_I_G"] + deltaX[6:9]
        imu["p_I_G"] = imu["p_I_G"] + deltaX[12:15]
        for c_i, cs in enumerate(self.cam_states):
            dq = self._build_update_quat(deltaX[15 + 6 * c_i:18 + 6 * c_i])
            cs.q_CG = quat_normalized(quat_mul(dq, cs.q_CG))
            cs.p_C_G = cs.p_C_G + deltaX[18 + 6 * c_i:21 + 6 * c_i]
        tempMat = np.eye(n, dtype=dt) - K @ T_H
        P_c = tempMat @ P @ tempMat.T + K @ R_n @ K.T
        P_c = ((P_c + P_c.T) / S(2)).astype(dt)
        self._split_P(P_c)

    def _build_update_quat(self, dtheta):
        """msckf.h:851-872."""
        S = self.S
        dq = S(0.5) * dtheta
        cs = np.dot(dq, dq)
        w = S(1) if cs > 1 else np.sqrt(S(1) - cs)
        return quat_normalized(np.array([-dq[0], -dq[1], -dq[2], w], dtype=self.dt))

    # ------------------------------------------------------------------ msckf.h:1049-1098
    def _find_redundant_cam_states(self):
        rm = []
        cs = self.cam_states
        if len(cs) < 5:
            return rm
        dist_thresh = self.params["redundancy_distance_thresh"]
        angle_thresh = self.params["redundancy_angle_thresh"]
        kf_pos, kf_q = cs[0].p_C_G, cs[0].q_CG
        nxt = 1
        protected = len(cs) - 3
        while nxt != protected:
            d = cs[nxt].p_C_G - kf_pos
            distance = np.sqrt(np.dot(d, d))
            angle = quat_angular_distance(kf_q, cs[nxt].q_CG)
            if distance < dist_thresh and angle < angle_thresh:
                rm.append(cs[nxt].state_id)
            else:
                kf_pos, kf_q = cs[nxt].p_C_G, cs[nxt].q_CG
            nxt += 1
            if len(cs) - len(rm) <= self.params["max_cam_states"]:
                break
        num_over_max = (len(cs) - len(rm)) - self.params["max_cam_states"]
        for i in range(num_over_max):
            if cs[i].state_id not in rm:
                rm.append(cs[i].state_id)
        if len(rm) < 2:
            rm = []
        rm.sort()
        return rm

    # ------------------------------------------------------------------ msckf.h:453-682
    def pruneRedundantStates(self):
        dt, S = self.dt, self.S
        if len(self.cam_states) < 20:
            return
        rm_ids = self._find_redundant_cam_states()
        for feature in self.feature_tracks:
            involved = []
            obs_id = None
            for cam_id in rm_ids:
                if cam_id in feature.cam_state_indices:
                    involved.append(cam_id)
                    obs_id = feature.cam_state_indices.index(cam_id)
            if len(involved) == 0:
                continue
            if len(involved) == 1:
                del feature.observations[obs_id]
                del feature.cam_state_indices[obs_id]
                continue
            if not feature.initialized:
                assoc = [c for c in self.cam_states if c.state_id in feature.cam_state_indices]

                def drop():
                    for cam_id in involved:
                        if cam_id in feature.cam_state_indices:
                            k = feature.cam_state_indices.index(cam_id)
                            del feature.cam_state_indices[k]
                            del feature.observations[k]

                if not self._check_motion(feature.observations[0], assoc):
                    drop()
                    continue
                ok, p_f_G = self._initialize_position(assoc, feature.observations)
                if not ok:
                    drop()
                    continue
                feature.initialized = True
                feature.p_f_G = p_f_G
                self.map.append(p_f_G)
        n = 15 + 6 * len(self.cam_states)
        u_var, v_var = self.noise["u_var_prime"], self.noise["v_var_prime"]
        H_blocks, r_blocks, Ro_blocks = [], [], []
        for feature in self.feature_tracks:
            involved, involved_obs = [], []
            for cam_id in rm_ids:
                if cam_id in feature.cam_state_indices:
                    involved.append(cam_id)
                    involved_obs.append(feature.observations[feature.cam_state_indices.index(cam_id)])
            nObs = len(involved)
            if nObs == 0:
                continue
            involved_states, cam_idx = [], []
            for pos, c in enumerate(self.cam_states):
                if c.state_id in involved:
                    involved_states.append(c)
                    cam_idx.append(pos)
            r_j = self._calc_residual(feature.p_f_G, involved_states, involved_obs)
            R_diag = np.tile(np.array([u_var, v_var], dtype=dt), nObs)
            H_x_j, A_j = self._calc_meas_jacobian(feature.p_f_G, cam_idx)
            r_x_j = (A_j.T @ r_j).astype(dt)
            R_x_j = (A_j.T @ (R_diag[:, None] * A_j)).astype(dt)
            ok, _ = self._gating_test(H_x_j, r_x_j, nObs - 1)
            if ok:
                H_blocks.append(H_x_j)
                r_blocks.append(r_x_j)
                Ro_blocks.append(R_x_j)
            for cam_id in involved:
                if cam_id in feature.cam_state_indices:
                    k = feature.cam_state_indices.index(cam_id)
                    del feature.cam_state_indices[k]
                    del feature.observations[k]
        if H_blocks:
            self._measurement_update(np.vstack(H_blocks), np.concatenate(r_blocks), Ro_blocks)
        num_states = len(self.cam_states)
        keep = []
        new_states = []
        for pos, c in enumerate(self.cam_states):
            if c.state_id in rm_ids:
                self.pruned_states.append(c)
            else:
                keep.append(pos)
                new_states.append(c)
        if len(keep) != num_states:
            self.cam_states = new_states
            self._gather_cov(keep)

    def _gather_cov(self, keep_positions):
        """matrix_utils.h:58-87 square_slice / column_slice on 6-blocks."""
        idx = np.concatenate([np.arange(6 * k, 6 * k + 6) for k in keep_positions]) if keep_positions else np.zeros(0, int)
        self.cam_covar = self.cam_covar[np.ix_(idx, idx)].copy()
        self.imu_cam_covar = self.imu_cam_covar[:, idx].copy()

    # ------------------------------------------------------------------ msckf.h:685-761
    def pruneEmptyStates(self):
        max_states = self.params["max_cam_states"]
        if len(self.cam_states) < max_states:
            return
        num = len(self.cam_states)
        last_to_remove = num - max_states - 1
        if self.cam_states[0].tracked_feature_ids:
            return
        for i in range(1, num - max_states):
            if self.cam_states[i].tracked_feature_ids:
                last_to_remove = i - 1
                break
        ndel = last_to_remove + 1
        if ndel <= 0:
            return
        for i in range(ndel):
            self.pruned_states.append(self.cam_states[i])
        self.cam_states = self.cam_states[ndel:]
        self._gather_cov(list(range(ndel, num)))

    # ------------------------------------------------------------------ msckf.h:765-807
    def finish(self):
        for i in range(len(self.tracked_feature_ids)):
            fid = self.tracked_feature_ids[i]
            states, idx = self._remove_tracked_feature(fid)
            if len(states) >= self.params["min_track_length"]:
                tr = TrackToResidualize()
                src = self.feature_tracks[i]
                if src.feature_id != fid:
                    for ft in self.feature_tracks:
                        if ft.feature_id == fid:
                            src = ft
                            break
                tr.feature_id = src.feature_id
                tr.observations = [o.copy() for o in src.observations]
                tr.initialized = src.initialized
                if src.initialized:
                    tr.p_f_G = src.p_f_G.copy()
                tr.cam_states = states
                tr.cam_state_indices = idx
                self.feature_tracks_to_residualize.append(tr)
            self.tracks_to_remove.append(fid)
        self.marginalize()

    # ------------------------------------------------------------------ getters :810-848
    def getNumCamStates(self):
        return len(self.cam_states)

    def getImuState(self):
        return {k: v.copy() for k, v in self.imu.items()}

    def getCovariance(self):
        return self._assemble_P()

    def getPrunedStates(self):
        self.pruned_states.sort(key=lambda c: c.state_id)
        return self.pruned_states
