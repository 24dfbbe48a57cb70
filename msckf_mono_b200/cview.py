"""ctypes binding of the "C view of msckf_mono::MSCKF<_S>" (include/msckf_mono_c.h).

The same function set is exported by the B200 engine (libmsckf_b200.so, prefix
``msckf_mono_``) and -- for tests and the CPU baseline only -- by the oracle
(oracle/libmsckf_oracle.so, prefix ``msckf_oracle_``); this class is prefix-agnostic so the
parity tests drive both through identical calls.  All scalars cross the boundary as double.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_ip = C.POINTER(C.c_int)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def pack_camera(cam):
    return np.concatenate([[cam["c_u"], cam["c_v"], cam["f_u"], cam["f_v"], cam["b"]],
                           np.asarray(cam["q_CI"], float), np.asarray(cam["p_C_I"], float)])


def pack_noise(nz):
    return np.concatenate([[nz["u_var_prime"], nz["v_var_prime"]], np.asarray(nz["Q_imu"], float).ravel(),
                           np.asarray(nz["initial_imu_covar"], float).ravel()])


def pack_params(p):
    return np.array([p["max_gn_cost_norm"], p["min_rcond"], p["translation_threshold"],
                     p["redundancy_angle_thresh"], p["redundancy_distance_thresh"],
                     p["min_track_length"], p["max_track_length"], p["max_cam_states"]], dtype=float)


def pack_imu_state(s):
    return np.concatenate([np.asarray(s[k], float) for k in ("p_I_G", "v_I_G", "b_g", "b_a", "g", "q_IG")])


def marginalize_batch(filters, threads=8):
    """marginalize() on independent filters through the C view's batched entry point: launch on all, collect on all, the
    per-filter host work spread over `threads` host threads (engine filters only)."""
    if not filters:
        return
    f0 = filters[0]
    arr = (C.c_void_p * len(filters))(*[f.h for f in filters])
    f0._chk(f0._f("marginalize_batch")(arr, C.c_int(len(filters)), C.c_int(int(threads))), "marginalize_batch")


class FilterBatch:
    """msckf_mono::MSCKFBatch<_S> through the C view: a persistent device batch over engine filters of one dtype."""

    def __init__(self, filters, threads=1):
        self.filters = list(filters)
        f0 = self.filters[0]
        self.lib = f0.lib
        arr = (C.c_void_p * len(self.filters))(*[f.h for f in self.filters])
        self.h = C.c_void_p()
        f0._chk(self.lib.msckf_mono_batch_create(arr, C.c_int(len(self.filters)), C.c_int(int(threads)), C.byref(self.h)), "batch_create")

    def marginalize(self):
        self.filters[0]._chk(self.lib.msckf_mono_batch_marginalize(self.h), "batch_marginalize")

    def handle(self):
        fn = self.lib.msckf_mono_batch_handle
        fn.restype = C.c_void_p
        return fn(self.h)

    def close(self):
        if self.h:
            self.lib.msckf_mono_batch_destroy.argtypes = [C.c_void_p]
            self.lib.msckf_mono_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CFilter:
    """MSCKF<_S> surface over a C view.  dtype: np.float32 or np.float64."""

    def __init__(self, lib_path, prefix, dtype=np.float64, create_args=()):
        self.lib = C.CDLL(str(lib_path))
        self.prefix = prefix
        self.dtype = np.dtype(dtype)
        self.h = C.c_void_p()
        code = 0 if self.dtype == np.float32 else 1
        rc = self._f("create")(C.c_int(code), *create_args, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"{prefix}create failed rc={rc}")

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _chk(self, rc, what):
        if rc < 0:
            msg = ""
            try:
                fn = getattr(self.lib, self.prefix + "last_error")
                fn.restype = C.c_char_p
                msg = (fn() or b"").decode()
            except AttributeError:
                pass
            raise RuntimeError(f"{self.prefix}{what} failed rc={rc} {msg}")
        return rc

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _round(self, a):
        """inputs are rounded to the filter dtype first, so double(float(x)) is exact."""
        return np.asarray(a, dtype=self.dtype).astype(np.float64)

    # ---- MSCKF<_S> surface -------------------------------------------------------------
    def initialize(self, camera, noise, params, imu_state):
        a, pa = _d(self._round(pack_camera(camera)))
        b, pb = _d(self._round(pack_noise(noise)))
        c, pc = _d(pack_params({k: (self._round(v) if not isinstance(v, int) else v) for k, v in params.items()}))
        d, pd = _d(self._round(pack_imu_state(imu_state)))
        self._chk(self._f("initialize")(self.h, pa, pb, pc, pd), "initialize")

    def propagate(self, omega, a, dT):
        m, pm = _d(self._round(np.concatenate([omega, a, [dT]])))
        self._chk(self._f("propagate")(self.h, pm), "propagate")

    def augmentState(self, state_id, time):
        self._chk(self._f("augment_state")(self.h, C.c_int(int(state_id)), C.c_double(float(self._round(time)))), "augment_state")

    def _obs(self, z, ids):
        z = self._round(np.asarray(z, float).reshape(-1, 2))
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        assert len(ids) == len(z)
        return z, z.ctypes.data_as(_dp), ids, ids.ctypes.data_as(_u64p), C.c_int(len(ids))

    def update(self, z, ids):
        z, pz, ids, pi, n = self._obs(z, ids)
        self._chk(self._f("update")(self.h, pz, pi, n), "update")

    def addFeatures(self, z, ids):
        z, pz, ids, pi, n = self._obs(z, ids)
        self._chk(self._f("add_features")(self.h, pz, pi, n), "add_features")

    def marginalize(self):
        self._chk(self._f("marginalize")(self.h), "marginalize")

    def pruneRedundantStates(self):
        self._chk(self._f("prune_redundant_states")(self.h), "prune_redundant_states")

    def pruneEmptyStates(self):
        self._chk(self._f("prune_empty_states")(self.h), "prune_empty_states")

    def finish(self):
        self._chk(self._f("finish")(self.h), "finish")

    # ---- getters -----------------------------------------------------------------------
    def getNumCamStates(self):
        return self._chk(self._f("get_num_cam_states")(self.h), "get_num_cam_states")

    def getImuState(self):
        o = np.zeros(29)
        self._chk(self._f("get_imu_state")(self.h, o.ctypes.data_as(_dp)), "get_imu_state")
        return {"p_I_G": o[0:3], "v_I_G": o[3:6], "b_g": o[6:9], "b_a": o[9:12], "g": o[12:15], "q_IG": o[15:19],
                "p_I_G_null": o[19:22], "v_I_G_null": o[22:25], "q_IG_null": o[25:29]}

    def getCamStates(self):
        M = self.getNumCamStates()
        poses = np.zeros((max(M, 1), 7))
        ids = np.zeros((max(M, 1), 2), dtype=np.int32)
        times = np.zeros(max(M, 1))
        self._chk(self._f("get_cam_states")(self.h, poses.ctypes.data_as(_dp), ids.ctypes.data_as(_ip),
                                          times.ctypes.data_as(_dp)), "get_cam_states")
        return {"p_C_G": poses[:M, 0:3], "q_CG": poses[:M, 3:7], "state_id": ids[:M, 0],
                "last_correlated_id": ids[:M, 1], "time": times[:M]}

    def getCamTrackedIds(self, cam):
        cap = 4096
        out = np.zeros(cap, dtype=np.uint64)
        n = self._chk(self._f("get_cam_tracked_ids")(self.h, C.c_int(cam), out.ctypes.data_as(_u64p), C.c_int(cap)), "get_cam_tracked_ids")
        return out[:n].copy()

    def getCovariance(self):
        n = 15 + 6 * self.getNumCamStates()
        out = np.zeros((n, n))
        n2 = self._chk(self._f("get_covariance")(self.h, out.ctypes.data_as(_dp)), "get_covariance")
        assert n2 == n, (n2, n)
        return out

    def getMap(self):
        cap = 8192
        out = np.zeros((cap, 3))
        n = self._chk(self._f("get_map")(self.h, out.ctypes.data_as(_dp), C.c_int(cap)), "get_map")
        return out[:n].copy()

    def getPrunedStates(self):
        cap = 65536
        poses = np.zeros((cap, 7))
        ids = np.zeros((cap, 2), dtype=np.int32)
        n = self._chk(self._f("get_pruned_states")(self.h, poses.ctypes.data_as(_dp), ids.ctypes.data_as(_ip), C.c_int(cap)), "get_pruned_states")
        return {"p_C_G": poses[:n, :3].copy(), "q_CG": poses[:n, 3:].copy(), "state_id": ids[:n, 0].copy()}

    def getTrackedFeatureIds(self):
        cap = 65536
        out = np.zeros(cap, dtype=np.uint64)
        n = self._chk(self._f("get_tracked_feature_ids")(self.h, out.ctypes.data_as(_u64p), C.c_int(cap)), "get_tracked_feature_ids")
        return out[:n].copy()

    # ---- test hooks --------------------------------------------------------------------
    def lastReport(self):
        cap = 8192
        flags = np.zeros((cap, 4), dtype=np.int32)
        gamma = np.zeros(cap)
        pfg = np.zeros((cap, 3))
        n = self._chk(self._f("last_report")(self.h, flags.ctypes.data_as(_ip), gamma.ctypes.data_as(_dp),
                                           pfg.ctypes.data_as(_dp), C.c_int(cap)), "last_report")
        return {"cm_passed": flags[:n, 0].copy(), "valid": flags[:n, 1].copy(), "accepted": flags[:n, 2].copy(),
                "rows": flags[:n, 3].copy(), "gamma": gamma[:n].copy(), "p_f_G": pfg[:n].copy()}

    def counters(self):
        out = (C.c_long * 8)()
        self._chk(self._f("get_counters")(self.h, out), "get_counters")
        return {"num_residualized": out[0], "pfg_shifted": out[1], "pfg_oob": out[2], "n_updates": out[3],
                "rows_kept": out[4], "m": out[5]}

    def setOption(self, key, value):
        self._chk(self._f("set_option")(self.h, C.c_int(key), C.c_double(value)), "set_option")

    def lastDeltaX(self):
        cap = 15 + 6 * 512
        out = np.zeros(cap)
        n = self._chk(self._f("last_delta_x")(self.h, out.ctypes.data_as(_dp), C.c_int(cap)), "last_delta_x")
        return out[:n].copy()

    def packQueued(self):
        """the batch queued by update()/finish() in the engine's flat SoA form (engine view only)."""
        cap_t, cap_o = 8192, 8192 * 64
        off = np.zeros(cap_t + 1, dtype=np.int32)
        obs = np.zeros(2 * cap_o)
        idx = np.zeros(cap_o, dtype=np.int32)
        n = self._chk(self._f("pack_queued")(self.h, off.ctypes.data_as(_ip), obs.ctypes.data_as(_dp), idx.ctypes.data_as(_ip),
                                            C.c_int(cap_t), C.c_int(cap_o)), "pack_queued")
        tot = int(off[n])
        return off[:n + 1].copy(), obs[:2 * tot].copy(), idx[:tot].copy()

    def engineHandle(self):
        fn = self._f("engine")
        fn.restype = C.c_void_p
        return fn(self.h)

    def marginalizeLaunch(self):
        self._chk(self._f("marginalize_launch")(self.h), "marginalize_launch")

    def marginalizeCollect(self):
        self._chk(self._f("marginalize_collect")(self.h), "marginalize_collect")

    def queuedTracks(self):
        cap = 8192
        ids = np.zeros(cap, dtype=np.uint64)
        nobs = np.zeros(cap, dtype=np.int32)
        n = self._chk(self._f("queued_tracks")(self.h, ids.ctypes.data_as(_u64p), nobs.ctypes.data_as(_ip), C.c_int(cap)), "queued_tracks")
        return ids[:n].copy(), nobs[:n].copy()
