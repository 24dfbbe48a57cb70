/* include/msckf_b200.h -- C-ABI of the B200 MSCKF measurement-update engine (libmsckf_b200.so).
 *
 * The reference (daniilidis-group/msckf_mono) has no plugin / FFI boundary: its filter is the header-only
 * C++ class template msckf_mono::MSCKF<_S> (include/msckf_mono/msckf.h:31-1512).  This C-ABI is what the
 * replacement class of the same name (include/msckf_mono/msckf.h in THIS repo) binds: each entry point below
 * names the reference member function(s) whose numerics it replaces.  All integer bookkeeping (track lists,
 * clone ids, pruning decisions) stays on the host in the class shim, exactly as in the reference.
 *
 * Conventions: plain pointers and sizes only.  `dtype` selects the scalar type of every `void*` array
 * (MSCKF_B200_F32: float, MSCKF_B200_F64: double) -- the reference's template parameter _S.  Quaternions are
 * (x,y,z,w) (Eigen coeffs() order).  Error-state order: IMU [dtheta, db_g, dv, db_a, dp] then per clone
 * [dtheta_C, dp_C] (msckf.h:885-889, :1376-1390).  Every function returns 0 on success, a negative
 * msckf_b200_status otherwise; there is NO CPU fallback -- a failed device call is a hard error.
 * A handle owns its device buffers and one CUDA stream; calls on one handle must be serialised by the
 * caller (like the reference class), different handles are independent.
 */
#ifndef MSCKF_B200_H_
#define MSCKF_B200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum msckf_b200_dtype { MSCKF_B200_F32 = 0, MSCKF_B200_F64 = 1 };
enum msckf_b200_status {
  MSCKF_B200_OK = 0,
  MSCKF_B200_ERR_CUDA = -1,      /* a CUDA runtime call or kernel failed */
  MSCKF_B200_ERR_CAPACITY = -2,  /* more clones / tracks / observations than msckf_b200_create reserved */
  MSCKF_B200_ERR_ARG = -3,       /* invalid argument (e.g. track longer than 98 observations, msckf.h:91) */
  MSCKF_B200_ERR_NO_DEVICE = -4, /* no CUDA device: the engine never computes on the CPU */
  MSCKF_B200_ERR_STATE = -5,     /* call order violated */
  MSCKF_B200_ERR_NUMERIC = -6    /* the update produced a non-finite delta-x or covariance entry.  The update HAS been applied, like in
                                    the reference (whose only anomaly handling is a print, msckf.h:1405-1409): returned by _fetch
                                    after the report was filled, so the caller decides (the class shim prints and carries on) */
};
enum msckf_b200_mode {
  MSCKF_B200_MARGINALIZE = 0, /* MSCKF::marginalize() msckf.h:336-449: loop A + loop B + measurementUpdate */
  MSCKF_B200_TRIANGULATE = 1, /* checkMotion :980-1025 + initializePosition :1147-1285 only (pruneRedundantStates :488-531) */
  MSCKF_B200_RESIDUALIZE = 2  /* loop B + measurementUpdate at caller-given p_f_G (pruneRedundantStates :545-614) */
};

typedef struct msckf_b200_engine msckf_b200_engine;

typedef struct {
  int dtype;       /* msckf_b200_dtype */
  int device;      /* CUDA device ordinal */
  int max_clones;  /* initial capacity M_max (covariance is (15+6*M_max)^2) */
  int max_tracks;  /* initial capacity of one track batch */
  int max_obs;     /* initial capacity of one batch's total observation count */
} msckf_b200_config;
/* The three capacities are INITIAL sizes: like the reference, whose window and track lists are unbounded std::vectors, the
 * engine grows its device buffers on demand (x2, contents preserved).  Hard limits that remain: 98 observations per track
 * (the reference's chi-square table has 99 entries, msckf.h:91 -- longer tracks index past it there) and the window the
 * tail kernel's shared-memory panels can hold (about 135 clones): both fail with MSCKF_B200_ERR_CAPACITY / _ARG. */

/* One batch of feature tracks, flat SoA (types.h:102-113 featureTrackToResidualize without the clone copies:
 * the clone poses are device resident and addressed by POSITION in the sliding window, msckf.h:1481). */
typedef struct {
  int n_tracks;
  const int* obs_offset;  /* [n_tracks+1] prefix offsets into obs / clone_index */
  const void* obs;        /* [2*obs_offset[n_tracks]] normalised image coordinates (u,v), dtype scalars */
  const int* clone_index; /* [obs_offset[n_tracks]] positional index of the observing clone */
  const void* p_f_G;      /* MSCKF_B200_RESIDUALIZE only: [3*n_tracks] feature positions; else NULL */
} msckf_b200_tracks;

/* Per-track results; every pointer may be NULL (not wanted).  Caller-owned host buffers of n_tracks entries. */
typedef struct {
  int* cm_ok;    /* checkMotion result */
  int* tri_ok;   /* initializePosition validity */
  int* valid;    /* valid_tracks[] of msckf.h:348 */
  int* accepted; /* passed gatingTest */
  void* gamma;   /* Mahalanobis distance, dtype scalars */
  void* p_f_G;   /* [3*n_tracks] triangulated position, dtype scalars */
  int m;         /* out: stacked rows of accepted tracks (stack_counter, msckf.h:443) */
  int rank;      /* out: independent rows kept by the compression */
} msckf_b200_report;
/* (a non-finite result is signalled by _fetch returning MSCKF_B200_ERR_NUMERIC, with the report filled) */

int msckf_b200_create(const msckf_b200_config* cfg, msckf_b200_engine** out);
int msckf_b200_destroy(msckf_b200_engine* e);

/* MSCKF::initialize msckf.h:72-97.
 * camera: q_CI[4], p_C_I[3]; noise: u_var_prime, v_var_prime, Q_imu[144], initial_imu_covar[225] (row-major);
 * params: max_gn_cost_norm, translation_threshold; imu_state: p_I_G[3], v_I_G[3], b_g[3], b_a[3], g[3], q_IG[4]. */
int msckf_b200_initialize(msckf_b200_engine* e, const void* camera, const void* noise, const void* params, const void* imu_state);
/* MSCKF::propagate msckf.h:101-145.  reading: omega[3], a[3], dT */
int msckf_b200_propagate(msckf_b200_engine* e, const void* reading);
/* k consecutive propagate() calls (readings: [k][7] scalars) in one kernel launch per 16 readings; same arithmetic, reading
 * by reading.  The class shim queues the IMU readings that arrive between two images (src/ros_interface.cpp:92-97) */
int msckf_b200_propagate_n(msckf_b200_engine* e, const void* readings, int k);
/* MSCKF::augmentState msckf.h:148-212 (numeric part: clone pose + covariance augmentation) */
int msckf_b200_augment(msckf_b200_engine* e);
/* marginalize / pruneRedundantStates numerics, asynchronous on the handle's stream.  The input arrays are
 * copied to pinned staging before return; results are fetched (and the stream synchronised) by _fetch.
 * A track with a single observation cannot be triangulated or projected (2L - 3 < 1 rows): it is reported as rejected
 * (cm_ok = tri_ok = valid = accepted = 0), not as an error. */
int msckf_b200_update_async(msckf_b200_engine* e, int mode, const msckf_b200_tracks* tracks);
int msckf_b200_fetch(msckf_b200_engine* e, msckf_b200_report* report);
/* Zero-copy packing: host pointers INTO the engine's pinned input block for a batch of n_tracks tracks / n_obs observations
 * (out->obs_offset, ->obs, ->clone_index, ->p_f_G; out->n_tracks = n_tracks).  The caller fills them and passes the same struct
 * to _update / _update_async / _stage, which then skips its own copy.  Valid until the next call on this handle. */
int msckf_b200_input_buffer(msckf_b200_engine* e, int n_tracks, int n_obs, msckf_b200_tracks* out);
/* update_async split in two: _stage validates and copies the batch into HBM, _launch enqueues the kernels
 * (bench.py times _launch alone: "inputs already resident in HBM") */
int msckf_b200_stage(msckf_b200_engine* e, int mode, const msckf_b200_tracks* tracks);
int msckf_b200_launch(msckf_b200_engine* e);
/* _launch bracketed by CUDA events on the handle's stream (kernels only); *ms = device time */
int msckf_b200_launch_timed(msckf_b200_engine* e, float* ms);
/* with option key 1 set, per-kernel device times of the last launch (launch order); returns the count */
int msckf_b200_kernel_times(msckf_b200_engine* e, float* ms, const char** names, int cap);
/* with option key 1 set, %globaltimer stamps (ns) taken by the tail kernel at its phase boundaries; 0-terminated */
int msckf_b200_tail_profile(msckf_b200_engine* e, unsigned long long* out, int cap);
/* update_async + fetch */
int msckf_b200_update(msckf_b200_engine* e, int mode, const msckf_b200_tracks* tracks, msckf_b200_report* report);
/* ---- device-side batches of independent filters (sequences / Monte-Carlo trials sharing one GPU) ----------
 * A batch groups n engines (same dtype, same device).  One update of the whole group is ONE launch per kernel -- the filter
 * index rides in blockIdx.z, one thread-block cluster per filter runs the serial tail -- replayed as ONE CUDA graph, with ONE
 * packed host->device copy of all track batches and ONE packed device->host copy of all reports.  Results are bit-identical
 * to n separate msckf_b200_update calls.  While a batch exists its engines share the batch's stream (their other calls --
 * propagate, augment, prune, getters -- stay valid and are ordered with the batch's updates); destroying the batch gives
 * every engine its own stream back.  An engine belongs to at most one batch. */
typedef struct msckf_b200_batch msckf_b200_batch;
int msckf_b200_batch_create(msckf_b200_engine** engines, int n, msckf_b200_batch** out);
int msckf_b200_batch_destroy(msckf_b200_batch* b);
/* tracks[n] / reports[n]: one per engine, in the order given to _create; reports may be NULL.  Up to `threads` host threads
 * share the validation and packing of the n track batches (<= 1: the calling thread does all of it); the engine uses fewer
 * for small batches -- one per 8 MB of observations: packing from several cores makes the following host-to-device copy slower
 * than the packing gets faster (measured, profiles/README.md). */
int msckf_b200_batch_update_async(msckf_b200_batch* b, int mode, const msckf_b200_tracks* tracks, int threads);
int msckf_b200_batch_fetch(msckf_b200_batch* b, msckf_b200_report* reports);
int msckf_b200_batch_update(msckf_b200_batch* b, int mode, const msckf_b200_tracks* tracks, msckf_b200_report* reports, int threads);
/* the same split as _stage / _launch / _launch_timed of a single engine (bench.py: inputs resident in HBM) */
int msckf_b200_batch_stage(msckf_b200_batch* b, int mode, const msckf_b200_tracks* tracks, int threads);
int msckf_b200_batch_launch(msckf_b200_batch* b);
int msckf_b200_batch_launch_timed(msckf_b200_batch* b, float* ms);
/* per-kernel device times of the last batch launch (option key 1 of the FIRST engine set); returns the count */
int msckf_b200_batch_kernel_times(msckf_b200_batch* b, float* ms, const char** names, int cap);
long long msckf_b200_batch_launch_count(const msckf_b200_batch* b);
void* msckf_b200_batch_stream(msckf_b200_batch* b);
/* Convenience: one update on each of n engines through a batch.  If the engines already form a batch (in this order) it is
 * used; otherwise a temporary one is created and destroyed around the call.  Returns the first non-zero status. */
int msckf_b200_update_batch(msckf_b200_engine** engines, int n, int mode, const msckf_b200_tracks* tracks, msckf_b200_report* reports,
                            int threads);
/* covariance / pose gather of pruneEmptyStates msckf.h:685-761 and pruneRedundantStates :616-681:
 * keep[] = ascending positional indices of the clones that survive (n_keep = 0: only the IMU block stays).
 * Asynchronous: keep[] travels as a kernel argument, nothing is synchronised and captured graphs stay valid. */
int msckf_b200_prune(msckf_b200_engine* e, const int* keep, int n_keep);

int msckf_b200_num_clones(msckf_b200_engine* e);
/* imu: p,v,b_g,b_a,g,q_IG, p_null,v_null,q_null (29 scalars); clone_poses: [M*7] p(3) q(4).  Either may be NULL. */
int msckf_b200_get_state(msckf_b200_engine* e, void* imu, void* clone_poses);
/* full (15+6M)^2 covariance, row-major, dtype scalars */
int msckf_b200_get_covariance(msckf_b200_engine* e, void* out);
/* overwrite the covariance (same layout; checkpoint restore, Monte-Carlo initialisation) */
int msckf_b200_set_covariance(msckf_b200_engine* e, const void* in);
/* counters[0..7]: num_feature_tracks_residualized_, pfg_shifted, pfg_oob, n_updates, last m, last rank, 0, 0 */
int msckf_b200_get_counters(msckf_b200_engine* e, long long* counters);
/* last delta-x (fp64), returns its length */
int msckf_b200_last_delta_x(msckf_b200_engine* e, double* out, int cap);
/* diagnostics of the last update's rank decision: for every index of the compressed system, the pivot of the basis Gram matrix
 * relative to its original diagonal (the squared sine against the span of the previous basis vectors) at the moment it was
 * compared with the rank threshold (option 0).  Directions in the null space of H_o show up as rounding noise here; the gap
 * between that noise and the smallest kept pivot is the margin of the decision.  Returns the dimension 15 + 6 M of that update (0 before the first one). */
int msckf_b200_rank_pivots(msckf_b200_engine* e, double* out, int cap);
/* option keys: 0 = rank threshold of the compression (relative pivot of the basis Gram matrix = squared sine to the span of the previous basis vectors, default 1e-9:
 *                  measured pivots of null directions are rounding noise up to ~2e-11, the smallest real pivot of the test windows is > 1e-3,
 *                  see msckf_b200_rank_pivots);
 *              1 = record per-kernel CUDA events in _launch (profiling aid, default off);
 *              2 = replay the update's kernel sequence as a CUDA graph when the batch signature repeats (default on);
 *              3 = fuse the forward substitution into the blocked Cholesky of the tail kernel where the window allows it
 *                  (15 + 6M <= 255; default on; 0 = always run it as a separate sweep -- same results up to rounding);
 *              4 = launch the update's kernels with programmatic dependent launch (default on);
 *              5 = Gram products of the compression on the FP64 tensor-core path (mma.sync m8n8k4 f64, SASS DMMA; default on;
 *                  0 = SIMT DFMA tiles -- same results to fp64 rounding);
 *              6 = threads that work on one track in the feature kernel: 128 (a CTA per track: lowest latency), 64, or 32
 *                  (a warp per track, two tracks per CTA: highest throughput); 0 = by batch size (default: 128 for one
 *                  filter, 32 for a device batch).  The results are bit-identical for every value;
 *              7 = CTAs that run the tail's chains of diagonal blocks: 2 (default where the device accepts clusters of 9: Gamma's
 *                  and S''s chains side by side on two CTAs, seven worker CTAs) or 1 (cluster of 8, both chains on one CTA);
 *                  0 = default.  Same rank decisions; results agree to fp64 rounding. */
int msckf_b200_set_option(msckf_b200_engine* e, int key, double value);
/* checkpoint / resume: copy the complete filter state of src into dst (same dtype and capacities) */
int msckf_b200_copy_state(msckf_b200_engine* dst, const msckf_b200_engine* src);
/* number of kernels launched by this handle so far (bench.py's gpu_launches) */
long long msckf_b200_launch_count(const msckf_b200_engine* e);
/* the handle's cudaStream_t (for event timing on the launching stream) */
void* msckf_b200_stream(msckf_b200_engine* e);
int msckf_b200_synchronize(msckf_b200_engine* e);
const char* msckf_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MSCKF_B200_H_ */
