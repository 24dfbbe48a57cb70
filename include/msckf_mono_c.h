/* include/msckf_mono_c.h -- C view of the drop-in class msckf_mono::MSCKF<_S> (include/msckf_mono/msckf.h).
 *
 * One function per public member of the reference class (msckf.h:72-848), so that non-C++ hosts (the
 * Python parity tests and bench.py via ctypes) drive exactly the call sequence a C++ caller would.  All
 * scalars cross this view as double (exact widening of the filter's _S = float); the class beneath is
 * instantiated for float or double.  Returns 0 / a count on success, a negative value on failure
 * (msckf_mono_last_error() holds the message).  Exported by libmsckf_b200.so.
 */
#ifndef MSCKF_MONO_C_H_
#define MSCKF_MONO_C_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int msckf_mono_create(int dtype /*0 = float, 1 = double*/, void** out);
void msckf_mono_destroy(void* h);
/* device ordinal and engine capacities (0 = derive from MSCKFParams); call before initialize */
int msckf_mono_set_engine_options(void* h, int device, int max_clones, int max_tracks, int max_obs);
/* camera[12]: c_u,c_v,f_u,f_v,b,q_CI(xyzw),p_C_I; noise[371]: u_var',v_var',Q_imu[144],initial_imu_covar[225];
 * params[8]: max_gn_cost_norm,min_rcond,translation_threshold,redundancy_angle_thresh,redundancy_distance_thresh,
 *            min_track_length,max_track_length,max_cam_states; imu_state[19]: p,v,b_g,b_a,g,q_IG(xyzw)      msckf.h:72 */
int msckf_mono_initialize(void* h, const double* camera, const double* noise, const double* params, const double* imu_state);
int msckf_mono_propagate(void* h, const double* omega_a_dT);                                  /* msckf.h:101 */
int msckf_mono_augment_state(void* h, int state_id, double time);                             /* msckf.h:148 */
int msckf_mono_update(void* h, const double* measurements, const uint64_t* ids, int n);       /* msckf.h:215 */
int msckf_mono_add_features(void* h, const double* features, const uint64_t* ids, int n);     /* msckf.h:302 */
int msckf_mono_marginalize(void* h);                                                          /* msckf.h:336 */
int msckf_mono_prune_redundant_states(void* h);                                               /* msckf.h:453 */
int msckf_mono_prune_empty_states(void* h);                                                   /* msckf.h:685 */
int msckf_mono_finish(void* h);                                                               /* msckf.h:765 */
int msckf_mono_get_num_cam_states(void* h);                                                   /* msckf.h:810 */
int msckf_mono_get_imu_state(void* h, double* out29);                                         /* msckf.h:815 */
int msckf_mono_get_cam_states(void* h, double* poses7, int* ids2, double* times);             /* msckf.h:835 */
int msckf_mono_get_cam_tracked_ids(void* h, int cam, uint64_t* out, int cap);
int msckf_mono_get_covariance(void* h, double* out);
int msckf_mono_get_map(void* h, double* out, int cap);                                        /* msckf.h:820 */
int msckf_mono_get_pruned_states(void* h, double* poses7, int* ids2, int cap);                /* msckf.h:840 */
int msckf_mono_get_tracked_feature_ids(void* h, uint64_t* out, int cap);
/* diagnostics */
int msckf_mono_last_report(void* h, int* flags4 /*cm,valid,accepted,rows*/, double* gamma, double* pfg, int cap);
int msckf_mono_get_counters(void* h, long* out8);
int msckf_mono_set_option(void* h, int key, double value);
int msckf_mono_last_delta_x(void* h, double* out, int cap);
int msckf_mono_queued_tracks(void* h, uint64_t* ids, int* nobs, int cap);
/* the track batch queued by update()/finish() in the engine's flat SoA form (obs as double); returns n_tracks */
int msckf_mono_pack_queued(void* h, int* obs_offset, double* obs, int* clone_index, int cap_tracks, int cap_obs);
/* pipelining helpers: marginalize() = launch + collect */
int msckf_mono_marginalize_launch(void* h);
int msckf_mono_marginalize_collect(void* h);
/* marginalize() on n independent filters of one scalar type as ONE device batch (msckf_mono::MSCKFBatch<_S>: one launch per
 * kernel with the filter index in blockIdx.z, one CUDA graph, one packed copy each way); `threads` host threads share the
 * packing.  Bit-identical to marginalize() on each filter.  _batch_create makes the batch persistent (the filters share its
 * stream while it lives); msckf_mono_marginalize_batch is the one-shot form (temporary batch around one call). */
int msckf_mono_batch_create(void** handles, int n, int threads, void** out);
void msckf_mono_batch_destroy(void* batch);
int msckf_mono_batch_marginalize(void* batch);
void* msckf_mono_batch_handle(void* batch); /* the msckf_b200_batch* beneath */
int msckf_mono_marginalize_batch(void** handles, int n, int threads);
/* the msckf_b200_engine* beneath (for CUDA-event timing on its stream, launch counts, state copies) */
void* msckf_mono_engine(void* h);
/* copy the complete filter (host bookkeeping + device state) of src into dst; both must be initialised alike */
int msckf_mono_clone_from(void* dst, void* src);
const char* msckf_mono_last_error(void);
#ifdef __cplusplus
}
#endif
#endif
