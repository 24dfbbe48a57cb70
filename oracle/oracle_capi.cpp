// oracle/oracle_capi.cpp -- C view of the CPU oracle (msckf_oracle::MSCKF<float|double>) for ctypes.
// TEST INFRASTRUCTURE ONLY.  All scalars cross this boundary as double (exact widening of float).
#include <cstring>
#include <memory>
#include "msckf_oracle.hpp"
using namespace msckf_oracle;

namespace {
struct Base {
  virtual ~Base() {}
  virtual void initialize(const double* cam, const double* noise, const double* params, const double* imu) = 0;
  virtual void propagate(const double* m) = 0;
  virtual void augment(int id, double t) = 0;
  virtual void update(const double* z, const uint64_t* ids, int n) = 0;
  virtual void add(const double* z, const uint64_t* ids, int n) = 0;
  virtual void marginalize() = 0;
  virtual void prune_redundant() = 0;
  virtual void prune_empty() = 0;
  virtual void finish() = 0;
  virtual int num_cam() = 0;
  virtual void imu_state(double* out) = 0;
  virtual void cam_states(double* poses, int* ids, double* times) = 0;
  virtual int cam_tracked(int cam, uint64_t* out, int cap) = 0;
  virtual int covariance(double* out) = 0;
  virtual int map(double* out, int cap) = 0;
  virtual int pruned(double* poses, int* ids, int cap) = 0;
  virtual int tracked_ids(uint64_t* out, int cap) = 0;
  virtual int report(int* flags, double* gamma, double* pfg, int cap) = 0;
  virtual void counters(long* out) = 0;
  virtual void set_option(int key, double v) = 0;
  virtual int last_dx(double* out, int cap) = 0;
  virtual int track_info(int i, uint64_t* id, int* nobs, int* initialized) = 0;
  virtual int queued_tracks(uint64_t* ids, int* nobs, int cap) = 0;
};

template <class S>
struct Impl : Base {
  MSCKF<S> f;
  static V3<S> v3(const double* p) { return {{(S)p[0], (S)p[1], (S)p[2]}}; }
  static Quat<S> q4(const double* p) { return {(S)p[0], (S)p[1], (S)p[2], (S)p[3]}; }
  void initialize(const double* c, const double* nz, const double* pr, const double* im) override {
    Camera<S> cam;
    cam.c_u = (S)c[0]; cam.c_v = (S)c[1]; cam.f_u = (S)c[2]; cam.f_v = (S)c[3]; cam.b = (S)c[4];
    cam.q_CI = q4(c + 5); cam.p_C_I = v3(c + 9);
    NoiseParams<S> noise;
    noise.u_var_prime = (S)nz[0]; noise.v_var_prime = (S)nz[1];
    for (int i = 0; i < 144; ++i) noise.Q_imu.a[i] = (S)nz[2 + i];
    for (int i = 0; i < 225; ++i) noise.initial_imu_covar.a[i] = (S)nz[146 + i];
    MSCKFParams<S> p;
    p.max_gn_cost_norm = (S)pr[0]; p.min_rcond = (S)pr[1]; p.translation_threshold = (S)pr[2];
    p.redundancy_angle_thresh = (S)pr[3]; p.redundancy_distance_thresh = (S)pr[4];
    p.min_track_length = (int)pr[5]; p.max_track_length = (int)pr[6]; p.max_cam_states = (int)pr[7];
    ImuState<S> s;
    s.p_I_G = v3(im); s.v_I_G = v3(im + 3); s.b_g = v3(im + 6); s.b_a = v3(im + 9); s.g = v3(im + 12);
    s.q_IG = q4(im + 15);
    f.initialize(cam, noise, p, s);
  }
  void propagate(const double* m) override {
    ImuReading<S> r; r.omega = v3(m); r.a = v3(m + 3); r.dT = (S)m[6];
    f.propagate(r);
  }
  void augment(int id, double t) override { f.augmentState(id, (S)t); }
  static void conv(const double* z, const uint64_t* ids, int n, std::vector<Obs<S>>& o, std::vector<size_t>& i) {
    o.resize(n); i.resize(n);
    for (int k = 0; k < n; ++k) { o[k].u = (S)z[2 * k]; o[k].v = (S)z[2 * k + 1]; i[k] = (size_t)ids[k]; }
  }
  void update(const double* z, const uint64_t* ids, int n) override {
    std::vector<Obs<S>> o; std::vector<size_t> i; conv(z, ids, n, o, i); f.update(o, i);
  }
  void add(const double* z, const uint64_t* ids, int n) override {
    std::vector<Obs<S>> o; std::vector<size_t> i; conv(z, ids, n, o, i); f.addFeatures(o, i);
  }
  void marginalize() override { f.marginalize(); }
  void prune_redundant() override { f.pruneRedundantStates(); }
  void prune_empty() override { f.pruneEmptyStates(); }
  void finish() override { f.finish(); }
  int num_cam() override { return (int)f.getNumCamStates(); }
  void imu_state(double* o) override {
    const ImuState<S> s = f.getImuState();
    auto put3 = [&](int off, const V3<S>& v) { for (int i = 0; i < 3; ++i) o[off + i] = v[i]; };
    auto put4 = [&](int off, const Quat<S>& q) { o[off] = q.x; o[off + 1] = q.y; o[off + 2] = q.z; o[off + 3] = q.w; };
    put3(0, s.p_I_G); put3(3, s.v_I_G); put3(6, s.b_g); put3(9, s.b_a); put3(12, s.g); put4(15, s.q_IG);
    put3(19, s.p_I_G_null); put3(22, s.v_I_G_null); put4(25, s.q_IG_null);
  }
  static void put_pose(double* p, const CamState<S>& c) {
    for (int i = 0; i < 3; ++i) p[i] = c.p_C_G[i];
    p[3] = c.q_CG.x; p[4] = c.q_CG.y; p[5] = c.q_CG.z; p[6] = c.q_CG.w;
  }
  void cam_states(double* poses, int* ids, double* times) override {
    const auto cs = f.getCamStates();
    for (size_t k = 0; k < cs.size(); ++k) {
      put_pose(poses + 7 * k, cs[k]);
      ids[2 * k] = cs[k].state_id; ids[2 * k + 1] = cs[k].last_correlated_id;
      times[k] = cs[k].time;
    }
  }
  int cam_tracked(int cam, uint64_t* out, int cap) override {
    const auto c = f.getCamState(cam);
    const int n = (int)c.tracked_feature_ids.size();
    for (int i = 0; i < std::min(n, cap); ++i) out[i] = c.tracked_feature_ids[i];
    return n;
  }
  int covariance(double* out) override {
    const Mat<S> P = f.getCovariance();
    for (size_t i = 0; i < P.a.size(); ++i) out[i] = P.a[i];
    return P.r;
  }
  int map(double* out, int cap) override {
    const auto m = f.getMap();
    for (int i = 0; i < std::min((int)m.size(), cap); ++i) for (int k = 0; k < 3; ++k) out[3 * i + k] = m[i][k];
    return (int)m.size();
  }
  int pruned(double* poses, int* ids, int cap) override {
    const auto ps = f.getPrunedStates();
    for (int k = 0; k < std::min((int)ps.size(), cap); ++k) {
      put_pose(poses + 7 * k, ps[k]); ids[2 * k] = ps[k].state_id; ids[2 * k + 1] = ps[k].last_correlated_id;
    }
    return (int)ps.size();
  }
  int tracked_ids(uint64_t* out, int cap) override {
    const auto& t = f.trackedFeatureIds();
    for (int i = 0; i < std::min((int)t.size(), cap); ++i) out[i] = t[i];
    return (int)t.size();
  }
  int report(int* flags, double* gamma, double* pfg, int cap) override {
    const auto& r = f.last_report;
    for (int i = 0; i < std::min((int)r.size(), cap); ++i) {
      flags[4 * i] = r[i].cm_passed; flags[4 * i + 1] = r[i].valid; flags[4 * i + 2] = r[i].accepted; flags[4 * i + 3] = r[i].rows;
      gamma[i] = r[i].gamma;
      for (int k = 0; k < 3; ++k) pfg[3 * i + k] = r[i].p_f_G[k];
    }
    return (int)r.size();
  }
  void counters(long* o) override {
    o[0] = (long)f.numResidualized(); o[1] = f.pfg_shifted; o[2] = f.pfg_oob; o[3] = f.n_updates;
    o[4] = f.last_rows_kept; o[5] = f.last_m;
  }
  void set_option(int key, double v) override {
    if (key == 0) f.faithful_max_rows = (int)v;
    if (key == 1) f.drop_null_rows = v != 0;
    if (key == 2) f.null_row_tol = v;
  }
  int last_dx(double* out, int cap) override {
    for (int i = 0; i < std::min((int)f.last_deltaX.size(), cap); ++i) out[i] = f.last_deltaX[i];
    return (int)f.last_deltaX.size();
  }
  int track_info(int i, uint64_t* id, int* nobs, int* initialized) override {
    const auto& t = f.featureTracks();
    if (i < 0 || i >= (int)t.size()) return (int)t.size();
    *id = t[i].feature_id; *nobs = (int)t[i].observations.size(); *initialized = t[i].initialized;
    return (int)t.size();
  }
  int queued_tracks(uint64_t* ids, int* nobs, int cap) override {
    const auto& q = f.tracksToResidualize();
    for (int i = 0; i < std::min((int)q.size(), cap); ++i) { ids[i] = q[i].feature_id; nobs[i] = (int)q[i].observations.size(); }
    return (int)q.size();
  }
};
}  // namespace

// exceptions (the stale-queue guard) must not cross the C boundary: -1 + msckf_oracle_last_error()
static thread_local std::string g_oracle_err;
template <class F>
static int oguard(F&& f) {
  try { f(); return 0; } catch (const std::exception& e) { g_oracle_err = e.what(); return -1; }
}

extern "C" {
const char* msckf_oracle_last_error(void) { return g_oracle_err.c_str(); }
int msckf_oracle_create(int dtype, void** out) {
  Base* b = (dtype == 0) ? (Base*)new Impl<float>() : (Base*)new Impl<double>();
  *out = b;
  return 0;
}
void msckf_oracle_destroy(void* h) { delete (Base*)h; }
int msckf_oracle_initialize(void* h, const double* cam, const double* noise, const double* params, const double* imu) { ((Base*)h)->initialize(cam, noise, params, imu); return 0; }
int msckf_oracle_propagate(void* h, const double* m) { ((Base*)h)->propagate(m); return 0; }
int msckf_oracle_augment_state(void* h, int id, double t) { ((Base*)h)->augment(id, t); return 0; }
int msckf_oracle_update(void* h, const double* z, const uint64_t* ids, int n) { ((Base*)h)->update(z, ids, n); return 0; }
int msckf_oracle_add_features(void* h, const double* z, const uint64_t* ids, int n) { ((Base*)h)->add(z, ids, n); return 0; }
int msckf_oracle_marginalize(void* h) { return oguard([&] { ((Base*)h)->marginalize(); }); }
int msckf_oracle_prune_redundant_states(void* h) { ((Base*)h)->prune_redundant(); return 0; }
int msckf_oracle_prune_empty_states(void* h) { ((Base*)h)->prune_empty(); return 0; }
int msckf_oracle_finish(void* h) { return oguard([&] { ((Base*)h)->finish(); }); }
int msckf_oracle_get_num_cam_states(void* h) { return ((Base*)h)->num_cam(); }
int msckf_oracle_get_imu_state(void* h, double* out) { ((Base*)h)->imu_state(out); return 0; }
int msckf_oracle_get_cam_states(void* h, double* poses, int* ids, double* times) { ((Base*)h)->cam_states(poses, ids, times); return 0; }
int msckf_oracle_get_cam_tracked_ids(void* h, int cam, uint64_t* out, int cap) { return ((Base*)h)->cam_tracked(cam, out, cap); }
int msckf_oracle_get_covariance(void* h, double* out) { return ((Base*)h)->covariance(out); }
int msckf_oracle_get_map(void* h, double* out, int cap) { return ((Base*)h)->map(out, cap); }
int msckf_oracle_get_pruned_states(void* h, double* poses, int* ids, int cap) { return ((Base*)h)->pruned(poses, ids, cap); }
int msckf_oracle_get_tracked_feature_ids(void* h, uint64_t* out, int cap) { return ((Base*)h)->tracked_ids(out, cap); }
int msckf_oracle_last_report(void* h, int* flags, double* gamma, double* pfg, int cap) { return ((Base*)h)->report(flags, gamma, pfg, cap); }
int msckf_oracle_get_counters(void* h, long* out) { ((Base*)h)->counters(out); return 0; }
int msckf_oracle_set_option(void* h, int key, double v) { ((Base*)h)->set_option(key, v); return 0; }
int msckf_oracle_last_delta_x(void* h, double* out, int cap) { return ((Base*)h)->last_dx(out, cap); }
int msckf_oracle_track_info(void* h, int i, uint64_t* id, int* nobs, int* initialized) { return ((Base*)h)->track_info(i, id, nobs, initialized); }
int msckf_oracle_queued_tracks(void* h, uint64_t* ids, int* nobs, int cap) { return ((Base*)h)->queued_tracks(ids, nobs, cap); }
}
