// msckf_mono_b200/csrc/engine.cu -- host side of the B200 engine and the C-ABI of include/msckf_b200.h.
// One handle = one filter: device-resident covariance / clone poses / IMU state, one CUDA stream,
// pinned staging for the per-call track batch and report.  No CPU fallback anywhere.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/msckf_b200.h"
#include "common.cuh"
#include "feature_kernels.cuh"
#include "gram_kernels.cuh"
#include "state_kernels.cuh"
#include "tail_kernels.cuh"
#include "tail_cluster.cuh"
#include "tail_fused.cuh"

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(MSCKF_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

constexpr int kMaxSplit = 16;
constexpr int kTailCluster = 8;  // portable cluster size: the serial EKF tail runs on 8 SMs of one GPC
constexpr size_t kSmemBudget = 220 * 1024;

static const double kChi2_005[99] = {
#include "chi2_table.inc"
};

struct EngineBase {
  virtual ~EngineBase() {}
  virtual int initialize(const void*, const void*, const void*, const void*) = 0;
  virtual int propagate(const void*) = 0;
  virtual int augment() = 0;
  virtual int update_async(int mode, const msckf_b200_tracks*) = 0;
  virtual int fetch(msckf_b200_report*) = 0;
  virtual int prune(const int*, int) = 0;
  virtual int get_state(void*, void*) = 0;
  virtual int get_covariance(void*) = 0;
  virtual int get_counters(long long*) = 0;
  virtual int last_dx(double*, int) = 0;
  virtual int copy_from(const EngineBase*) = 0;
  virtual int sync() = 0;
  virtual int stage(int mode, const msckf_b200_tracks*) = 0;
  virtual int launch() = 0;
  virtual int launch_timed(float* ms) = 0;
  virtual int kernel_times(float* ms, const char** names, int cap) = 0;
  virtual int tail_profile(unsigned long long* out, int cap) = 0;
  virtual void use_graph_reset() = 0;
  bool profile = false;
  bool use_graph = true;
  bool no_fused_tail = false;  // option 3 = 0: substitution as a separate sweep even where the fused form fits
  int dtype = 0, device = 0, Mmax = 0, Tmax = 0, Omax = 0;
  int M = 0;
  double rank_thr = 1e-11;
  long long launches = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;               // side branch of the update (k_blockdiag runs beside k_gram)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

template <class S>
struct Engine : EngineBase {
  // device
  mb::DevState<S>* d_st = nullptr;
  S *d_P = nullptr, *d_P2 = nullptr, *d_poses = nullptr, *d_poses2 = nullptr;
  int *d_off = nullptr, *d_idx = nullptr, *d_cm = nullptr, *d_tri = nullptr, *d_valid = nullptr, *d_src = nullptr,
      *d_accept = nullptr, *d_rows = nullptr, *d_rowoff = nullptr, *d_scratch = nullptr, *d_keep = nullptr, *d_m = nullptr,
      *d_rank = nullptr, *d_keepclones = nullptr, *d_cmeff = nullptr;
  unsigned long long* d_csnap = nullptr;
  unsigned long long* d_prof = nullptr;
  unsigned* d_done = nullptr;  // k_jac's CTA ticket counter
  S *d_obs = nullptr, *d_pfg = nullptr, *d_pfg_given = nullptr, *d_gamma = nullptr, *d_Xg = nullptr, *d_rg = nullptr,
    *d_Vg = nullptr, *d_taug = nullptr;
  double *d_Z = nullptr, *d_Yq = nullptr, *d_ur = nullptr, *d_G1p = nullptr, *d_G2p = nullptr, *d_D1 = nullptr, *d_D2 = nullptr,
         *d_bb = nullptr, *d_T2 = nullptr, *d_R2 = nullptr, *d_r2 = nullptr, *d_TP = nullptr, *d_S2 = nullptr, *d_W = nullptr, *d_G = nullptr,
         *d_y = nullptr, *d_dx = nullptr, *d_idiag = nullptr;
  // pinned host
  // One packed input block and one packed report block per update (a single H2D and a single D2H copy): the typed
  // pointers below (d_off .. d_pfg_given, d_cmeff .. d_gamma and their host mirrors) are carved out of them by layout(),
  // tightly for the batch at hand -- deterministic in (N, O), which the CUDA-graph key contains.
  unsigned char *d_in = nullptr, *h_in = nullptr, *d_rep = nullptr, *h_rep = nullptr;
  size_t in_bytes = 0, rep_bytes = 0;
  int *h_off = nullptr, *h_idx = nullptr, *h_cmeff = nullptr, *h_cm = nullptr, *h_tri = nullptr, *h_valid = nullptr, *h_accept = nullptr,
      *h_mr = nullptr /*m, rank*/;
  S *h_obs = nullptr, *h_pfg_in = nullptr, *h_pfg = nullptr, *h_gamma = nullptr;
  static size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }
  void layout(int N, int O) {
    size_t o = 0;
    auto take = [&](unsigned char* dbase, unsigned char* hbase, size_t bytes, void** dp, void** hp) {
      *dp = dbase + o; *hp = hbase + o; o += up16(bytes);
    };
    take(d_in, h_in, sizeof(int) * (N + 1), (void**)&d_off, (void**)&h_off);
    take(d_in, h_in, sizeof(int) * (size_t)O, (void**)&d_idx, (void**)&h_idx);
    take(d_in, h_in, sizeof(S) * 2 * (size_t)O, (void**)&d_obs, (void**)&h_obs);
    take(d_in, h_in, sizeof(S) * 3 * (size_t)N, (void**)&d_pfg_given, (void**)&h_pfg_in);
    in_bytes = o;
    o = 0;
    take(d_rep, h_rep, sizeof(int) * 2, (void**)&d_m, (void**)&h_mr);
    d_rank = d_m + 1;
    take(d_rep, h_rep, sizeof(int) * (size_t)N, (void**)&d_cmeff, (void**)&h_cmeff);
    take(d_rep, h_rep, sizeof(int) * (size_t)N, (void**)&d_cm, (void**)&h_cm);
    take(d_rep, h_rep, sizeof(int) * (size_t)N, (void**)&d_tri, (void**)&h_tri);
    take(d_rep, h_rep, sizeof(int) * (size_t)N, (void**)&d_valid, (void**)&h_valid);
    take(d_rep, h_rep, sizeof(int) * (size_t)N, (void**)&d_accept, (void**)&h_accept);
    take(d_rep, h_rep, sizeof(S) * 3 * (size_t)N, (void**)&d_pfg, (void**)&h_pfg);
    take(d_rep, h_rep, sizeof(S) * (size_t)N, (void**)&d_gamma, (void**)&h_gamma);
    rep_bytes = o;
  }
  mb::DevState<S>* h_st = nullptr;
  int nmax = 0, ldp = 0, ld = 0;
  int pending_n = 0, pending_mode = -1;
  bool initialized = false;
  // staged batch
  int st_N = 0, st_O = 0, st_Lmax = 0, st_mode = -1;
  bool staged = false;
  bool timed_region = false;
  // CUDA-graph replay of the update's kernel sequence when the batch signature repeats (launch-bound inner loop)
  cudaGraphExec_t g_exec = nullptr;
  unsigned long long g_key = ~0ull, last_key = ~0ull;
  int g_nodes = 0;
  void drop_graph() {
    if (g_exec) { cudaGraphExecDestroy(g_exec); g_exec = nullptr; }
    g_key = ~0ull; last_key = ~0ull;
  }
  // optional per-kernel CUDA-event profile of the last launch (option key 1)
  static constexpr int kMaxEv = 24;
  cudaEvent_t ev[kMaxEv] = {};
  const char* ev_name[kMaxEv] = {};
  int n_ev = 0;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  void mark(const char* name) {
    if (!profile || n_ev >= kMaxEv) return;
    if (!ev[n_ev]) cudaEventCreate(&ev[n_ev]);
    cudaEventRecord(ev[n_ev], stream);
    ev_name[n_ev] = name;
    n_ev++;
  }

  int alloc() {
    nmax = 15 + 6 * Mmax;
    ldp = (nmax + 3) & ~3;
    ld = (nmax + 1) & ~1;
    const int cmax = 6 * Mmax;
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    CK(cudaMalloc(&d_st, sizeof(mb::DevState<S>)));
    CK(cudaMalloc(&d_P, sizeof(S) * (size_t)ldp * nmax));
    CK(cudaMalloc(&d_P2, sizeof(S) * (size_t)ldp * nmax));
    CK(cudaMalloc(&d_poses, sizeof(S) * mb::kPoseStride * (size_t)(Mmax + 1)));
    CK(cudaMalloc(&d_poses2, sizeof(S) * mb::kPoseStride * (size_t)(Mmax + 1)));
    CK(cudaMemsetAsync(d_P, 0, sizeof(S) * (size_t)ldp * nmax, stream));
    CK(cudaMemsetAsync(d_P2, 0, sizeof(S) * (size_t)ldp * nmax, stream));
    CK(cudaMemsetAsync(d_poses, 0, sizeof(S) * mb::kPoseStride * (size_t)(Mmax + 1), stream));
    CK(cudaMemsetAsync(d_poses2, 0, sizeof(S) * mb::kPoseStride * (size_t)(Mmax + 1), stream));
    const size_t T = Tmax, O = Omax;
    for (int** p : {&d_src, &d_rows, &d_scratch}) CK(cudaMalloc(p, sizeof(int) * T));
    {
      const size_t cap_in = up16(sizeof(int) * (T + 1)) + up16(sizeof(int) * O) + up16(sizeof(S) * 2 * O) + up16(sizeof(S) * 3 * T);
      const size_t cap_rep = up16(sizeof(int) * 2) + 5 * up16(sizeof(int) * T) + up16(sizeof(S) * 3 * T) + up16(sizeof(S) * T);
      CK(cudaMalloc(&d_in, cap_in)); CK(cudaMallocHost(&h_in, cap_in));
      CK(cudaMalloc(&d_rep, cap_rep)); CK(cudaMallocHost(&h_rep, cap_rep));
      CK(cudaMemsetAsync(d_rep, 0, cap_rep, stream));
      layout((int)T, (int)O);
    }
    CK(cudaMalloc(&d_rowoff, sizeof(int) * (T + 1)));
    CK(cudaMalloc(&d_csnap, sizeof(unsigned long long)));
    CK(cudaMalloc(&d_prof, sizeof(unsigned long long) * 80));
    CK(cudaMalloc(&d_done, sizeof(unsigned)));
    CK(cudaMemsetAsync(d_done, 0, sizeof(unsigned), stream));
    CK(cudaMemsetAsync(d_prof, 0, sizeof(unsigned long long) * 80, stream));
    CK(cudaMalloc(&d_keep, sizeof(int) * nmax));
    CK(cudaMalloc(&d_keepclones, sizeof(int) * (Mmax + 1)));
    CK(cudaMalloc(&d_Xg, sizeof(S) * 12 * O));
    CK(cudaMalloc(&d_rg, sizeof(S) * 2 * O));
    CK(cudaMalloc(&d_Vg, sizeof(S) * 6 * O));
    CK(cudaMalloc(&d_taug, sizeof(S) * 3 * T));
    CK(cudaMalloc(&d_Z, sizeof(double) * 3 * T * cmax));
    CK(cudaMalloc(&d_Yq, sizeof(double) * 3 * T * cmax));
    CK(cudaMalloc(&d_ur, sizeof(double) * 3 * T));
    CK(cudaMalloc(&d_G1p, sizeof(double) * kMaxSplit * (size_t)cmax * cmax));
    CK(cudaMalloc(&d_G2p, sizeof(double) * kMaxSplit * (size_t)cmax * cmax));
    CK(cudaMalloc(&d_D1, sizeof(double) * 36 * Mmax));
    CK(cudaMalloc(&d_D2, sizeof(double) * 36 * Mmax));
    CK(cudaMalloc(&d_bb, sizeof(double) * 6 * Mmax));
    for (double** p : {&d_T2, &d_R2, &d_TP, &d_S2, &d_W, &d_G}) CK(cudaMalloc(p, sizeof(double) * (size_t)ld * nmax));
    CK(cudaMalloc(&d_r2, sizeof(double) * nmax));
    CK(cudaMalloc(&d_idiag, sizeof(double) * 2 * nmax));
    CK(cudaMalloc(&d_y, sizeof(double) * nmax));
    CK(cudaMalloc(&d_dx, sizeof(double) * nmax));
    CK(cudaMemsetAsync(d_dx, 0, sizeof(double) * nmax, stream));
    CK(cudaMallocHost(&h_st, sizeof(mb::DevState<S>)));
    // opt in to large dynamic shared memory
    CK(cudaFuncSetAttribute(mb::k_tri<S, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    CK(cudaFuncSetAttribute(mb::k_jac<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    CK(cudaFuncSetAttribute(mb::k_tail_fused<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    CK(cudaFuncSetAttribute(mb::k_tail<S, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    CK(cudaFuncSetAttribute(mb::k_tail<S, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }
  ~Engine() override {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (g_exec) cudaGraphExecDestroy(g_exec);
    void* dv[] = {d_st, d_P, d_P2, d_poses, d_poses2, d_in, d_rep, d_src, d_rows, d_rowoff,
                  d_scratch, d_keep, d_keepclones, d_csnap, d_prof, d_done, d_Xg, d_rg, d_Vg, d_taug, d_Z, d_Yq,
                  d_ur, d_G1p, d_G2p, d_D1, d_D2, d_bb, d_T2, d_R2, d_TP, d_S2, d_W, d_G, d_r2, d_y, d_dx, d_idiag};
    for (void* p : dv) if (p) cudaFree(p);
    void* hv[] = {h_in, h_rep, h_st};
    for (void* p : hv) if (p) cudaFreeHost(p);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (stream2) cudaStreamDestroy(stream2);
    if (stream) cudaStreamDestroy(stream);
  }

  int initialize(const void* cam_, const void* noise_, const void* params_, const void* imu_) override {
    const S* cam = (const S*)cam_; const S* nz = (const S*)noise_; const S* pr = (const S*)params_; const S* im = (const S*)imu_;
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(stream));
    mb::DevState<S>& s = *h_st;
    memset(&s, 0, sizeof(s));
    for (int i = 0; i < 4; ++i) s.q_CI[i] = cam[i];
    for (int i = 0; i < 3; ++i) s.p_C_I[i] = cam[4 + i];
    s.u_var = nz[0]; s.v_var = nz[1];
    for (int i = 0; i < 144; ++i) s.Q_imu[i] = nz[2 + i];
    s.max_gn_cost_norm = pr[0]; s.translation_threshold = pr[1];
    for (int i = 0; i < 3; ++i) {
      s.p_I_G[i] = im[i]; s.v_I_G[i] = im[3 + i]; s.b_g[i] = im[6 + i]; s.b_a[i] = im[9 + i]; s.g[i] = im[12 + i];
      s.p_I_G_null[i] = im[i]; s.v_I_G_null[i] = im[3 + i];
    }
    for (int i = 0; i < 4; ++i) { s.q_IG[i] = im[15 + i]; s.q_IG_null[i] = im[15 + i]; }
    for (int i = 0; i < 99; ++i) s.chi2[i] = (S)kChi2_005[i];  // msckf.h:91-95
    CK(cudaMemcpyAsync(d_st, h_st, sizeof(s), cudaMemcpyHostToDevice, stream));
    // imu_covar_ = initial_imu_covar (msckf.h:86)
    std::vector<S> P0((size_t)ldp * 15, S(0));
    for (int i = 0; i < 15; ++i)
      for (int j = 0; j < 15; ++j) P0[(size_t)i * ldp + j] = nz[146 + 15 * i + j];
    CK(cudaMemsetAsync(d_P, 0, sizeof(S) * (size_t)ldp * nmax, stream));
    CK(cudaMemcpyAsync(d_P, P0.data(), sizeof(S) * P0.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    M = 0;
    initialized = true;
    pending_mode = -1;
    drop_graph();
    return 0;
  }

  int propagate(const void* r_) override {
    if (!initialized) return fail(MSCKF_B200_ERR_STATE, "propagate before initialize");
    const S* r = (const S*)r_;
    CK(cudaSetDevice(device));
    mb::k_propagate<S><<<1, 256, 0, stream>>>(d_st, d_P, ldp, M, r[0], r[1], r[2], r[3], r[4], r[5], r[6]);
    launches++;
    CK(cudaGetLastError());
    return 0;
  }

  int augment() override {
    if (!initialized) return fail(MSCKF_B200_ERR_STATE, "augment before initialize");
    if (M + 1 > Mmax) return fail(MSCKF_B200_ERR_CAPACITY, "augmentState: clone capacity exceeded");
    CK(cudaSetDevice(device));
    mb::k_augment<S><<<1, 256, 0, stream>>>(d_st, d_P, ldp, M, d_poses);
    launches++;
    CK(cudaGetLastError());
    M += 1;
    return 0;
  }

  int update_async(int mode, const msckf_b200_tracks* tr) override {
    int rc = stage(mode, tr);
    if (rc != 0) return rc;
    return launch();
  }

  // validate + copy the batch to the device (pinned staging, asynchronous)
  int stage(int mode, const msckf_b200_tracks* tr) override {
    if (!initialized) return fail(MSCKF_B200_ERR_STATE, "update before initialize");
    if (pending_mode >= 0) return fail(MSCKF_B200_ERR_STATE, "previous update not fetched");
    const int N = tr->n_tracks;
    if (N < 0 || N > Tmax) return fail(MSCKF_B200_ERR_CAPACITY, "track batch exceeds max_tracks");
    CK(cudaSetDevice(device));
    staged = false;
    st_N = N; st_mode = mode; st_O = 0; st_Lmax = 0;
    if (N == 0) { staged = true; return 0; }
    if (M < 1) return fail(MSCKF_B200_ERR_STATE, "update without clones");
    const int O = tr->obs_offset[N];
    if (O > Omax) return fail(MSCKF_B200_ERR_CAPACITY, "track batch exceeds max_obs");
    int Lmax = 0;
    for (int t = 0; t < N; ++t) {
      const int L = tr->obs_offset[t + 1] - tr->obs_offset[t];
      if (L < 1 || L > 98) return fail(MSCKF_B200_ERR_ARG, "track length must be in [1,98] (chi-square table, msckf.h:91)");
      if (mode != MSCKF_B200_TRIANGULATE && L < 2) return fail(MSCKF_B200_ERR_ARG, "residualised track needs >= 2 observations");
      Lmax = std::max(Lmax, L);
    }
    for (int o = 0; o < O; ++o)
      if (tr->clone_index[o] < 0 || tr->clone_index[o] >= M) return fail(MSCKF_B200_ERR_ARG, "clone_index out of range");
    if (mode == MSCKF_B200_RESIDUALIZE && !tr->p_f_G) return fail(MSCKF_B200_ERR_ARG, "RESIDUALIZE needs p_f_G");
    layout(N, O);
    memcpy(h_off, tr->obs_offset, sizeof(int) * (N + 1));
    memcpy(h_idx, tr->clone_index, sizeof(int) * O);
    memcpy(h_obs, tr->obs, sizeof(S) * 2 * (size_t)O);
    if (mode == MSCKF_B200_RESIDUALIZE) memcpy(h_pfg_in, tr->p_f_G, sizeof(S) * 3 * (size_t)N);
    CK(cudaMemcpyAsync(d_in, h_in, mode == MSCKF_B200_RESIDUALIZE ? in_bytes : (size_t)((unsigned char*)d_pfg_given - d_in), cudaMemcpyHostToDevice, stream));
    st_O = O; st_Lmax = Lmax;
    staged = true;
    return 0;
  }

  // launch the kernels of the staged batch and queue the report copies (all asynchronous)
  int launch() override {
    if (!staged) return fail(MSCKF_B200_ERR_STATE, "launch without a staged batch");
    staged = false;
    const int N = st_N, mode = st_mode, Lmax = st_Lmax;
    CK(cudaSetDevice(device));
    pending_n = N;
    pending_mode = mode;
    n_ev = 0;
    if (N == 0) return 0;
    const unsigned long long key = ((unsigned long long)mode << 60) ^ ((unsigned long long)N << 40) ^ ((unsigned long long)st_O << 16) ^
                                   ((unsigned long long)Lmax << 8) ^ (unsigned long long)M;
    bool capturing = false;
    const long long launches_before = launches;
    if (use_graph && !profile) {
      if (g_exec && key == g_key) {
        CK(cudaGraphLaunch(g_exec, stream));
        launches += g_nodes;
        return queue_report(N, mode);
      }
      if (key == last_key) {  // second time in a row with this signature: capture it
        CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        capturing = true;
      }
      last_key = key;
    }
    mark("begin");
    const int n = 15 + 6 * M, c = 6 * M;
    mb::FeatArgs<S> a;
    a.n_tracks = N; a.M = M; a.Lmax = Lmax; a.ldp = ldp; a.has_yf = 1;
    a.obs_off = d_off; a.obs = d_obs; a.clone_idx = d_idx; a.poses = d_poses; a.P = d_P; a.st = d_st;
    a.pfg = d_pfg; a.counter_snap = d_csnap; a.cm_eff = d_cmeff; a.cm_ok = d_cm; a.tri_ok = d_tri; a.valid = d_valid; a.src = d_src;
    a.pfg_given = (mode == MSCKF_B200_RESIDUALIZE) ? d_pfg_given : nullptr;
    a.accept = d_accept; a.gamma = d_gamma; a.rows = d_rows; a.Xg = d_Xg; a.rg = d_rg; a.Vg = d_Vg; a.taug = d_taug;
    a.Z = d_Z; a.Yq = d_Yq; a.ur = d_ur;
    a.row_off = d_rowoff; a.m_out = d_m; a.done = d_done;
    a.prof = profile ? (d_prof + 40) : nullptr;
    const size_t pose_bytes = 16 + sizeof(S) * mb::kPoseStride * (size_t)M;
    if (mode != MSCKF_B200_RESIDUALIZE) {
      const size_t smem = pose_bytes + sizeof(S) * 4 * 14 * (size_t)Lmax;
      if (smem > kSmemBudget) return fail(MSCKF_B200_ERR_CAPACITY, "k_tri shared memory");
      mb::k_tri<S, 4><<<(N + 3) / 4, 128, smem, stream>>>(a);
      launches++;
      mark("k_tri");
    }
    if (mode != MSCKF_B200_TRIANGULATE) {
      size_t jsmem = mb::jac_smem_bytes<S>(Lmax, M, true);
      a.has_yf = 1;
      if (jsmem > kSmemBudget) { jsmem = mb::jac_smem_bytes<S>(Lmax, M, false); a.has_yf = 0; }  // longest fp64 tracks: CTA-wide gate only
      if (jsmem > kSmemBudget) return fail(MSCKF_B200_ERR_CAPACITY, "k_jac shared memory");
      mb::k_jac<S><<<N, mb::JT, jsmem, stream>>>(a, d_st, mode == MSCKF_B200_RESIDUALIZE ? 1 : 0);
      launches++;
      mark("k_jac");
      const double du = (double)h_st->u_var, dv = (double)h_st->v_var;
      // k_blockdiag and k_gram both depend on k_jac only: two branches (a fork / join in the captured graph)
      cudaStream_t sb = profile ? stream : stream2;  // per-kernel event timing keeps everything on one stream
      if (!profile) { CK(cudaEventRecord(ev_fork, stream)); CK(cudaStreamWaitEvent(stream2, ev_fork, 0)); }
      mb::k_blockdiag<S><<<M, 128, 0, sb>>>(N, d_off, d_idx, d_accept, d_Xg, d_rg, du, dv, d_D1, d_D2, d_bb);
      launches++;
      mark("k_blockdiag");
      if (!profile) CK(cudaEventRecord(ev_join, stream2));
      const int K = 3 * N;
      int nsplit = std::max(1, std::min(kMaxSplit, K / 96));
      int kchunk = (K + nsplit - 1) / nsplit;
      kchunk = (kchunk + mb::GK - 1) / mb::GK * mb::GK;
      nsplit = (K + kchunk - 1) / kchunk;
      const int ntile = (c + mb::GT - 1) / mb::GT;
      mb::k_gram<<<dim3(ntile * (ntile + 1) / 2, nsplit), 256, 0, stream>>>(d_Z, d_Yq, K, c, kchunk, d_G1p, d_G2p);
      launches++;
      mark("k_gram");
      if (!profile) CK(cudaStreamWaitEvent(stream, ev_join, 0));
      const int agrid = std::min(592, (n * n + 255) / 256);
      mb::k_assemble<<<agrid, 256, 0, stream>>>(n, ld, K, nsplit, d_G1p, d_G2p, d_D1, d_D2, d_bb, d_Z, d_ur, d_m, d_T2, d_R2, d_r2);
      launches++;
      mark("k_assemble");
      mb::k_rows<S><<<N, 128, sizeof(double) * 12 * (size_t)Lmax, stream>>>(N, n, ld, d_off, d_idx, d_accept, d_rowoff, d_m, d_Xg, d_rg, d_Vg,
                                                                            d_taug, d_Z, du, dv, d_T2, d_R2, d_r2);
      launches++;
      mark("k_rows");
      const dim3 tg((n + 31) / 32, (n + 31) / 32);
      mb::k_gemm_tp<S><<<tg, mb::kGemmThreads, 0, stream>>>(n, ld, d_T2, d_P, ldp, d_TP);
      mark("k_gemm_tp");
      mb::k_gemm_s<<<tg, mb::kGemmThreads, 0, stream>>>(n, ld, d_TP, d_T2, d_R2, d_S2);
      mark("k_gemm_s");
      launches += 2;
      // rank decision + Cholesky + substitution + covariance/state update: one cluster kernel (scratch for Gamma: d_G)
      {
        const int ldt = (n + 3) & ~3;
        auto smem_for = [&](int NB, bool) {
          return sizeof(double) * ((size_t)2 * NB * (NB + 1) + 2 + (size_t)2 * NB * ldt + ((n + 1) & ~1) + 2 * NB) + 64;
        };
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kTailCluster, 1, 1);
        cfg.blockDim = dim3(mb::kTailThreads, 1, 1);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kTailCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        unsigned long long* pf = profile ? d_prof : (unsigned long long*)nullptr;
        // fused substitution: the RHS columns of one CTA must fit one block of W rows ((n + 1) / cluster <= 32)
        const bool fused = (n + 1 + kTailCluster - 1) / kTailCluster <= 32 && mb::tail_fused_smem_bytes(n) <= kSmemBudget && !no_fused_tail;
        if (fused) {
          cfg.dynamicSmemBytes = mb::tail_fused_smem_bytes(n);
          // T'' is dead after k_gemm_s: the kernel patches it into Gamma in place
          CK(cudaLaunchKernelEx(&cfg, mb::k_tail_fused<S>, n, ld, d_T2, d_S2, rank_thr, d_rank, (const int*)d_m,
                                (const double*)d_TP, (const double*)d_r2, d_W, d_y, d_dx, d_G /*scratch*/, pf));
        } else if (smem_for(32, false) <= kSmemBudget) {
          cfg.dynamicSmemBytes = smem_for(32, false);
          CK(cudaLaunchKernelEx(&cfg, mb::k_tail<S, 32>, n, ld, M, (const double*)d_T2, d_G, d_S2, d_R2 /*receives L (R'' is consumed by k_gemm_s)*/,
                                d_keep, d_idiag, rank_thr, d_rank, (const int*)d_m, (const double*)d_TP, (const double*)d_r2, d_W, d_y, d_P, ldp, d_st,
                                d_poses, d_dx, pf));
        } else if (smem_for(16, false) <= kSmemBudget) {
          cfg.dynamicSmemBytes = smem_for(16, false);
          CK(cudaLaunchKernelEx(&cfg, mb::k_tail<S, 16>, n, ld, M, (const double*)d_T2, d_G, d_S2, d_R2, d_keep, d_idiag, rank_thr, d_rank,
                                (const int*)d_m, (const double*)d_TP, (const double*)d_r2, d_W, d_y, d_P, ldp, d_st, d_poses, d_dx, pf));
        } else return fail(MSCKF_B200_ERR_CAPACITY, "k_tail shared memory");
      }
      launches++;
      mark("k_tail");
      {
        const int nt32 = (n + 31) / 32;
        mb::k_syrk<S><<<nt32 * (nt32 + 1) / 2, mb::kGemmThreads, 0, stream>>>(n, ld, d_W, d_P, ldp, d_m);
        launches++;
        mark("k_syrk");
        mb::k_inject<S><<<1, 1024, sizeof(double) * n, stream>>>(n, ld, M, d_W, d_y, d_st, d_poses, d_dx, d_m, d_rank);
        launches++;
        mark("k_inject");
      }
    }
    if (capturing) {
      cudaGraph_t graph = nullptr;
      CK(cudaStreamEndCapture(stream, &graph));
      if (g_exec) { cudaGraphExecDestroy(g_exec); g_exec = nullptr; }
      CK(cudaGraphInstantiate(&g_exec, graph, 0));
      CK(cudaGraphDestroy(graph));
      g_key = key;
      g_nodes = (int)(launches - launches_before);
      CK(cudaGraphLaunch(g_exec, stream));
    }
    CK(cudaGetLastError());
    return queue_report(N, mode);
  }

  int queue_report(int N, int mode) {
    if (timed_region) CK(cudaEventRecord(ev_t1, stream));
    // report back (pinned), still asynchronous: one copy of the packed block
    (void)mode;
    if (N > 0) CK(cudaMemcpyAsync(h_rep, d_rep, rep_bytes, cudaMemcpyDeviceToHost, stream));
    return 0;
  }

  int fetch(msckf_b200_report* rep) override {
    if (pending_mode < 0) return fail(MSCKF_B200_ERR_STATE, "fetch without a pending update");
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(stream));
    const int N = pending_n, mode = pending_mode;
    pending_mode = -1;
    if (!rep) return 0;
    rep->m = 0; rep->rank = 0;
    if (N == 0) return 0;
    if (mode != MSCKF_B200_RESIDUALIZE) {
      if (rep->cm_ok) memcpy(rep->cm_ok, mode == MSCKF_B200_MARGINALIZE ? h_cmeff : h_cm, sizeof(int) * N);
      if (rep->tri_ok) memcpy(rep->tri_ok, h_tri, sizeof(int) * N);
      if (rep->p_f_G) memcpy(rep->p_f_G, h_pfg, sizeof(S) * 3 * (size_t)N);
    }
    if (mode != MSCKF_B200_TRIANGULATE) {
      if (rep->valid) memcpy(rep->valid, h_valid, sizeof(int) * N);
      if (rep->accepted) memcpy(rep->accepted, h_accept, sizeof(int) * N);
      if (rep->gamma) memcpy(rep->gamma, h_gamma, sizeof(S) * N);
      rep->m = h_mr[0];
      rep->rank = h_mr[1];
    }
    return 0;
  }

  int prune(const int* keep, int n_keep) override {
    if (n_keep < 0 || n_keep > M) return fail(MSCKF_B200_ERR_ARG, "prune: bad keep count");
    for (int i = 0; i < n_keep; ++i)
      if (keep[i] < 0 || keep[i] >= M || (i && keep[i] <= keep[i - 1])) return fail(MSCKF_B200_ERR_ARG, "prune: keep[] must be ascending positions");
    if (n_keep == M) return 0;
    CK(cudaSetDevice(device));
    CK(cudaMemcpyAsync(d_keepclones, keep, sizeof(int) * std::max(n_keep, 1), cudaMemcpyHostToDevice, stream));
    const int n_new = 15 + 6 * n_keep;
    mb::k_gather<S><<<std::min(296, (n_new * n_new + 255) / 256), 256, 0, stream>>>(n_new, d_keepclones, n_keep, d_P, d_P2, ldp, d_poses, d_poses2);
    launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));  // keep[] is caller memory
    std::swap(d_P, d_P2);
    std::swap(d_poses, d_poses2);
    drop_graph();  // captured kernel arguments hold the old buffers
    M = n_keep;
    return 0;
  }

  int get_state(void* imu_, void* poses_) override {
    CK(cudaSetDevice(device));
    std::vector<S> tmp;
    if (imu_) CK(cudaMemcpyAsync(h_st, d_st, sizeof(mb::DevState<S>), cudaMemcpyDeviceToHost, stream));
    if (poses_ && M > 0) {
      tmp.resize((size_t)mb::kPoseStride * M);
      CK(cudaMemcpyAsync(tmp.data(), d_poses, sizeof(S) * tmp.size(), cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));  // one synchronisation for both copies
    if (imu_) {
      S* o = (S*)imu_;
      const mb::DevState<S>& s = *h_st;
      for (int i = 0; i < 3; ++i) {
        o[i] = s.p_I_G[i]; o[3 + i] = s.v_I_G[i]; o[6 + i] = s.b_g[i]; o[9 + i] = s.b_a[i]; o[12 + i] = s.g[i];
        o[19 + i] = s.p_I_G_null[i]; o[22 + i] = s.v_I_G_null[i];
      }
      for (int i = 0; i < 4; ++i) { o[15 + i] = s.q_IG[i]; o[25 + i] = s.q_IG_null[i]; }
    }
    if (poses_ && M > 0) {
      S* o = (S*)poses_;
      for (int k = 0; k < M; ++k) {
        const S* p = tmp.data() + mb::kPoseStride * k;
        o[7 * k] = p[4]; o[7 * k + 1] = p[5]; o[7 * k + 2] = p[6];
        o[7 * k + 3] = p[0]; o[7 * k + 4] = p[1]; o[7 * k + 5] = p[2]; o[7 * k + 6] = p[3];
      }
    }
    return 0;
  }

  int get_covariance(void* out) override {
    CK(cudaSetDevice(device));
    const int n = 15 + 6 * M;
    CK(cudaMemcpy2DAsync(out, sizeof(S) * n, d_P, sizeof(S) * ldp, sizeof(S) * n, n, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return n;
  }

  int get_counters(long long* c) override {
    CK(cudaSetDevice(device));
    CK(cudaMemcpyAsync(h_st, d_st, sizeof(mb::DevState<S>), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    c[0] = (long long)h_st->num_residualized; c[1] = (long long)h_st->pfg_shifted; c[2] = (long long)h_st->pfg_oob;
    c[3] = (long long)h_st->n_updates; c[4] = h_st->last_m; c[5] = h_st->last_rank; c[6] = 0; c[7] = 0;
    return 0;
  }

  int last_dx(double* out, int cap) override {
    CK(cudaSetDevice(device));
    const int n = 15 + 6 * M;
    std::vector<double> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), d_dx, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int i = 0; i < std::min(n, cap); ++i) out[i] = tmp[i];
    return n;
  }

  int copy_from(const EngineBase* src_) override {
    const Engine<S>* src = dynamic_cast<const Engine<S>*>(src_);
    if (!src || src->Mmax != Mmax || src->device != device) return fail(MSCKF_B200_ERR_ARG, "copy_state: incompatible engines");
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(src->stream));
    CK(cudaMemcpyAsync(d_st, src->d_st, sizeof(mb::DevState<S>), cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(d_P, src->d_P, sizeof(S) * (size_t)ldp * nmax, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(d_poses, src->d_poses, sizeof(S) * mb::kPoseStride * (size_t)(Mmax + 1), cudaMemcpyDeviceToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    // captured kernel arguments hold scalars of the filter (noise variances, rank threshold): keep the graph only if equal
    if (M != src->M || rank_thr != src->rank_thr || h_st->u_var != src->h_st->u_var || h_st->v_var != src->h_st->v_var) drop_graph();
    rank_thr = src->rank_thr;
    pending_mode = -1;
    memcpy(h_st, src->h_st, sizeof(mb::DevState<S>));
    M = src->M;
    initialized = src->initialized;
    return 0;
  }

  int sync() override {
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }

  // launch() bracketed by CUDA events on this handle's stream; the staged inputs are already in HBM.
  // The report copies are queued after the stop event, so the time is kernels only.
  int launch_timed(float* ms) override {
    CK(cudaSetDevice(device));
    if (!ev_t0) { CK(cudaEventCreate(&ev_t0)); CK(cudaEventCreate(&ev_t1)); }
    CK(cudaStreamSynchronize(stream));
    timed_region = true;
    CK(cudaEventRecord(ev_t0, stream));
    int rc = launch();
    timed_region = false;
    if (rc != 0) return rc;
    CK(cudaStreamSynchronize(stream));
    CK(cudaEventElapsedTime(ms, ev_t0, ev_t1));
    return 0;
  }

  void use_graph_reset() override { drop_graph(); }

  int tail_profile(unsigned long long* out, int cap) override {
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(stream));
    CK(cudaMemcpy(out, d_prof, sizeof(unsigned long long) * std::min(cap, 80), cudaMemcpyDeviceToHost));
    return std::min(cap, 80);
  }

  int kernel_times(float* ms, const char** names, int cap) override {
    CK(cudaSetDevice(device));
    CK(cudaStreamSynchronize(stream));
    int k = 0;
    for (int i = 1; i < n_ev && k < cap; ++i, ++k) {
      CK(cudaEventElapsedTime(&ms[k], ev[i - 1], ev[i]));
      names[k] = ev_name[i];
    }
    return k;
  }
};
}  // namespace

struct msckf_b200_engine { EngineBase* impl; };

extern "C" {
const char* msckf_b200_last_error(void) { return g_err.c_str(); }

int msckf_b200_create(const msckf_b200_config* cfg, msckf_b200_engine** out) {
  if (!cfg || !out) return fail(MSCKF_B200_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MSCKF_B200_ERR_NO_DEVICE, "no CUDA device: the B200 engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(MSCKF_B200_ERR_ARG, "bad device ordinal");
  if (cfg->max_clones < 1 || cfg->max_tracks < 1 || cfg->max_obs < 1) return fail(MSCKF_B200_ERR_ARG, "bad capacities");
  EngineBase* b = nullptr;
  if (cfg->dtype == MSCKF_B200_F32) b = new Engine<float>();
  else if (cfg->dtype == MSCKF_B200_F64) b = new Engine<double>();
  else return fail(MSCKF_B200_ERR_ARG, "bad dtype");
  if (getenv("MSCKF_B200_NO_GRAPH")) b->use_graph = false;  // e.g. under ncu: profile plain launches
  if (getenv("MSCKF_B200_NO_FUSED_TAIL")) b->no_fused_tail = true;
  b->dtype = cfg->dtype; b->device = cfg->device; b->Mmax = cfg->max_clones; b->Tmax = cfg->max_tracks; b->Omax = cfg->max_obs;
  int rc = (cfg->dtype == MSCKF_B200_F32) ? static_cast<Engine<float>*>(b)->alloc() : static_cast<Engine<double>*>(b)->alloc();
  if (rc != 0) { delete b; return rc; }
  *out = new msckf_b200_engine{b};
  return 0;
}
int msckf_b200_destroy(msckf_b200_engine* e) {
  if (!e) return 0;
  delete e->impl;
  delete e;
  return 0;
}
int msckf_b200_initialize(msckf_b200_engine* e, const void* camera, const void* noise, const void* params, const void* imu_state) {
  return e->impl->initialize(camera, noise, params, imu_state);
}
int msckf_b200_propagate(msckf_b200_engine* e, const void* reading) { return e->impl->propagate(reading); }
int msckf_b200_augment(msckf_b200_engine* e) { return e->impl->augment(); }
int msckf_b200_update_async(msckf_b200_engine* e, int mode, const msckf_b200_tracks* tracks) { return e->impl->update_async(mode, tracks); }
int msckf_b200_fetch(msckf_b200_engine* e, msckf_b200_report* report) { return e->impl->fetch(report); }
int msckf_b200_update(msckf_b200_engine* e, int mode, const msckf_b200_tracks* tracks, msckf_b200_report* report) {
  int rc = e->impl->update_async(mode, tracks);
  if (rc != 0) { e->impl->fetch(nullptr); return rc; }
  return e->impl->fetch(report);
}
int msckf_b200_update_batch(msckf_b200_engine** engines, int n, int mode, const msckf_b200_tracks* tracks, msckf_b200_report* reports,
                            int threads) {
  if (n <= 0) return 0;
  if (!engines || !tracks) return fail(MSCKF_B200_ERR_ARG, "null argument");
  const int T = std::max(1, std::min(threads, n));
  std::vector<int> rc(T, 0);
  std::vector<std::string> err(T);
  auto work = [&](int w) {
    for (int i = w; i < n; i += T) {
      const int r = engines[i]->impl->update_async(mode, &tracks[i]);
      if (r != 0 && rc[w] == 0) { rc[w] = r; err[w] = g_err; }
    }
    for (int i = w; i < n; i += T) {
      const int r = engines[i]->impl->fetch(reports ? &reports[i] : nullptr);
      if (r != 0 && rc[w] == 0) { rc[w] = r; err[w] = g_err; }
    }
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < T; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto& t : pool) t.join();
  for (int w = 0; w < T; ++w)
    if (rc[w] != 0) return fail(rc[w], err[w]);
  return 0;
}
int msckf_b200_prune(msckf_b200_engine* e, const int* keep, int n_keep) { return e->impl->prune(keep, n_keep); }
int msckf_b200_num_clones(msckf_b200_engine* e) { return e->impl->M; }
int msckf_b200_get_state(msckf_b200_engine* e, void* imu, void* clone_poses) { return e->impl->get_state(imu, clone_poses); }
int msckf_b200_get_covariance(msckf_b200_engine* e, void* out) { return e->impl->get_covariance(out); }
int msckf_b200_get_counters(msckf_b200_engine* e, long long* counters) { return e->impl->get_counters(counters); }
int msckf_b200_last_delta_x(msckf_b200_engine* e, double* out, int cap) { return e->impl->last_dx(out, cap); }
int msckf_b200_stage(msckf_b200_engine* e, int mode, const msckf_b200_tracks* tracks) { return e->impl->stage(mode, tracks); }
int msckf_b200_launch(msckf_b200_engine* e) { return e->impl->launch(); }
int msckf_b200_launch_timed(msckf_b200_engine* e, float* ms) { return e->impl->launch_timed(ms); }
int msckf_b200_kernel_times(msckf_b200_engine* e, float* ms, const char** names, int cap) { return e->impl->kernel_times(ms, names, cap); }
int msckf_b200_tail_profile(msckf_b200_engine* e, unsigned long long* out, int cap) { return e->impl->tail_profile(out, cap); }
int msckf_b200_set_option(msckf_b200_engine* e, int key, double value) {
  if (key == 0) { e->impl->rank_thr = value; e->impl->use_graph_reset(); return 0; }
  if (key == 1) { e->impl->profile = value != 0; return 0; }
  if (key == 2) { e->impl->use_graph = value != 0; e->impl->use_graph_reset(); return 0; }
  if (key == 3) { e->impl->no_fused_tail = value == 0; e->impl->use_graph_reset(); return 0; }
  return fail(MSCKF_B200_ERR_ARG, "unknown option");
}
int msckf_b200_copy_state(msckf_b200_engine* dst, const msckf_b200_engine* src) { return dst->impl->copy_from(src->impl); }
long long msckf_b200_launch_count(const msckf_b200_engine* e) { return e->impl->launches; }
void* msckf_b200_stream(msckf_b200_engine* e) { return (void*)e->impl->stream; }
int msckf_b200_synchronize(msckf_b200_engine* e) { return e->impl->sync(); }
}
