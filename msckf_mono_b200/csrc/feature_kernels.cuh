// msckf_mono_b200/csrc/feature_kernels.cuh
// Per-feature stage of the MSCKF measurement update, one warp per feature track:
//   k_tri     : checkMotion (msckf.h:980-1025) + inverse-depth LM triangulation (msckf.h:1147-1285)
//   k_resolve : loop-A bookkeeping of marginalize() (msckf.h:352-399, incl. the p_f_G_vec index quirk of :419)
//   k_jac     : calcResidual (:960-978), calcMeasJacobian (:905-958, left null space by 3 Householder
//               reflectors of the column-pivoted QR of H_f), gatingTest (:1103-1124)
//   k_scan    : ordered stacking offsets (msckf.h:433-445)
// Clone poses are staged into shared memory with one TMA bulk copy per CTA.
#pragma once
#include "common.cuh"

namespace mb {

template <class S>
struct FeatArgs {
  int n_tracks, M, Lmax, ldp;
  const int* obs_off;    // [N+1]
  const S* obs;          // [sumL*2] normalised image coordinates
  const int* clone_idx;  // [sumL] positional index of the observing clone
  const S* poses;        // [M*8] current clone poses (q xyzw, p, pad)
  const S* P;            // [n x n], leading dim ldp, prior covariance
  const DevState<S>* st;
  // k_tri out
  S* pfg;       // [N*3]
  int* cm_ok;   // [N] checkMotion result
  int* tri_ok;  // [N] initializePosition validity
  // k_resolve out
  int* valid;  // [N]
  int* src;    // [N] index of the track whose p_f_G loop B uses (msckf.h:419)
  // k_jac out
  const S* pfg_given;  // optional [N*3]: residualize tracks at given positions (pruneRedundantStates)
  int* accept;         // [N]
  S* gamma;            // [N]
  int* rows;           // [N] rho_j = 2L-3
  S* Xg;               // [sumL*12] H_x blocks (2x6 per observation)
  S* rg;               // [sumL*2]  residuals
  S* Vg;               // [sumL*2*3] Householder vectors of the null-space projection (unit lower trapezoid)
  S* taug;             // [N*3]
  double* Z;           // [3N x c]  U_j^T X_j scattered to clone columns
  double* Yq;          // [3N x c]  U_j^T D X_j - 1/2 (U_j^T D U_j) Z_j
  double* ur;          // [3N]      U_j^T r_j
};

// Eigen::LDLT (diagonal pivoting) solve of a symmetric 3x3 system, msckf.h:1222.
template <class S>
__device__ __forceinline__ void ldlt3_solve(S A[3][3], const S b[3], S x[3]) {
  int tr[3];
  S temp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int idx = k;
    S big = tabs(A[k][k]);
    for (int i = k + 1; i < 3; ++i)
      if (tabs(A[i][i]) > big) { big = tabs(A[i][i]); idx = i; }
    tr[k] = idx;
    if (idx != k) {
      for (int j = 0; j < k; ++j) { S t = A[k][j]; A[k][j] = A[idx][j]; A[idx][j] = t; }
      for (int i = idx + 1; i < 3; ++i) { S t = A[i][k]; A[i][k] = A[i][idx]; A[i][idx] = t; }
      { S t = A[k][k]; A[k][k] = A[idx][idx]; A[idx][idx] = t; }
      for (int i = k + 1; i < idx; ++i) { S t = A[i][k]; A[i][k] = A[idx][i]; A[idx][i] = t; }
    }
    if (k > 0) {
      for (int j = 0; j < k; ++j) temp[j] = A[j][j] * A[k][j];
      S s = 0;
      for (int j = 0; j < k; ++j) s += A[k][j] * temp[j];
      A[k][k] -= s;
      for (int i = k + 1; i < 3; ++i) {
        S t = 0;
        for (int j = 0; j < k; ++j) t += A[i][j] * temp[j];
        A[i][k] -= t;
      }
    }
    const S akk = A[k][k];
    if (tabs(akk) > S(0))
      for (int i = k + 1; i < 3; ++i) A[i][k] /= akk;
  }
  x[0] = b[0]; x[1] = b[1]; x[2] = b[2];
#pragma unroll
  for (int k = 0; k < 3; ++k) { S t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
  x[1] -= A[1][0] * x[0];
  x[2] -= A[2][0] * x[0] + A[2][1] * x[1];
#pragma unroll
  for (int i = 0; i < 3; ++i) x[i] = (tabs(A[i][i]) > S(0)) ? x[i] / A[i][i] : S(0);
  x[1] -= A[2][1] * x[2];
  x[0] -= A[1][0] * x[1] + A[2][0] * x[2];
#pragma unroll
  for (int k = 2; k >= 0; --k) { S t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
}

// sum over the track's observations of the reprojection cost (msckf.h:1027-1047), warp-reduced
template <class S>
__device__ __forceinline__ S track_cost(const S* rel, const S* z, int L, int lane, const S x[3]) {
  S acc = 0;
  for (int i = lane; i < L; i += 32) {
    const S* T = rel + 12 * i;
    const S h0 = T[0] * x[0] + T[1] * x[1] + T[2] + x[2] * T[9];
    const S h1 = T[3] * x[0] + T[4] * x[1] + T[5] + x[2] * T[10];
    const S h2 = T[6] * x[0] + T[7] * x[1] + T[8] + x[2] * T[11];
    const S e0 = h0 / h2 - z[2 * i], e1 = h1 / h2 - z[2 * i + 1];
    acc += e0 * e0 + e1 * e1;
  }
  return warp_sum(acc);
}

template <class S, int WPB>
__global__ void __launch_bounds__(WPB * 32) k_tri(FeatArgs<S> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  S* poses = reinterpret_cast<S*>(smem_raw + 16);
  S* rel_all = poses + (size_t)a.M * kPoseStride;
  stage_table_tma(poses, a.poses, (unsigned)(a.M * kPoseStride * sizeof(S)), bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * WPB + warp;
  if (t >= a.n_tracks) return;
  const int o0 = a.obs_off[t], L = a.obs_off[t + 1] - o0;
  const int* idx = a.clone_idx + o0;
  const S* z = a.obs + 2 * (size_t)o0;
  S* rel = rel_all + (size_t)warp * a.Lmax * 12;
  // first clone: camera -> world
  const S* pose0 = poses + kPoseStride * idx[0];
  S C0[9];
  quat_to_rot(pose0, C0);
  const S p0[3] = {pose0[4], pose0[5], pose0[6]};
  // ---- checkMotion (msckf.h:980-1025)
  int cm = 0;
  if (L >= 2) {
    S d[3] = {z[0], z[1], S(1.0)};
    const S dn = tsqrt<S>(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] /= dn; d[1] /= dn; d[2] /= dn;
    const S dw[3] = {C0[0] * d[0] + C0[3] * d[1] + C0[6] * d[2], C0[1] * d[0] + C0[4] * d[1] + C0[7] * d[2],
                     C0[2] * d[0] + C0[5] * d[1] + C0[8] * d[2]};  // C0^T d
    S mo = 0;
    for (int i = 1 + lane; i < L; i += 32) {
      const S* ps = poses + kPoseStride * idx[i];
      const S tr[3] = {ps[4] - p0[0], ps[5] - p0[1], ps[6] - p0[2]};
      const S par = tr[0] * dw[0] + tr[1] * dw[1] + tr[2] * dw[2];
      const S o[3] = {tr[0] - par * dw[0], tr[1] - par * dw[1], tr[2] - par * dw[2]};
      const S nn = tsqrt<S>(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
      mo = nn > mo ? nn : mo;
    }
    mo = warp_max(mo);
    cm = mo > a.st->translation_threshold;
  }
  // ---- relative poses T_i^-1 * T_0 (msckf.h:1154-1168): rel = [R (9) | t (3)], first-clone frame -> clone i
  for (int i = lane; i < L; i += 32) {
    const S* ps = poses + kPoseStride * idx[i];
    S Ci[9];
    quat_to_rot(ps, Ci);
    S* T = rel + 12 * i;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)  // C_i * C_0^T
        T[3 * r + c] = Ci[3 * r] * C0[3 * c] + Ci[3 * r + 1] * C0[3 * c + 1] + Ci[3 * r + 2] * C0[3 * c + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const S a0 = Ci[3 * r] * p0[0] + Ci[3 * r + 1] * p0[1] + Ci[3 * r + 2] * p0[2];
      const S a1 = -(Ci[3 * r] * ps[4] + Ci[3 * r + 1] * ps[5] + Ci[3 * r + 2] * ps[6]);
      T[9 + r] = a0 + a1;
    }
  }
  __syncwarp();
  // ---- generateInitialGuess (msckf.h:1126-1145) from the first and the last observation
  S sol[3];
  {
    const S* T = rel + 12 * (L - 1);
    const S z1u = z[0], z1v = z[1], z2u = z[2 * (L - 1)], z2v = z[2 * (L - 1) + 1];
    const S m0 = T[0] * z1u + T[1] * z1v + T[2], m1 = T[3] * z1u + T[4] * z1v + T[5], m2 = T[6] * z1u + T[7] * z1v + T[8];
    const S A0 = m0 - z2u * m2, A1 = m1 - z2v * m2;
    const S b0 = z2u * T[11] - T[9], b1 = z2v * T[11] - T[10];
    const S depth = (S(1) / (A0 * A0 + A1 * A1)) * (A0 * b0 + A1 * b1);
    const S i0 = z1u * depth, i1 = z1v * depth, i2 = depth;
    sol[0] = i0 / i2; sol[1] = i1 / i2; sol[2] = S(1.0) / i2;
  }
  // ---- Levenberg-Marquardt (msckf.h:1178-1248); all control flow is warp-uniform
  S lambda = S(1e-3);
  int inner = 0, outer = 0;
  bool reduced = false;
  S delta_norm = 0;
  S total_cost = track_cost(rel, z, L, lane, sol);
  do {
    S s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // A00 A01 A02 A11 A12 A22 b0 b1 b2
    for (int i = lane; i < L; i += 32) {
      const S* T = rel + 12 * i;
      const S h0 = T[0] * sol[0] + T[1] * sol[1] + T[2] + sol[2] * T[9];
      const S h1 = T[3] * sol[0] + T[4] * sol[1] + T[5] + sol[2] * T[10];
      const S h2 = T[6] * sol[0] + T[7] * sol[1] + T[8] + sol[2] * T[11];
      // W = [R(:,0:2) | t]
      const S W0[3] = {T[0], T[1], T[9]}, W1[3] = {T[3], T[4], T[10]}, W2[3] = {T[6], T[7], T[11]};
      S J0[3], J1[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        J0[c] = S(1) / h2 * W0[c] - h0 / (h2 * h2) * W2[c];
        J1[c] = S(1) / h2 * W1[c] - h1 / (h2 * h2) * W2[c];
      }
      const S r0 = h0 / h2 - z[2 * i], r1 = h1 / h2 - z[2 * i + 1];
      const S e = tsqrt<S>(r0 * r0 + r1 * r1);
      const S w = (e <= S(0.01)) ? S(1.0) : S(0.01) / (S(2) * e);
      const S w2 = (w == S(1)) ? S(1) : w * w;
      s[0] += w2 * (J0[0] * J0[0] + J1[0] * J1[0]);
      s[1] += w2 * (J0[0] * J0[1] + J1[0] * J1[1]);
      s[2] += w2 * (J0[0] * J0[2] + J1[0] * J1[2]);
      s[3] += w2 * (J0[1] * J0[1] + J1[1] * J1[1]);
      s[4] += w2 * (J0[1] * J0[2] + J1[1] * J1[2]);
      s[5] += w2 * (J0[2] * J0[2] + J1[2] * J1[2]);
      s[6] += w2 * (J0[0] * r0 + J1[0] * r1);
      s[7] += w2 * (J0[1] * r0 + J1[1] * r1);
      s[8] += w2 * (J0[2] * r0 + J1[2] * r1);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = warp_sum(s[k]);
    do {
      S A[3][3] = {{s[0] + lambda, s[1], s[2]}, {s[1], s[3] + lambda, s[4]}, {s[2], s[4], s[5] + lambda}};
      const S b[3] = {s[6], s[7], s[8]};
      S delta[3];
      ldlt3_solve(A, b, delta);
      const S ns[3] = {sol[0] - delta[0], sol[1] - delta[1], sol[2] - delta[2]};
      delta_norm = tsqrt<S>(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
      const S new_cost = track_cost(rel, z, L, lane, ns);
      if (new_cost < total_cost) {
        reduced = true;
        sol[0] = ns[0]; sol[1] = ns[1]; sol[2] = ns[2];
        total_cost = new_cost;
        lambda = (S)((double)(lambda / 10) > 1e-10 ? (double)(lambda / 10) : 1e-10);
      } else {
        reduced = false;
        lambda = (S)((double)(lambda * 10) < 1e12 ? (double)(lambda * 10) : 1e12);
      }
    } while (inner++ < 10 && !reduced);
    inner = 0;
  } while (outer++ < 10 && delta_norm > S(5e-7));
  const S fin[3] = {sol[0] / sol[2], sol[1] / sol[2], S(1.0) / sol[2]};
  // ---- validity (msckf.h:1255-1276)
  int bad = 0;
  for (int i = lane; i < L; i += 32) {
    const S* T = rel + 12 * i;
    const S pz = T[6] * fin[0] + T[7] * fin[1] + T[8] * fin[2] + T[11];
    if (pz <= S(0)) bad = 1;
  }
  bad = __any_sync(0xffffffffu, bad);
  const S normalized_cost = total_cost / (S)(2 * (size_t)L * (size_t)L);
  int ok = !bad;
  if (normalized_cost > a.st->max_gn_cost_norm) ok = 0;
  if (lane == 0) {
    // p_f_G = C0^T * fin + p0 (msckf.h:1282), computed even when invalid
    a.pfg[3 * t + 0] = C0[0] * fin[0] + C0[3] * fin[1] + C0[6] * fin[2] + p0[0];
    a.pfg[3 * t + 1] = C0[1] * fin[0] + C0[4] * fin[1] + C0[7] * fin[2] + p0[1];
    a.pfg[3 * t + 2] = C0[2] * fin[0] + C0[5] * fin[1] + C0[8] * fin[2] + p0[2];
    a.cm_ok[t] = cm;
    a.tri_ok[t] = ok;
  }
}

// Loop A of marginalize() (msckf.h:352-399) as a bookkeeping pass over the per-track flags.
// Single CTA.  mode 0: marginalize semantics; mode 1: every track valid at its own/given position.
template <class S>
__global__ void k_resolve(FeatArgs<S> a, DevState<S>* st, int mode, int* pushed_scratch) {
  __shared__ int s_head_end, s_total_pushed;
  __shared__ unsigned long long s_counter;
  __shared__ int s_scan[1024];
  const int N = a.n_tracks;
  const int tid = threadIdx.x;
  if (mode == 1) {
    for (int k = tid; k < N; k += blockDim.x) { a.valid[k] = 1; a.src[k] = k; }
    return;
  }
  // head: walk sequentially while the residualised-track counter is <= 3 (checkMotion not yet applied)
  if (tid == 0) {
    unsigned long long counter = st->num_residualized;
    int k = 0;
    for (; k < N && counter <= 3; ++k) {
      pushed_scratch[k] = 1;  // checkMotion not consulted -> initializePosition ran -> p_f_G pushed
      a.cm_ok[k] = 1;         // report: "not rejected by checkMotion" (msckf.h:354)
      const int v = a.tri_ok[k];
      a.valid[k] = v;
      if (v) counter++;
    }
    s_head_end = k;
    s_counter = counter;
  }
  __syncthreads();
  const int he = s_head_end;
  for (int k = he + tid; k < N; k += blockDim.x) {
    const int cm = a.cm_ok[k];
    pushed_scratch[k] = cm;
    a.valid[k] = cm && a.tri_ok[k];
  }
  __syncthreads();
  // exclusive scan of pushed[] (chunked, blockDim.x == 1024) -> position in p_f_G_vec; inverse map pi[]
  int carry = 0;
  for (int base = 0; base < N; base += 1024) {
    const int k = base + tid;
    const int v = (k < N) ? pushed_scratch[k] : 0;
    s_scan[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = (tid >= o) ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += t;
      __syncthreads();
    }
    const int incl = s_scan[tid];
    if (k < N && v) a.src[carry + incl - 1] = k;  // pi[rank] = k  (temporarily stored in src[])
    const int tot = s_scan[1023];
    __syncthreads();
    carry += tot;
  }
  if (tid == 0) s_total_pushed = carry;
  __syncthreads();
  const int total = s_total_pushed;
  // src[] currently holds pi[0..total); loop B reads p_f_G_vec[iter] = pfg[pi[iter]] when iter < total,
  // else (out of bounds in the reference, UB) the track's own position.  Resolve in place via scratch.
  __syncthreads();
  for (int k = tid; k < N; k += blockDim.x) pushed_scratch[k] = (k < total) ? a.src[k] : k;
  __syncthreads();
  unsigned long long shifted = 0, oob = 0, nvalid = 0;
  for (int k = tid; k < N; k += blockDim.x) {
    const int sidx = pushed_scratch[k];
    a.src[k] = sidx;
    if (a.valid[k]) {
      if (k >= he) nvalid++;
      if (k < total) { if (sidx != k) shifted++; } else oob++;
    }
  }
  // block-reduce the three counters
  __shared__ unsigned long long s_red[3];
  if (tid == 0) { s_red[0] = s_red[1] = s_red[2] = 0; }
  __syncthreads();
  if (shifted) atomicAdd(&s_red[0], shifted);
  if (oob) atomicAdd(&s_red[1], oob);
  if (nvalid) atomicAdd(&s_red[2], nvalid);
  __syncthreads();
  if (tid == 0) {
    st->pfg_shifted += s_red[0];
    st->pfg_oob += s_red[1];
    st->num_residualized = s_counter + s_red[2];
  }
}

// packed lower-triangular symmetric storage
__device__ __forceinline__ int pk(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

template <class S>
__host__ __device__ inline size_t jac_warp_smem_elems(int L) {
  // X 12L | r 2L | V 6L | U64 6L doubles (= 12L floats worst case) | w 2L | p 2L | Ypacked L(2L+1)
  return (size_t)36 * L + (size_t)L * (2 * L + 1) + 4;
}

template <class S, int WPB>
__global__ void __launch_bounds__(WPB * 32) k_jac(FeatArgs<S> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  S* poses = reinterpret_cast<S*>(smem_raw + 16);
  S* ws_all = poses + (size_t)a.M * kPoseStride;
  stage_table_tma(poses, a.poses, (unsigned)(a.M * kPoseStride * sizeof(S)), bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * WPB + warp;
  if (t >= a.n_tracks) return;
  const int c = 6 * a.M;
  const int o0 = a.obs_off[t], L = a.obs_off[t + 1] - o0, L2 = 2 * L;
  double* Zr = a.Z + (size_t)3 * t * c;
  double* Yr = a.Yq + (size_t)3 * t * c;
  if (!a.valid[t]) {  // not residualised: contributes nothing
    for (int k = lane; k < 3 * c; k += 32) { Zr[k] = 0.0; Yr[k] = 0.0; }
    if (lane == 0) { a.accept[t] = 0; a.gamma[t] = S(0); a.rows[t] = 0; a.ur[3 * t] = a.ur[3 * t + 1] = a.ur[3 * t + 2] = 0.0; }
    return;
  }
  const int* idx = a.clone_idx + o0;
  const S* z = a.obs + 2 * (size_t)o0;
  S* wbase = ws_all + (size_t)warp * jac_warp_smem_elems<S>(a.Lmax);
  // U in fp64 (8-byte aligned: the per-warp region starts 16-byte aligned and its size is even in S units)
  double* U = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(wbase) + 7) & ~uintptr_t(7));
  S* X = reinterpret_cast<S*>(U + 3 * L2);
  S* r = X + 12 * L;
  S* V = r + L2;
  S* wv = V + 3 * L2;
  S* pv = wv + L2;
  S* Y = pv + L2;
  const DevState<S>* st = a.st;
  const S g[3] = {st->g[0], st->g[1], st->g[2]};
  const S* pfsrc = a.pfg_given ? (a.pfg_given + 3 * t) : (a.pfg + 3 * a.src[t]);
  const S pf[3] = {pfsrc[0], pfsrc[1], pfsrc[2]};
  // ---- residual + measurement Jacobian blocks with the observability projection (msckf.h:915-950, 960-978)
  for (int i = lane; i < L; i += 32) {
    const S* ps = poses + kPoseStride * idx[i];
    S C[9];
    quat_to_rot(ps, C);
    const S d[3] = {pf[0] - ps[4], pf[1] - ps[5], pf[2] - ps[6]};
    const S pc[3] = {C[0] * d[0] + C[1] * d[1] + C[2] * d[2], C[3] * d[0] + C[4] * d[1] + C[5] * d[2],
                     C[6] * d[0] + C[7] * d[1] + C[8] * d[2]};
    const S Xc = pc[0], Yc = pc[1], Zc = pc[2];
    const S r0 = z[2 * i] - Xc / Zc, r1 = z[2 * i + 1] - Yc / Zc;
    r[2 * i] = r0; r[2 * i + 1] = r1;
    a.rg[2 * (size_t)(o0 + i)] = r0; a.rg[2 * (size_t)(o0 + i) + 1] = r1;
    const S iz = S(1) / Zc;
    const S J[2][3] = {{S(1) * iz, S(0) * iz, (-Xc / Zc) * iz}, {S(0) * iz, S(1) * iz, (-Yc / Zc) * iz}};
    // skew(pc)
    const S sk[3][3] = {{S(0), -pc[2], pc[1]}, {pc[2], S(0), -pc[0]}, {-pc[1], pc[0], S(0)}};
    S A[2][6];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        A[q][b] = J[q][0] * sk[0][b] + J[q][1] * sk[1][b] + J[q][2] * sk[2][b];
        A[q][3 + b] = -(J[q][0] * C[b] + J[q][1] * C[3 + b] + J[q][2] * C[6 + b]);
      }
    S u[6];
    u[0] = C[0] * g[0] + C[1] * g[1] + C[2] * g[2];
    u[1] = C[3] * g[0] + C[4] * g[1] + C[5] * g[2];
    u[2] = C[6] * g[0] + C[7] * g[1] + C[8] * g[2];
    u[3] = -d[2] * g[1] + d[1] * g[2];  // skew(d) * g
    u[4] = d[2] * g[0] - d[0] * g[2];
    u[5] = -d[1] * g[0] + d[0] * g[1];
    S uu = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) uu += u[b] * u[b];
    const S iuu = S(1) / uu;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      S Au = 0;
#pragma unroll
      for (int b = 0; b < 6; ++b) Au += A[q][b] * u[b];
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        const S hx = A[q][b] - Au * iuu * u[b];
        X[12 * i + 6 * q + b] = hx;
        a.Xg[12 * (size_t)(o0 + i) + 6 * q + b] = hx;
        if (b >= 3) V[3 * (2 * i + q) + (b - 3)] = -hx;  // H_f = -H_x(:,3:6)
      }
    }
  }
  __syncwarp();
  // ---- column-pivoted Householder QR of H_f (2L x 3): the trailing 2L-3 columns of Q are A_j (msckf.h:954-955)
  S tau[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    S nn[3] = {S(-1), S(-1), S(-1)};
    for (int cc = k; cc < 3; ++cc) {
      S sacc = 0;
      for (int row = k + lane; row < L2; row += 32) sacc += V[3 * row + cc] * V[3 * row + cc];
      nn[cc] = warp_sum(sacc);
    }
    int piv = k;
    S best = nn[k];
    for (int cc = k + 1; cc < 3; ++cc)
      if (nn[cc] > best) { best = nn[cc]; piv = cc; }
    if (piv != k)
      for (int row = lane; row < L2; row += 32) { const S tmp = V[3 * row + k]; V[3 * row + k] = V[3 * row + piv]; V[3 * row + piv] = tmp; }
    __syncwarp();
    S tacc = 0;
    for (int row = k + 1 + lane; row < L2; row += 32) tacc += V[3 * row + k] * V[3 * row + k];
    const S tail = warp_sum(tacc);
    const S c0 = V[3 * k + k];
    S beta, tk;
    if (tail <= S(sizeof(S) == 4 ? 1.17549435e-38 : 2.2250738585072014e-308)) {
      beta = c0; tk = S(0);
      for (int row = k + 1 + lane; row < L2; row += 32) V[3 * row + k] = S(0);
    } else {
      beta = tsqrt<S>(c0 * c0 + tail);
      if (c0 >= S(0)) beta = -beta;
      const S dd = c0 - beta;
      for (int row = k + 1 + lane; row < L2; row += 32) V[3 * row + k] /= dd;
      tk = (beta - c0) / beta;
    }
    tau[k] = tk;
    __syncwarp();
    if (tk != S(0)) {
      for (int cc = k + 1; cc < 3; ++cc) {
        S sacc = 0;
        for (int row = k + 1 + lane; row < L2; row += 32) sacc += V[3 * row + k] * V[3 * row + cc];
        S sdot = warp_sum(sacc) + V[3 * k + cc];
        sdot *= tk;
        for (int row = k + 1 + lane; row < L2; row += 32) V[3 * row + cc] -= V[3 * row + k] * sdot;
        __syncwarp();
        if (lane == 0) V[3 * k + cc] -= sdot;
        __syncwarp();
      }
    }
    if (lane == 0) V[3 * k + k] = S(1);  // unit diagonal of the Householder vector
    // zero above the diagonal of column k so V(:,k) is the full Householder vector v_k
    if (lane == 0) for (int row = 0; row < k; ++row) V[3 * row + k] = S(0);
    __syncwarp();
  }
  // export v_k, tau for the head-row kernel
  for (int e = lane; e < 3 * L2; e += 32) a.Vg[3 * 2 * (size_t)o0 + e] = V[e];
  if (lane < 3) a.taug[3 * t + lane] = tau[lane];
  // ---- exact reflectors for the Gram stage: H_k = I - tau64_k v_k v_k^T with tau64_k = 2 / (v_k^T v_k) evaluated in
  // fp64 from the stored vectors.  Q = H_0 H_1 H_2 is then orthogonal to 1e-16 whatever the filter precision, so
  // G_j = I - U_j U_j^T is an exact projector and the body Gram terms and the head rows (k_head) describe the same H_o.
  double tau64[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double sacc = 0.0;
    for (int row = k + lane; row < L2; row += 32) { const double vv = (double)V[3 * row + k]; sacc += vv * vv; }
    const double vtv = warp_sum(sacc);
    tau64[k] = (tau[k] != S(0)) ? 2.0 / vtv : 0.0;
  }
  // ---- U = Q(:,0:3) = H0 H1 H2 [I3; 0]  (fp64)
  for (int e = lane; e < 3 * L2; e += 32) U[e] = ((e / 3) == (e % 3)) ? 1.0 : 0.0;
  __syncwarp();
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    for (int cc = 0; cc < 3; ++cc) {
      double sacc = 0.0;
      for (int row = k + lane; row < L2; row += 32) sacc += (double)V[3 * row + k] * U[3 * row + cc];
      const double sdot = tau64[k] * warp_sum(sacc);
      for (int row = k + lane; row < L2; row += 32) U[3 * row + cc] -= sdot * (double)V[3 * row + k];
    }
    __syncwarp();
  }
  // U^T r in fp64 from the raw residual
  double urv[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double sacc = 0.0;
    for (int row = lane; row < L2; row += 32) sacc += U[3 * row + q] * (double)r[row];
    urv[q] = warp_sum(sacc);
  }
  // ---- r~ = H2 H1 H0 r in the filter precision (r_o = r~[3:], used by the gate only)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    S sacc = 0;
    for (int row = k + lane; row < L2; row += 32) sacc += V[3 * row + k] * r[row];
    const S sdot = tau[k] * warp_sum(sacc);
    for (int row = k + lane; row < L2; row += 32) r[row] -= sdot * V[3 * row + k];
    __syncwarp();
  }
  // ---- gating (msckf.h:1103-1124): gamma = r_o^T (H_o P H_o^T + u_var I)^-1 r_o with H_o = (Q^T X)[3:]
  // Y = X P_sub X^T, symmetric 2L x 2L, packed lower
  {
    const int npairs = L * (L + 1) / 2;
    const S* P = a.P;
    const int ldp = a.ldp;
    for (int p = lane; p < npairs; p += 32) {
      int i = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
      while (i * (i + 1) / 2 > p) --i;
      while ((i + 1) * (i + 2) / 2 <= p) ++i;
      const int k = p - i * (i + 1) / 2;  // k <= i
      const S* Pb = P + (size_t)(kImuDim + 6 * idx[i]) * ldp + (kImuDim + 6 * idx[k]);
      const S* Xi = X + 12 * i;
      const S* Xk = X + 12 * k;
      S T0[6], T1[6];
#pragma unroll
      for (int b = 0; b < 6; ++b) { T0[b] = S(0); T1[b] = S(0); }
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        const S x0 = Xi[q], x1 = Xi[6 + q];
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          const S pv_ = Pb[(size_t)q * ldp + b];
          T0[b] += x0 * pv_;
          T1[b] += x1 * pv_;
        }
      }
      S y00 = 0, y01 = 0, y10 = 0, y11 = 0;
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        y00 += T0[b] * Xk[b]; y01 += T0[b] * Xk[6 + b];
        y10 += T1[b] * Xk[b]; y11 += T1[b] * Xk[6 + b];
      }
      Y[pk(2 * i, 2 * k)] = y00;
      Y[pk(2 * i + 1, 2 * k)] = y10;
      Y[pk(2 * i + 1, 2 * k + 1)] = y11;
      if (i != k) Y[pk(2 * i, 2 * k + 1)] = y01;
    }
  }
  __syncwarp();
  // two-sided reflectors: Y <- H_k Y H_k on the trailing (>= k) block
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const S tk = tau[k];
    if (tk != S(0)) {
      for (int ar = k + lane; ar < L2; ar += 32) {
        S sacc = 0;
        for (int b = k; b <= ar; ++b) sacc += Y[pk(ar, b)] * V[3 * b + k];
        for (int b = ar + 1; b < L2; ++b) sacc += Y[pk(b, ar)] * V[3 * b + k];
        wv[ar] = sacc;
      }
      __syncwarp();
      S aacc = 0;
      for (int ar = k + lane; ar < L2; ar += 32) aacc += V[3 * ar + k] * wv[ar];
      const S alpha = warp_sum(aacc);
      const S hc = tk * tk * alpha * S(0.5);
      for (int ar = k + lane; ar < L2; ar += 32) pv[ar] = tk * wv[ar] - hc * V[3 * ar + k];
      __syncwarp();
      for (int ar = k + lane; ar < L2; ar += 32) {
        const S va = V[3 * ar + k], pa = pv[ar];
        for (int b = k; b <= ar; ++b) Y[pk(ar, b)] -= va * pv[b] + pa * V[3 * b + k];
      }
      __syncwarp();
    }
  }
  // S = Y[3:,3:] + u_var I ; Cholesky with the right-hand side r~[3:] carried as an extra row
  const S uvar = st->u_var;
  for (int j = 3 + lane; j < L2; j += 32) Y[pk(j, j)] += uvar;
  __syncwarp();
  bool chol_ok = true;
  for (int j = 3; j < L2; ++j) {
    const S dj = Y[pk(j, j)];
    if (!(dj > S(0))) { chol_ok = false; break; }
    const S ljj = tsqrt<S>(dj);
    __syncwarp();
    for (int i = j + 1 + lane; i < L2; i += 32) Y[pk(i, j)] /= ljj;
    if (lane == 0) { r[j] /= ljj; Y[pk(j, j)] = ljj; }
    __syncwarp();
    const S ej = r[j];
    for (int i = j + 1 + lane; i < L2; i += 32) {
      const S lij = Y[pk(i, j)];
      r[i] -= ej * lij;
      for (int cc = j + 1; cc <= i; ++cc) Y[pk(i, cc)] -= lij * Y[pk(cc, j)];
    }
    __syncwarp();
  }
  S gacc = 0;
  for (int j = 3 + lane; j < L2; j += 32) gacc += r[j] * r[j];
  S gam = warp_sum(gacc);
  const int acc = chol_ok && (gam < st->chi2[L]);  // table[dof+1], dof = L-1 (msckf.h:433,:1117)
  if (!chol_ok) gam = S(1e30);
  // ---- compact outputs for the Gram stage
  for (int k = lane; k < 3 * c; k += 32) { Zr[k] = 0.0; Yr[k] = 0.0; }
  // M = U^T D U (3x3 symmetric), D = diag(u_var, v_var, u_var, ...)
  double Mm[6] = {0, 0, 0, 0, 0, 0};
  const double du = (double)st->u_var, dv = (double)st->v_var;
  for (int row = lane; row < L2; row += 32) {
    const double dd = (row & 1) ? dv : du;
    const double u0 = U[3 * row], u1 = U[3 * row + 1], u2 = U[3 * row + 2];
    Mm[0] += dd * u0 * u0; Mm[1] += dd * u0 * u1; Mm[2] += dd * u0 * u2;
    Mm[3] += dd * u1 * u1; Mm[4] += dd * u1 * u2; Mm[5] += dd * u2 * u2;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) Mm[k] = warp_sum(Mm[k]);
  const double M3[3][3] = {{Mm[0], Mm[1], Mm[2]}, {Mm[1], Mm[3], Mm[4]}, {Mm[2], Mm[4], Mm[5]}};
  __syncwarp();
  if (acc) {
    for (int i = lane; i < L; i += 32) {
      const int col0 = 6 * idx[i];
      double zb[3][6], yb[3][6];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double ua = U[3 * (2 * i) + q], ub = U[3 * (2 * i + 1) + q];
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          const double xa = X[12 * i + b], xb = X[12 * i + 6 + b];
          zb[q][b] = ua * xa + ub * xb;
          yb[q][b] = du * ua * xa + dv * ub * xb;
        }
      }
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int b = 0; b < 6; ++b) {
          Zr[(size_t)q * c + col0 + b] = zb[q][b];
          Yr[(size_t)q * c + col0 + b] = yb[q][b] - 0.5 * (M3[q][0] * zb[0][b] + M3[q][1] * zb[1][b] + M3[q][2] * zb[2][b]);
        }
    }
  }
  if (lane == 0) {
    a.accept[t] = acc;
    a.gamma[t] = gam;
    a.rows[t] = acc ? (L2 - 3) : 0;
    for (int q = 0; q < 3; ++q) a.ur[3 * t + q] = acc ? urv[q] : 0.0;
  }
}

// Ordered stacking (msckf.h:433-445): exclusive prefix sum of the accepted blocks' row counts.
__global__ void k_scan(int N, const int* rows, int* row_off /*[N+1]*/, int* m_out) {
  __shared__ int s_scan[1024];
  const int tid = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < N; base += 1024) {
    const int k = base + tid;
    const int v = (k < N) ? rows[k] : 0;
    s_scan[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int t = (tid >= o) ? s_scan[tid - o] : 0;
      __syncthreads();
      s_scan[tid] += t;
      __syncthreads();
    }
    if (k < N) row_off[k] = carry + s_scan[tid] - v;
    const int tot = s_scan[1023];
    __syncthreads();
    carry += tot;
  }
  if (tid == 0) { row_off[N] = carry; *m_out = carry; }
}

}  // namespace mb
