// msckf_mono_b200/csrc/tail_kernels.cuh
// Dense n x n tail of the EKF update (msckf.h:1369-1418) in square-root form, fp64 throughout:
//   TP  = T'' P                      S'' = TP T''^T + R''          (k_gemm_tp, k_gemm_s)
//   S'' = L L^T restricted to the independent rows (rank-revealing skip-Cholesky, k_chol)
//   W   = L^-1 [TP | r'']            (k_trsm)
//   P+  = P - W^T W  (== Joseph form for the optimal gain of the projected system),  dx = W^T (L^-1 r'')
//   state injection (msckf.h:1373-1391, buildUpdateQuat :851-872)
#pragma once
#include "common.cuh"

namespace mb {

// TP[a][b] = sum_k T2[a][k] * P[k][b]      (T2 columns < 15 are zero)
template <class S>
__global__ void __launch_bounds__(256) k_gemm_tp(int n, int ld, const double* __restrict__ T2, const S* __restrict__ P, int ldp,
                                                double* __restrict__ TP) {
  __shared__ double sA[32][17], sB[16][33];
  const int ta = blockIdx.y * 32, tb = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int kb = kImuDim; kb < n; kb += 16) {
    for (int e = threadIdx.x; e < 32 * 16; e += 256) {
      const int r = e / 16, kk = e % 16;
      const int a = ta + r, k = kb + kk;
      sA[r][kk] = (a < n && k < n) ? T2[(size_t)a * ld + k] : 0.0;
    }
    for (int e = threadIdx.x; e < 16 * 32; e += 256) {
      const int kk = e / 32, cc = e % 32;
      const int k = kb + kk, b = tb + cc;
      sB[kk][cc] = (k < n && b < n) ? (double)P[(size_t)k * ldp + b] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const double a0 = sA[ty][kk], a1 = sA[ty + 16][kk], b0 = sB[kk][tx], b1 = sB[kk][tx + 16];
      acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = ta + ty + 16 * i, b = tb + tx + 16 * j;
      if (a < n && b < n) TP[(size_t)a * ld + b] = acc[i][j];
    }
}

// S2[a][b] = sum_k TP[a][k] * T2[b][k] + R2[a][b]
__global__ void __launch_bounds__(256) k_gemm_s(int n, int ld, const double* __restrict__ TP, const double* __restrict__ T2,
                                               const double* __restrict__ R2, double* __restrict__ S2) {
  __shared__ double sA[32][17], sB[32][17];
  const int ta = blockIdx.y * 32, tb = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int kb = kImuDim; kb < n; kb += 16) {
    for (int e = threadIdx.x; e < 32 * 16; e += 256) {
      const int r = e / 16, kk = e % 16;
      const int k = kb + kk;
      sA[r][kk] = (ta + r < n && k < n) ? TP[(size_t)(ta + r) * ld + k] : 0.0;
      sB[r][kk] = (tb + r < n && k < n) ? T2[(size_t)(tb + r) * ld + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const double a0 = sA[ty][kk], a1 = sA[ty + 16][kk], b0 = sB[tx][kk], b1 = sB[tx + 16][kk];
      acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = ta + ty + 16 * i, b = tb + tx + 16 * j;
      if (a < n && b < n) S2[(size_t)a * ld + b] = acc[i][j] + R2[(size_t)a * ld + b];
    }
}

// Rank-revealing Cholesky in natural order ("skip" variant): index k is dropped when its pivot falls below
// thr * (original diagonal) -- i.e. row k of Q'' is (numerically) in the span of the previous ones.  Dropped
// rows/columns are removed from the factor (L_kk = 1, rest 0).  Single CTA, blocked (32), lower triangle in place.
// decide = 1: rank decisions are taken here (pivot <= thr * original diagonal, or rank cap reached) and written to keep[].
// decide = 0: keep[] is given (decided on the Gram matrix of the basis, k_gamma); only non-positive pivots are dropped.
__global__ void __launch_bounds__(1024) k_chol(int n, int ld, double* __restrict__ A, int* __restrict__ keep, double thr,
                                              int* __restrict__ rank_out, const int* __restrict__ m_in, int decide) {
  extern __shared__ double sm[];
  double* d0 = sm;                       // [n]
  double* D = d0 + ((n + 1) & ~1);       // [32][33]
  double* Lp = D + 32 * 33;              // [(n) x 33] panel rows below the diagonal block
  int* skeep = reinterpret_cast<int*>(Lp + (size_t)n * 33);  // [n]
  __shared__ int s_rank;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the stacked system has m rows: span(Q'') cannot have more than min(m, n) dimensions
  const int rank_cap = min(*m_in, n);
  for (int k = tid; k < n; k += 1024) { d0[k] = A[(size_t)k * ld + k]; skeep[k] = decide ? 1 : keep[k]; }
  if (tid == 0) s_rank = 0;
  __syncthreads();
  for (int kb = 0; kb < n; kb += 32) {
    const int nb = min(32, n - kb);
    for (int e = tid; e < nb * nb; e += 1024) {
      const int i = e / nb, j = e % nb;
      D[i * 33 + j] = (j <= i) ? A[(size_t)(kb + i) * ld + kb + j] : 0.0;
    }
    __syncthreads();
    if (warp == 0) {
      int rank_now = s_rank;
      for (int j = 0; j < nb; ++j) {
        const double piv = D[j * 33 + j];
        const double dj0 = d0[kb + j];
        const bool drop = decide ? (!(dj0 > 0.0) || !(piv > thr * dj0) || rank_now >= rank_cap)
                                 : (skeep[kb + j] == 0 || !(piv > 0.0));
        if (!drop) rank_now++;
        __syncwarp();
        if (drop) {
          if (lane == 0) { skeep[kb + j] = 0; D[j * 33 + j] = 1.0; }
          if (lane > j && lane < nb) D[lane * 33 + j] = 0.0;
          if (lane < j) D[j * 33 + lane] = 0.0;
        } else {
          const double ljj = sqrt(piv);
          if (lane == 0) { skeep[kb + j] = 1; D[j * 33 + j] = ljj; }
          if (lane > j && lane < nb) D[lane * 33 + j] /= ljj;
          __syncwarp();
          if (lane > j && lane < nb) {
            const double lij = D[lane * 33 + j];
            for (int cc = j + 1; cc <= lane; ++cc) D[lane * 33 + cc] -= lij * D[cc * 33 + j];
          }
        }
        __syncwarp();
      }
      if (lane == 0) s_rank = rank_now;
    }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += 1024) {
      const int i = e / nb, j = e % nb;
      if (j <= i) A[(size_t)(kb + i) * ld + kb + j] = D[i * 33 + j];
    }
    // panel rows below: x D^T = a
    const int r0 = kb + nb;
    for (int row = r0 + tid; row < n; row += 1024) {
      double x[32];
#pragma unroll 1
      for (int j = 0; j < nb; ++j) {
        double v = A[(size_t)row * ld + kb + j];
        for (int cc = 0; cc < j; ++cc) v -= x[cc] * D[j * 33 + cc];
        const bool kp = skeep[kb + j] != 0;  // written by warp 0 before the barrier above
        x[j] = kp ? v / D[j * 33 + j] : 0.0;
        A[(size_t)row * ld + kb + j] = x[j];
        Lp[(size_t)(row - r0) * 33 + j] = x[j];
      }
    }
    __syncthreads();
    // trailing update of the lower triangle: A[i][i2] -= sum_j Lp[i][j] Lp[i2][j], i >= i2 >= r0
    const int nr = n - r0;
    const int npair = nr * (nr + 1) / 2;
    for (int p = tid; p < npair; p += 1024) {
      int i = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
      while (i * (i + 1) / 2 > p) --i;
      while ((i + 1) * (i + 2) / 2 <= p) ++i;
      const int i2 = p - i * (i + 1) / 2;
      double s = 0.0;
#pragma unroll 8
      for (int j = 0; j < nb; ++j) s += Lp[(size_t)i * 33 + j] * Lp[(size_t)i2 * 33 + j];
      A[(size_t)(r0 + i) * ld + r0 + i2] -= s;
    }
    __syncthreads();
  }
  // dropped rows: clear the part of the row left of the diagonal that earlier panels produced
  for (int e = tid; e < n * n; e += 1024) {
    const int k = e / n, cc = e % n;
    if (cc < k && !skeep[k]) A[(size_t)k * ld + cc] = 0.0;
  }
  for (int k = tid; k < n; k += 1024) keep[k] = skeep[k];
  if (tid == 0) *rank_out = s_rank;
}

// Gram matrix of the basis Q'' = [E_h | H_c]:  Gamma = [[I_h, H_h], [H_h^T, Lambda]]  (all of it is already in T'').
// Its rank-revealing Cholesky decides which basis vectors are independent -- a purely geometric decision,
// unaffected by how large the prior covariance is relative to the measurement noise.
__global__ void __launch_bounds__(256) k_gamma(int n, int ld, const double* __restrict__ T2, const int* __restrict__ m_in,
                                              double* __restrict__ Gm) {
  const int m = *m_in;
  const bool full = m <= n;
  const int h = full ? m : kImuDim;
  const size_t total = (size_t)n * n;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int a = (int)(e / n), b = (int)(e % n);
    double v;
    if (full) v = (a == b && a < h) ? 1.0 : 0.0;  // all rows explicit and orthonormal: Gamma = I_m
    else if (a < kImuDim && b < kImuDim) v = (a == b && a < h) ? 1.0 : 0.0;
    else if (a < kImuDim) v = T2[(size_t)a * ld + b];
    else if (b < kImuDim) v = T2[(size_t)b * ld + a];
    else v = T2[(size_t)a * ld + b];
    Gm[(size_t)a * ld + b] = v;
  }
}

// W = L^-1 [TP | r''] by column slabs of 32 (column n is r'').  Dropped rows produce zeros.
__global__ void __launch_bounds__(256) k_trsm(int n, int ld, const double* __restrict__ Lm, const int* __restrict__ keep,
                                             const double* __restrict__ TP, const double* __restrict__ r2,
                                             double* __restrict__ Wm, double* __restrict__ yv) {
  extern __shared__ double sm[];
  double* Ws = sm;              // [n][33]
  double* D = Ws + (size_t)n * 33;  // [32][33]
  double* Lpan = D + 32 * 33;       // [n][33] panel of L below the diagonal block
  int* skeep = reinterpret_cast<int*>(Lpan + (size_t)n * 33);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col0 = blockIdx.x * 32;
  for (int k = tid; k < n; k += 256) skeep[k] = keep[k];
  __syncthreads();
  for (int e = tid; e < n * 32; e += 256) {
    const int row = e / 32, cc = e % 32, col = col0 + cc;
    double v = 0.0;
    if (skeep[row]) {
      if (col < n) v = TP[(size_t)row * ld + col];
      else if (col == n) v = r2[row];
    }
    Ws[(size_t)row * 33 + cc] = v;
  }
  __syncthreads();
  for (int kb = 0; kb < n; kb += 32) {
    const int nb = min(32, n - kb);
    const int r0 = kb + nb;
    for (int e = tid; e < nb * nb; e += 256) {
      const int i = e / nb, j = e % nb;
      D[i * 33 + j] = (j <= i) ? Lm[(size_t)(kb + i) * ld + kb + j] : 0.0;
    }
    for (int e = tid; e < (n - r0) * nb; e += 256) {  // coalesced panel load
      const int i = e / nb, j = e % nb;
      Lpan[(size_t)i * 33 + j] = Lm[(size_t)(r0 + i) * ld + kb + j];
    }
    __syncthreads();
    if (warp == 0) {
      for (int j = 0; j < nb; ++j) {
        const int row = kb + j;
        double x = Ws[(size_t)row * 33 + lane];
        for (int cc = 0; cc < j; ++cc) x -= D[j * 33 + cc] * Ws[(size_t)(kb + cc) * 33 + lane];
        Ws[(size_t)row * 33 + lane] = skeep[row] ? x / D[j * 33 + j] : 0.0;
      }
    }
    __syncthreads();
    for (int e = tid; e < (n - r0) * 32; e += 256) {
      const int i = e / 32, cc = e % 32;
      double s = 0.0;
#pragma unroll 8
      for (int j = 0; j < nb; ++j) s += Lpan[(size_t)i * 33 + j] * Ws[(size_t)(kb + j) * 33 + cc];
      Ws[(size_t)(r0 + i) * 33 + cc] -= s;
    }
    __syncthreads();
  }
  for (int e = tid; e < n * 32; e += 256) {
    const int row = e / 32, cc = e % 32, col = col0 + cc;
    const double v = Ws[(size_t)row * 33 + cc];
    if (col < n) Wm[(size_t)row * ld + col] = v;
    else if (col == n) yv[row] = v;
  }
}

// P <- P - W^T W  (written in the filter precision; exactly symmetric by construction)
template <class S>
__global__ void __launch_bounds__(256) k_syrk_apply(int n, int ld, const double* __restrict__ Wm, S* __restrict__ P, int ldp) {
  __shared__ double sA[16][33], sB[16][33];
  const int ta = blockIdx.y * 32, tb = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int kb = 0; kb < n; kb += 16) {
    for (int e = threadIdx.x; e < 16 * 32; e += 256) {
      const int kk = e / 32, cc = e % 32, k = kb + kk;
      sA[kk][cc] = (k < n && ta + cc < n) ? Wm[(size_t)k * ld + ta + cc] : 0.0;
      sB[kk][cc] = (k < n && tb + cc < n) ? Wm[(size_t)k * ld + tb + cc] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const double a0 = sA[kk][ty], a1 = sA[kk][ty + 16], b0 = sB[kk][tx], b1 = sB[kk][tx + 16];
      acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int a = ta + ty + 16 * i, b = tb + tx + 16 * j;
      if (a < n && b < n) P[(size_t)a * ldp + b] = (S)((double)P[(size_t)a * ldp + b] - acc[i][j]);
    }
}

// dx = W^T y, then the state correction of msckf.h:1373-1391.  Single CTA.
template <class S>
__global__ void __launch_bounds__(256) k_inject(int n, int ld, int M, const double* __restrict__ Wm, const double* __restrict__ yv,
                                               DevState<S>* st, S* __restrict__ poses, double* __restrict__ dx_out,
                                               const int* __restrict__ m_in, const int* __restrict__ rank_in) {
  extern __shared__ double sdx[];
  const int tid = threadIdx.x;
  if (*m_in == 0) {  // nothing accepted: the reference returns before touching the state (msckf.h:401-403,:1328)
    for (int a = tid; a < n; a += 256) dx_out[a] = 0.0;
    return;
  }
  for (int a = tid; a < n; a += 256) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += Wm[(size_t)k * ld + a] * yv[k];
    sdx[a] = s;
    dx_out[a] = s;
  }
  __syncthreads();
  if (tid == 0) {
    const S dth[3] = {(S)sdx[0], (S)sdx[1], (S)sdx[2]};
    S uq[4], qn[4];
    build_update_quat(dth, uq);
    quat_mul(uq, st->q_IG, qn);  // not renormalised (msckf.h:1376-1378)
    for (int i = 0; i < 4; ++i) st->q_IG[i] = qn[i];
    for (int i = 0; i < 3; ++i) {
      st->b_g[i] += (S)sdx[3 + i];
      st->v_I_G[i] += (S)sdx[6 + i];
      st->b_a[i] += (S)sdx[9 + i];
      st->p_I_G[i] += (S)sdx[12 + i];
    }
    st->n_updates += 1;
    st->last_m = *m_in;
    st->last_rank = *rank_in;
    double nn = 0.0;
    for (int a = 0; a < n; ++a) nn += sdx[a] * sdx[a];
    st->last_dx_norm = sqrt(nn);
  }
  for (int ci = tid; ci < M; ci += 256) {
    S* ps = poses + kPoseStride * ci;
    const S dth[3] = {(S)sdx[15 + 6 * ci], (S)sdx[16 + 6 * ci], (S)sdx[17 + 6 * ci]};
    S uq[4], qn[4];
    build_update_quat(dth, uq);
    quat_mul(uq, ps, qn);
    quat_normalize(qn);
    ps[0] = qn[0]; ps[1] = qn[1]; ps[2] = qn[2]; ps[3] = qn[3];
    ps[4] += (S)sdx[18 + 6 * ci]; ps[5] += (S)sdx[19 + 6 * ci]; ps[6] += (S)sdx[20 + 6 * ci];
  }
}

}  // namespace mb
