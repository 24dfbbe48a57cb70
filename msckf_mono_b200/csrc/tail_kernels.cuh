// msckf_mono_b200/csrc/tail_kernels.cuh
// Dense n x n tail of the EKF update (msckf.h:1369-1418) in square-root form, fp64 throughout:
//   TP  = T'' P                      S'' = TP T''^T + R''          (k_gemm_tp, k_gemm_s: this file, all SMs)
//   S'' = L L^T over the independent rows, W = L^-1 [TP | r''], P+ = P - W^T W (== Joseph form for the optimal
//   gain of the projected system), dx = W^T (L^-1 r''), state injection       (k_tail: tail_cluster.cuh, one cluster)
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

}  // namespace mb
