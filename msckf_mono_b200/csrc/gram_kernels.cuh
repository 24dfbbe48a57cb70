// msckf_mono_b200/csrc/gram_kernels.cuh
// Compression of the stacked measurement (msckf.h:1343-1366) in projector / Gram form.
//
// The reference keeps Q_1 = the first n columns of the Householder Q of the stacked H_o (rows 0..14 pass
// through because the 15 leading columns of H_o are zero) and forms T_H = Q_1^T H_o, r_n = Q_1^T r_o,
// R_n = Q_1^T R_o Q_1.  The EKF update only depends on span(Q_1) = span(E_15) + range(H_o) (+ directions
// of numerically zero pivots, which are rounding noise in the reference).  We use the basis
// Q'' = [E_15 | H_c] (H_c = H_o(:,15:)) of that subspace directly:
//     T'' = Q''^T H_o   = [[0, H_o(0:15,15:)], [0, Lambda]],    Lambda = H_c^T H_c
//     r'' = Q''^T r_o   = [r_o(0:15); beta],                    beta   = H_c^T r_o
//     R'' = Q''^T R_o Q''= [[R_o(0:15,0:15), (R_o H_c)(0:15,:)], [sym, Psi]],  Psi = H_c^T R_o H_c
// and since H_o,j = A_j^T X_j with A_j A_j^T = I - U_j U_j^T (U_j = first 3 columns of the per-feature Q),
//     Lambda = blkdiag(sum X_ji^T X_ji)     - Z^T Z            Z_j  = U_j^T X_j
//     Psi    = blkdiag(sum X_ji^T D X_ji)   - Z^T Yq - Yq^T Z   Yq_j = U_j^T D X_j - 1/2 (U_j^T D U_j) Z_j
//     beta   = blk(sum X_ji^T r_ji)         - Z^T (U^T r)
// so the stacked m x n matrix is never materialised: the Gram terms are one (c x 3N)(3N x c) GEMM pair.
// All accumulation is fp64 (for both filter precisions).
#pragma once
#include "common.cuh"

namespace mb {

constexpr int GT = 32;   // output tile
constexpr int GK = 16;   // k-step

// Partial Gram products over a K-split: G1p[s] = Z^T Z, G2p[s] = Z^T Yq + Yq^T Z, upper tiles only.
__global__ void __launch_bounds__(256) k_gram(const double* __restrict__ Z, const double* __restrict__ Yq, int K, int c,
                                             int kchunk, double* __restrict__ G1p, double* __restrict__ G2p) {
  __shared__ double sZa[GK][GT + 1], sZb[GK][GT + 1], sYa[GK][GT + 1], sYb[GK][GT + 1];
  // decode the (ta <= tb) tile pair
  const int ntile = (c + GT - 1) / GT;
  int pidx = blockIdx.x, ta = 0;
  while (pidx >= ntile - ta) { pidx -= ntile - ta; ++ta; }
  const int tb = ta + pidx;
  const int split = blockIdx.y;
  const int k0 = split * kchunk, k1 = min(K, k0 + kchunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 2x2 outputs each
  double a1[2][2] = {{0, 0}, {0, 0}}, a2[2][2] = {{0, 0}, {0, 0}};
  for (int kb = k0; kb < k1; kb += GK) {
    for (int e = threadIdx.x; e < GK * GT; e += 256) {
      const int kk = e / GT, cc = e % GT;
      const int k = kb + kk;
      const int ca = ta * GT + cc, cb = tb * GT + cc;
      const bool kin = k < k1;
      sZa[kk][cc] = (kin && ca < c) ? Z[(size_t)k * c + ca] : 0.0;
      sYa[kk][cc] = (kin && ca < c) ? Yq[(size_t)k * c + ca] : 0.0;
      sZb[kk][cc] = (kin && cb < c) ? Z[(size_t)k * c + cb] : 0.0;
      sYb[kk][cc] = (kin && cb < c) ? Yq[(size_t)k * c + cb] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const double za0 = sZa[kk][ty], za1 = sZa[kk][ty + 16], ya0 = sYa[kk][ty], ya1 = sYa[kk][ty + 16];
      const double zb0 = sZb[kk][tx], zb1 = sZb[kk][tx + 16], yb0 = sYb[kk][tx], yb1 = sYb[kk][tx + 16];
      a1[0][0] += za0 * zb0; a1[0][1] += za0 * zb1; a1[1][0] += za1 * zb0; a1[1][1] += za1 * zb1;
      a2[0][0] += za0 * yb0 + ya0 * zb0; a2[0][1] += za0 * yb1 + ya0 * zb1;
      a2[1][0] += za1 * yb0 + ya1 * zb0; a2[1][1] += za1 * yb1 + ya1 * zb1;
    }
    __syncthreads();
  }
  double* o1 = G1p + (size_t)split * c * c;
  double* o2 = G2p + (size_t)split * c * c;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ra = ta * GT + ty + 16 * i, cb = tb * GT + tx + 16 * j;
      if (ra < c && cb < c) { o1[(size_t)ra * c + cb] = a1[i][j]; o2[(size_t)ra * c + cb] = a2[i][j]; }
    }
}

// Block-diagonal terms, one CTA per clone: D1 = sum X^T X, D2 = sum X^T D X (6x6), b = sum X^T r (6),
// over the accepted tracks' observations of that clone.  Deterministic (fixed feature->thread map, tree reduce).
template <class S>
__global__ void __launch_bounds__(128) k_blockdiag(int N, const int* __restrict__ obs_off, const int* __restrict__ clone_idx,
                                                  const int* __restrict__ accept, const S* __restrict__ Xg, const S* __restrict__ rg,
                                                  double du, double dv, double* __restrict__ D1, double* __restrict__ D2,
                                                  double* __restrict__ bb) {
  const int clone = blockIdx.x;
  double acc[78];
#pragma unroll
  for (int k = 0; k < 78; ++k) acc[k] = 0.0;
  for (int j = threadIdx.x; j < N; j += 128) {
    if (!accept[j]) continue;
    const int o0 = obs_off[j], o1 = obs_off[j + 1];
    for (int o = o0; o < o1; ++o) {
      if (clone_idx[o] != clone) continue;
      double x0[6], x1[6];
#pragma unroll
      for (int b = 0; b < 6; ++b) { x0[b] = (double)Xg[12 * (size_t)o + b]; x1[b] = (double)Xg[12 * (size_t)o + 6 + b]; }
      const double r0 = (double)rg[2 * (size_t)o], r1 = (double)rg[2 * (size_t)o + 1];
#pragma unroll
      for (int p = 0; p < 6; ++p) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          acc[6 * p + q] += x0[p] * x0[q] + x1[p] * x1[q];
          acc[36 + 6 * p + q] += du * x0[p] * x0[q] + dv * x1[p] * x1[q];
        }
        acc[72 + p] += x0[p] * r0 + x1[p] * r1;
      }
    }
  }
  __shared__ double red[4][78];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 78; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 78) {
    const int k = threadIdx.x;
    const double v = (red[0][k] + red[1][k]) + (red[2][k] + red[3][k]);
    if (k < 36) D1[36 * (size_t)clone + k] = v;
    else if (k < 72) D2[36 * (size_t)clone + (k - 36)] = v;
    else bb[6 * (size_t)clone + (k - 72)] = v;
  }
}

// Assemble T'' (n x n), r'' (n), R'' (n x n) except the <=15 head rows (k_head fills those afterwards).
__global__ void __launch_bounds__(256) k_assemble(int n, int ld, int K, int nsplit, const double* __restrict__ G1p,
                                                 const double* __restrict__ G2p, const double* __restrict__ D1,
                                                 const double* __restrict__ D2, const double* __restrict__ bb,
                                                 const double* __restrict__ Z, const double* __restrict__ ur,
                                                 double* __restrict__ T2, double* __restrict__ R2, double* __restrict__ r2) {
  const int c = n - kImuDim;
  const size_t total = (size_t)n * n;
  for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
    const int a = (int)(e / n), b = (int)(e % n);
    double tv = 0.0, rv = 0.0;
    if (a >= kImuDim && b >= kImuDim) {
      const int ac = a - kImuDim, bc = b - kImuDim;
      // the Gram kernel wrote upper tiles only
      const bool upper = (ac / GT) <= (bc / GT);
      const size_t off = upper ? ((size_t)ac * c + bc) : ((size_t)bc * c + ac);
      double g1 = 0.0, g2 = 0.0;
      for (int s = 0; s < nsplit; ++s) { g1 += G1p[(size_t)s * c * c + off]; g2 += G2p[(size_t)s * c * c + off]; }
      double d1 = 0.0, d2 = 0.0;
      if (ac / 6 == bc / 6) { d1 = D1[36 * (size_t)(ac / 6) + 6 * (ac % 6) + (bc % 6)]; d2 = D2[36 * (size_t)(ac / 6) + 6 * (ac % 6) + (bc % 6)]; }
      tv = d1 - g1;
      rv = d2 - g2;
    }
    T2[(size_t)a * ld + b] = tv;
    R2[(size_t)a * ld + b] = rv;
  }
  // beta
  for (int a = blockIdx.x * 256 + threadIdx.x; a < n; a += gridDim.x * 256) {
    double v = 0.0;
    if (a >= kImuDim) {
      const int ac = a - kImuDim;
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += Z[(size_t)k * c + ac] * ur[k];
      v = bb[ac] - s;
    }
    r2[a] = v;
  }
}

// Head rows: stacked rows 0..14 pass through the reference's QR untouched (msckf.h:1343-1363 with 15 zero
// leading columns).  They belong to the first accepted feature(s) in stacking order.  Single CTA.
template <class S>
__global__ void __launch_bounds__(256) k_head(int N, int n, int ld, const int* __restrict__ obs_off, const int* __restrict__ clone_idx,
                                             const int* __restrict__ accept, const int* __restrict__ row_off,
                                             const S* __restrict__ Xg, const S* __restrict__ rg, const S* __restrict__ Vg,
                                             const S* __restrict__ taug, const double* __restrict__ Z, double du, double dv,
                                             double* __restrict__ T2, double* __restrict__ R2, double* __restrict__ r2, int Lmax) {
  extern __shared__ double sh[];  // cols[18][2*Lmax]
  const int c = n - kImuDim;
  const int tid = threadIdx.x;
  __shared__ double s_dot[18];
  __shared__ double s_adu[15][3];
  for (int j = 0; j < N; ++j) {
    if (!accept[j]) continue;
    const int h0 = row_off[j];
    if (h0 >= kImuDim) break;  // offsets are non-decreasing
    const int o0 = obs_off[j], L = obs_off[j + 1] - o0, L2 = 2 * L;
    const int rho = L2 - 3;
    const int nh = min(rho, kImuDim - h0);
    const int ncol = 3 + nh;
    const int ldc = 2 * Lmax;
    // cols(:,q) = Q e_q = H0 H1 H2 e_q, q = 0..2+nh  (q<3: U_j, q>=3: the head columns of A_j)
    for (int e = tid; e < ncol * L2; e += 256) { const int q = e / L2, row = e % L2; sh[q * ldc + row] = (row == q) ? 1.0 : 0.0; }
    __syncthreads();
    for (int k = 2; k >= 0; --k) {
      const double tk = (double)taug[3 * j + k];
      if (tid < ncol) {
        double s = 0.0;
        for (int row = k; row < L2; ++row) s += (double)Vg[3 * (2 * (size_t)o0 + row) + k] * sh[tid * ldc + row];
        s_dot[tid] = tk * s;
      }
      __syncthreads();
      for (int e = tid; e < ncol * L2; e += 256) {
        const int q = e / L2, row = e % L2;
        if (row >= k) sh[q * ldc + row] -= s_dot[q] * (double)Vg[3 * (2 * (size_t)o0 + row) + k];
      }
      __syncthreads();
    }
    // A_h^T D U (nh x 3)
    if (tid < nh * 3) {
      const int t = tid / 3, q = tid % 3;
      double s = 0.0;
      for (int row = 0; row < L2; ++row) s += sh[(3 + t) * ldc + row] * ((row & 1) ? dv : du) * sh[q * ldc + row];
      s_adu[t][q] = s;
    }
    __syncthreads();
    // r_h, R_hh
    for (int e = tid; e < nh * (nh + 1); e += 256) {
      const int t = e / (nh + 1), u = e % (nh + 1);
      double s = 0.0;
      if (u == nh) {
        for (int row = 0; row < L2; ++row) s += sh[(3 + t) * ldc + row] * (double)rg[2 * (size_t)o0 + row];
        r2[h0 + t] = s;
      } else {
        for (int row = 0; row < L2; ++row) s += sh[(3 + t) * ldc + row] * ((row & 1) ? dv : du) * sh[(3 + u) * ldc + row];
        R2[(size_t)(h0 + t) * ld + (h0 + u)] = s;
      }
    }
    // H_h rows and R_hH rows: nonzero only in the track's clone blocks
    for (int e = tid; e < nh * L * 6; e += 256) {
      const int t = e / (L * 6), rem = e % (L * 6), i = rem / 6, b = rem % 6;
      const int col = 6 * clone_idx[o0 + i] + b;
      const double x0 = (double)Xg[12 * (size_t)(o0 + i) + b], x1 = (double)Xg[12 * (size_t)(o0 + i) + 6 + b];
      const double a0 = sh[(3 + t) * ldc + 2 * i], a1 = sh[(3 + t) * ldc + 2 * i + 1];
      const double hv = a0 * x0 + a1 * x1;
      double rv = du * a0 * x0 + dv * a1 * x1;
      for (int q = 0; q < 3; ++q) rv -= s_adu[t][q] * Z[(size_t)(3 * j + q) * c + col];
      T2[(size_t)(h0 + t) * ld + kImuDim + col] = hv;
      R2[(size_t)(h0 + t) * ld + kImuDim + col] = rv;
      R2[(size_t)(kImuDim + col) * ld + (h0 + t)] = rv;
    }
    __syncthreads();
  }
}

}  // namespace mb
