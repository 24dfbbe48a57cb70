// msckf_mono_b200/csrc/tail_diag.cuh -- the 32 x 32 diagonal block of the tail's blocked factorisation (round 2).
//
// Round 1 factorised a diagonal block with one warp per matrix, a shuffle + shared-memory round trip per pivot and a second
// pair of warps building the inverses one pivot behind (the S'' warp polling Gamma's keep / drop decision through a
// volatile flag): ~540 ns per pivot, 17 us per block, half of the kernel.  Two measurements shaped this version
// (%globaltimer stamps inside the function, profiles/r02_*):
//   * per-pivot communication is what costs: here the block is itself blocked by panels of 4 columns, the two 4 x 4
//     diagonal micro-blocks (Gamma's and S''s) are factorised REDUNDANTLY in registers by every thread of the two warps that
//     need them -- static indices, right-looking, so the dependent chain of a pivot is rsqrt -> multiply -> one FMA -- and
//     the keep / drop decision of each pivot (Gamma's pivot against thr x its original diagonal, the rank cap; S'' follows)
//     is evaluated by every thread from the same numbers: nobody waits for a flag, and the only synchronisation is two CTA
//     barriers per panel;
//   * code that runs once per call is bound by instruction fetch (~5-7 cycles per instruction from L2 once the body exceeds
//     the 32 KB L1.5 instruction cache; a first version with 8 x 8 micro-blocks and a fully unrolled 32-step inverse was
//     3 157 instructions and took 21 us per block): every loop here is rolled except the 4 x 4 micro-block itself.
// After the eight panels the inverses of the two factors are built block row by block row (8 x 8 blocks):
// Linv_ij = -Linv_ii (sum_k L_ik Linv_kj), two short dot products per element, all threads.  Only the inverses leave the
// block -- the panels below it are X = rows x Linv^T -- so the factor of the diagonal block is never exported.
#pragma once
#include "common.cuh"

namespace mb {

constexpr int kFB = 32;        // block size of the fused form
constexpr int kFLD = kFB + 2;  // row stride of the 32 x 32 shared blocks: even (16-byte loads), conflict-free per quarter warp

struct TfRank { int g, a; };

// 1 / sqrt(p), p > 0 finite: hardware seed (MUFU.RSQ64H, ~2^-22) + two Newton steps in fp64
__device__ __forceinline__ double tf_rsqrt_pos(double p) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  const double hp = 0.5 * p;
  y = y * (1.5 - hp * y * y);
  y = y * (1.5 - hp * y * y);
  return y;
}

__device__ __noinline__ TfRank tf_factor_block(double* __restrict__ DG, double* __restrict__ DA, double* __restrict__ LIG,
                                               double* __restrict__ LIA, double* __restrict__ idg, double* __restrict__ ida,
                                               const double* __restrict__ d0, int kb, int nb, double thr, int rank_cap, bool full,
                                               TfRank rk, int tid, int nthreads, double* __restrict__ TA /*[32][kFLD] scratch*/,
                                               double* __restrict__ TG /*[32][kFLD] scratch*/, unsigned long long* dprof = nullptr) {
  constexpr int LD = kFLD, PW = 4;
  int dpi = 0;
  auto dstamp = [&]() {
    if (dprof && tid == 0 && dpi < 14) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_) : : "memory"); dprof[dpi++] = t_; }
  };
  dstamp();
  const int half = tid >> 5, lane = tid & 31;
#pragma unroll 1
  for (int c0 = 0; c0 < kFB; c0 += PW) {
    if (c0 >= nb) {  // columns beyond the matrix: dropped indices
      if (tid < PW) { ida[c0 + tid] = 0.0; idg[c0 + tid] = 0.0; }
      continue;
    }
    if (tid < 64) {
    // (1) the two 4 x 4 diagonal micro-blocks, redundantly in every thread of warps 0 and 1 (the row solvers of (2) need them
    // in registers; all eight warps doing it would make the half-rate FP64 pipe, not the pivot chain, the limit)
    double Gm[PW][PW], Am[PW][PW], ivg[PW], iva[PW];
#pragma unroll
    for (int a = 0; a < PW; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        Am[a][b] = DA[(c0 + a) * LD + c0 + b];
        Gm[a][b] = full ? 0.0 : DG[(c0 + a) * LD + c0 + b];
      }
#pragma unroll
    for (int j = 0; j < PW; ++j) {
      const double dk0 = d0[kb + c0 + j];
      const double pg = Gm[j][j], pa = Am[j][j];
      bool drop;
      if (full) drop = (c0 + j >= nb) || !(dk0 > 0.0) || rk.a >= rank_cap;
      else drop = (c0 + j >= nb) || !(dk0 > 0.0) || !(pg > thr * dk0) || rk.g >= rank_cap;
      if (!drop) rk.g++;
      const double ig = (drop || full) ? 0.0 : tf_rsqrt_pos(pg);
      const bool dropa = drop || !(pa > 0.0);
      if (!dropa) rk.a++;
      const double ia = dropa ? 0.0 : tf_rsqrt_pos(pa);
      ivg[j] = ig; iva[j] = ia;
#pragma unroll
      for (int a = j + 1; a < PW; ++a) { Gm[a][j] *= ig; Am[a][j] *= ia; }
#pragma unroll
      for (int b = j + 1; b < PW; ++b)
#pragma unroll
        for (int a = b; a < PW; ++a) { Gm[a][b] -= Gm[a][j] * Gm[b][j]; Am[a][b] -= Am[a][j] * Am[b][j]; }
    }
    if (tid < PW) {
      double vg = 0.0, va = 0.0;
#pragma unroll
      for (int j = 0; j < PW; ++j) if (j == tid) { vg = ivg[j]; va = iva[j]; }
      idg[c0 + tid] = vg; ida[c0 + tid] = va;
    }
    // (2) one row of the panel per thread (rows c0 .. nb-1; warp 0: S'', warp 1: Gamma), solved in registers; a row inside the
    // micro-block reproduces the factor's own row (entries right of its diagonal are masked)
    const int rrow = c0 + lane;
    if (rrow < nb && !(half == 1 && full)) {
      double* D = half ? DG : DA;
      double x[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) x[j] = (c0 + j <= rrow) ? D[rrow * LD + c0 + j] : 0.0;
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        x[j] *= half ? ivg[j] : iva[j];
        if (c0 + j > rrow) x[j] = 0.0;
#pragma unroll
        for (int jj = j + 1; jj < PW; ++jj) x[jj] -= x[j] * (half ? Gm[jj][j] : Am[jj][j]);
      }
#pragma unroll
      for (int j = 0; j < PW; ++j) if (c0 + j <= rrow) D[rrow * LD + c0 + j] = x[j];
    }
    }
    __syncthreads();
    // (3) trailing update inside the block: rows / columns c0+4 .. nb-1, lower triangle, both matrices
    const int q0 = c0 + PW, nr = nb - q0;
    if (nr > 0) {
      const int per = nr * nr, tot = full ? per : 2 * per;
#pragma unroll 1
      for (int e = tid; e < tot; e += nthreads) {
        const int mtx = e >= per ? 1 : 0, r2 = e - mtx * per, ii = r2 / nr, i = q0 + ii, j = q0 + (r2 - ii * nr);
        if (j > i) continue;
        double* D = mtx ? DG : DA;
        const double2 a01 = *reinterpret_cast<const double2*>(D + i * LD + c0), a23 = *reinterpret_cast<const double2*>(D + i * LD + c0 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(D + j * LD + c0), b23 = *reinterpret_cast<const double2*>(D + j * LD + c0 + 2);
        D[i * LD + j] -= (a01.x * b01.x + a01.y * b01.y) + (a23.x * b23.x + a23.y * b23.y);
      }
    }
    __syncthreads();
  }
  dstamp();  // factors done
  // ---- inverses, 8 x 8 blocks.  Diagonal blocks first: one column per thread (2 matrices x 4 blocks x 8 columns = 64 threads),
  // right-looking substitution inside the 8 x 8 block, the column kept in the output array (conflict-free: thread = column).
  constexpr int IB = 8;
  if (tid < 64 && !(half == 1 && full)) {
    const double* D = half ? DG : DA;
    const double* iv = half ? idg : ida;
    double* LI = half ? LIG : LIA;
    const int c = lane, b0 = c & ~(IB - 1);
#pragma unroll 1
    for (int k = 0; k < kFB; ++k) LI[k * LD + c] = (k == c) ? 1.0 : 0.0;  // (also zeroes everything above / outside)
#pragma unroll 1
    for (int j = c; j < b0 + IB; ++j) {
      const double y = LI[j * LD + c] * iv[j];
      LI[j * LD + c] = y;
#pragma unroll 1
      for (int k = j + 1; k < b0 + IB; ++k) LI[k * LD + c] -= D[k * LD + j] * y;
    }
  }
  __syncthreads();
  // block rows 1..3: T = sum_k L_ik Linv_kj over the finished block rows, then Linv_ij = -Linv_ii T
#pragma unroll 1
  for (int bi = 1; bi < kFB / IB; ++bi) {
    const int r0 = bi * IB, ncol = r0;  // columns 0 .. r0-1
    const int per = IB * ncol, tot = full ? per : 2 * per;
#pragma unroll 1
    for (int e = tid; e < tot; e += nthreads) {
      const int mtx = e >= per ? 1 : 0, r2 = e - mtx * per, rr = r2 / ncol, c = r2 - rr * ncol, r = r0 + rr;
      const double* D = mtx ? DG : DA;
      const double* LI = mtx ? LIG : LIA;
      double s0 = 0.0, s1 = 0.0;
      int k = c;
      for (; k + 1 < r0; k += 2) { s0 += D[r * LD + k] * LI[k * LD + c]; s1 += D[r * LD + k + 1] * LI[(k + 1) * LD + c]; }
      if (k < r0) s0 += D[r * LD + k] * LI[k * LD + c];
      (mtx ? TG : TA)[rr * LD + c] = s0 + s1;
    }
    __syncthreads();
#pragma unroll 1
    for (int e = tid; e < tot; e += nthreads) {
      const int mtx = e >= per ? 1 : 0, r2 = e - mtx * per, rr = r2 / ncol, c = r2 - rr * ncol, r = r0 + rr;
      double* LI = mtx ? LIG : LIA;
      const double* T = mtx ? TG : TA;
      double s = 0.0;
      for (int k = 0; k <= rr; ++k) s += LI[r * LD + r0 + k] * T[k * LD + c];
      LI[r * LD + c] = -s;
    }
    __syncthreads();
  }
  dstamp();  // inverses
  return rk;
}

}  // namespace mb
