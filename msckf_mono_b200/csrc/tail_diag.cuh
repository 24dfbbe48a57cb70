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
  const int half = (tid >> 5) & 1, lane = tid & 31;
  const bool extra = tid >= 64;  // warps 2, 3: the rows of the identity that turn into the inverses (see below)
  // The inverses ride along as 32 EXTRA ROWS per matrix: solving  Y L^T = I  row by row is exactly what the panel algorithm
  // does to the rows below the block (micro-solve against the diagonal micro-block, then the trailing update), and
  // Y = L^-T, i.e. extra row i, column j holds Linv[j][i] -- stored transposed, straight into LIA / LIG.  Row i of the identity
  // is zero left of column i, so it only joins from the panel that contains column i.  This replaces a separate inverse
  // stage (6 us and 7 CTA barriers per block in the first version of this file).
  for (int e = tid; e < kFB * LD; e += nthreads) {
    const int k = e / LD, c = e % LD;
    LIA[e] = (k == c) ? 1.0 : 0.0;
    LIG[e] = (k == c) ? 1.0 : 0.0;
  }
  __syncthreads();
#pragma unroll 1
  for (int c0 = 0; c0 < kFB; c0 += PW) {
    if (c0 >= nb) {  // columns beyond the matrix: dropped indices
      if (tid < PW) { ida[c0 + tid] = 0.0; idg[c0 + tid] = 0.0; }
      continue;
    }
    if (tid < 128) {
    // (1) the two 4 x 4 diagonal micro-blocks, redundantly in every thread of warps 0..3 (the row solvers of (2) need them
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
    if (!extra && rrow < nb && !(half == 1 && full)) {
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
    if (extra && lane < c0 + PW && lane < nb && !(half == 1 && full)) {  // extra row `lane` of the identity, active from its own column on
      double* LI = half ? LIG : LIA;
      double x[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) x[j] = LI[(c0 + j) * LD + lane];
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        x[j] *= half ? ivg[j] : iva[j];
#pragma unroll
        for (int jj = j + 1; jj < PW; ++jj) x[jj] -= x[j] * (half ? Gm[jj][j] : Am[jj][j]);
      }
#pragma unroll
      for (int j = 0; j < PW; ++j) LI[(c0 + j) * LD + lane] = x[j];
    }
    }
    __syncthreads();
    // (3) trailing update inside the block: rows / columns c0+4 .. nb-1, lower triangle, both matrices -- and of the extra rows:
    // rhs[i][j] -= sum_k Y[i][c0+k] L[j][c0+k] for the active rows i < q0, columns j >= q0.  Threads as a 16 x 16 grid (no
    // index divisions on this path: every instruction here is on the critical chain of the whole kernel).
    const int q0 = c0 + PW, nr = nb - q0;
    if (nr > 0) {
      const int ty = tid >> 4, tx = tid & 15;
#pragma unroll 1
      for (int i = q0 + ty; i < nb; i += 16) {
        const double2 a01 = *reinterpret_cast<const double2*>(DA + i * LD + c0), a23 = *reinterpret_cast<const double2*>(DA + i * LD + c0 + 2);
        const double2 g01 = *reinterpret_cast<const double2*>(DG + i * LD + c0), g23 = *reinterpret_cast<const double2*>(DG + i * LD + c0 + 2);
#pragma unroll 1
        for (int j = q0 + tx; j <= i; j += 16) {
          const double2 b01 = *reinterpret_cast<const double2*>(DA + j * LD + c0), b23 = *reinterpret_cast<const double2*>(DA + j * LD + c0 + 2);
          DA[i * LD + j] -= (a01.x * b01.x + a01.y * b01.y) + (a23.x * b23.x + a23.y * b23.y);
          if (!full) {
            const double2 h01 = *reinterpret_cast<const double2*>(DG + j * LD + c0), h23 = *reinterpret_cast<const double2*>(DG + j * LD + c0 + 2);
            DG[i * LD + j] -= (g01.x * h01.x + g01.y * h01.y) + (g23.x * h23.x + g23.y * h23.y);
          }
        }
      }
      const int nact = min(q0, nb);
#pragma unroll 1
      for (int j = q0 + ty; j < nb; j += 16) {
        const double2 b01 = *reinterpret_cast<const double2*>(DA + j * LD + c0), b23 = *reinterpret_cast<const double2*>(DA + j * LD + c0 + 2);
        const double2 h01 = *reinterpret_cast<const double2*>(DG + j * LD + c0), h23 = *reinterpret_cast<const double2*>(DG + j * LD + c0 + 2);
#pragma unroll 1
        for (int i = tx; i < nact; i += 16) {
          LIA[j * LD + i] -= (LIA[c0 * LD + i] * b01.x + LIA[(c0 + 1) * LD + i] * b01.y) + (LIA[(c0 + 2) * LD + i] * b23.x + LIA[(c0 + 3) * LD + i] * b23.y);
          if (!full) LIG[j * LD + i] -= (LIG[c0 * LD + i] * h01.x + LIG[(c0 + 1) * LD + i] * h01.y) + (LIG[(c0 + 2) * LD + i] * h23.x + LIG[(c0 + 3) * LD + i] * h23.y);
        }
      }
    }
    __syncthreads();
  }
  dstamp();  // factors + inverses done
  // rows / columns of dropped or out-of-range pivots of the inverses are zero by construction (1 / L_kk := 0)
  return rk;
}

}  // namespace mb
