// msckf_mono_b200/csrc/tail_cluster.cuh
// The serial part of the EKF tail as ONE thread-block-cluster kernel (8 CTAs on 8 SMs of one GPC):
//     Gamma (Gram matrix of the basis)      -> rank decision        }  rank-revealing Cholesky of Gamma and S''
//     S'' = L L^T over the kept indices                              }  in lockstep, blocked by NB
//     W = L^-1 [T''P | r'']                                         (forward substitution, RHS columns sharded)
//     P <- P - W^T W,   dx = W^T y,  state injection                (msckf.h:1373-1418)
// The matrices stay in global memory (they are L2 resident: n <= 639, fp64); cluster barriers (release/acquire
// at cluster scope, ~0.3 us) order the phases, so the O(n^3) trailing updates, the substitutions and the SYRK
// are spread over the cluster's SMs while the O(n NB^2) diagonal-block factorisations run on CTA 0.
// Inner loops are register tiled (4x4 outputs per thread, operands read as contiguous vectors from a transposed
// panel in shared memory): the fp64 FMA pipe, not the shared-memory port, is the limit.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace mb {
namespace cg = cooperative_groups;

constexpr int kTailThreads = 512;

template <int NB>
__device__ __forceinline__ void tc_load_diag(double* D, const double* A, int ld, int kb, int nb, int tid) {
  for (int e = tid; e < NB * NB; e += kTailThreads) {
    const int i = e / NB, j = e % NB;
    double v = (i == j) ? 1.0 : 0.0;  // identity padding when nb < NB
    if (i < nb && j < nb) v = (j <= i) ? A[(size_t)(kb + i) * ld + kb + j] : 0.0;
    D[i * (NB + 1) + j] = v;
  }
}

// one panel row: x D^T = a  (row in registers; dropped columns give 0)
template <int NB>
__device__ __forceinline__ void tc_panel_row(double* Mx, int ld, int row, int kb, int nb, const double* D, const int* bkeep) {
  constexpr int LD = NB + 1;
  double x[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? Mx[(size_t)row * ld + kb + j] : 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    double v = x[j];
#pragma unroll
    for (int cc = 0; cc < j; ++cc) v -= x[cc] * D[j * LD + cc];
    x[j] = ((j < nb) && bkeep[j]) ? v / D[j * LD + j] : 0.0;
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j < nb) Mx[(size_t)row * ld + kb + j] = x[j];
}

// C[4][4] += sum_j At[j][ra..ra+3] * Bt[j][rb..rb+3] over the NB panel columns (transposed panels, row stride ldt)
template <int NB>
__device__ __forceinline__ void tc_tile_4x4(const double* At, const double* Bt, int ldt, int ra, int rb, double acc[4][4]) {
#pragma unroll 4
  for (int j = 0; j < NB; ++j) {
    const double2 a01 = *reinterpret_cast<const double2*>(At + (size_t)j * ldt + ra);
    const double2 a23 = *reinterpret_cast<const double2*>(At + (size_t)j * ldt + ra + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(Bt + (size_t)j * ldt + rb);
    const double2 b23 = *reinterpret_cast<const double2*>(Bt + (size_t)j * ldt + rb + 2);
    const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[p][q] += a[p] * b[q];
  }
}

template <class S, int NB>
__global__ void __launch_bounds__(kTailThreads) k_tail(int n, int ld, int M, const double* __restrict__ T2, double* __restrict__ G,
                                                      double* __restrict__ A, int* __restrict__ keep, double thr,
                                                      int* __restrict__ rank_out, const int* __restrict__ m_in,
                                                      const double* __restrict__ TP, const double* __restrict__ r2,
                                                      double* __restrict__ Wm, double* __restrict__ yv, S* __restrict__ P, int ldp,
                                                      DevState<S>* st, S* __restrict__ poses, double* __restrict__ dx_out,
                                                      unsigned long long* __restrict__ prof /*optional phase timestamps*/) {
  cg::cluster_group cluster = cg::this_cluster();
  int prof_i = 0;
  auto stamp = [&]() {
    if (prof && blockIdx.x == 0 && threadIdx.x == 0 && prof_i < 64) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      prof[prof_i++] = t;
    }
  };
  stamp();
  const int crank = (int)cluster.block_rank();
  const int C = (int)cluster.num_blocks();
  const int tid = threadIdx.x;
  constexpr int LD = NB + 1;
  extern __shared__ __align__(16) double sm[];
  const int ldt = (n + 3) & ~3;              // transposed-panel row stride (16-byte aligned rows)
  double* DG = sm;                           // [NB][LD]
  double* DA = DG + NB * LD;                 // [NB][LD]
  double* PT_A = DA + NB * LD + ((NB * LD) & 1);  // [NB][ldt]  panel of A, transposed  (also W slab / SYRK tiles later)
  double* PT_G = PT_A + (size_t)NB * ldt;    // [NB][ldt]  panel of G, transposed
  double* d0 = PT_G + (size_t)NB * ldt;      // [n] original diagonal of Gamma (CTA 0)
  int* bkeep = reinterpret_cast<int*>(d0 + ((n + 1) & ~1));  // [NB] keep flags of the current block
  __shared__ double dgk[NB], dak[NB];
  __shared__ int s_rank;
  const int m = *m_in;
  const bool full = m <= n;  // all rows explicit and orthonormal: Gamma = I_m, nothing to decide, G untouched
  const int rank_cap = min(m, n);
  const int gtid = crank * kTailThreads + tid, gthreads = C * kTailThreads;

  if (m == 0) {  // nothing accepted: the reference returns before touching the state (msckf.h:401-403, :1328)
    for (int a = gtid; a < n; a += gthreads) dx_out[a] = 0.0;
    if (gtid == 0) *rank_out = 0;
    return;
  }
  // ---------------------------------------------------------------- Gamma = [[I_h, H_h], [H_h^T, Lambda]] from T''
  if (!full) {
    const size_t total = (size_t)n * n;
    for (size_t e = gtid; e < total; e += gthreads) {
      const int a = (int)(e / n), b = (int)(e % n);
      double v;
      if (a < kImuDim && b < kImuDim) v = (a == b) ? 1.0 : 0.0;
      else if (a < kImuDim) v = T2[(size_t)a * ld + b];
      else if (b < kImuDim) v = T2[(size_t)b * ld + a];
      else v = T2[(size_t)a * ld + b];
      G[(size_t)a * ld + b] = v;
    }
  }
  cluster.sync();
  stamp();  // [1] gamma built
  if (crank == 0) {
    for (int k = tid; k < n; k += kTailThreads) d0[k] = full ? (k < m ? 1.0 : 0.0) : G[(size_t)k * ld + k];
    if (tid == 0) s_rank = 0;
  }
  __syncthreads();
  // ---------------------------------------------------------------- blocked rank-revealing Cholesky (G decides, A follows)
  for (int kb = 0; kb < n; kb += NB) {
    const int nb = min(NB, n - kb);
    const int r0 = kb + nb;
    // phase 1 (CTA 0): diagonal blocks
    if (crank == 0) {
      if (!full) tc_load_diag<NB>(DG, G, ld, kb, nb, tid);
      tc_load_diag<NB>(DA, A, ld, kb, nb, tid);
      int rank_now = s_rank;
      for (int k = 0; k < nb; ++k) {
        __syncthreads();
        const double pg = full ? 1.0 : DG[k * LD + k], pa = DA[k * LD + k], dk0 = d0[kb + k];
        const bool drop = !(dk0 > 0.0) || !(pg > thr * dk0) || rank_now >= rank_cap || !(pa > 0.0);
        if (!drop) rank_now++;
        double lg = 1.0, la = 1.0;
        if (tid < 32 && !drop) { lg = sqrt(pg); la = sqrt(pa); }   // one warp computes the roots ...
        if (tid == 0) { dgk[k] = lg; dak[k] = la; bkeep[k] = drop ? 0 : 1; }
        __syncthreads();                                         // ... everybody reads them
        lg = dgk[k]; la = dak[k];
        for (int i = k + 1 + tid; i < nb; i += kTailThreads) {
          if (drop) { DA[i * LD + k] = 0.0; if (!full) DG[i * LD + k] = 0.0; }
          else { DA[i * LD + k] /= la; if (!full) DG[i * LD + k] /= lg; }
        }
        if (drop)
          for (int j = tid; j < k; j += kTailThreads) { DA[k * LD + j] = 0.0; if (!full) DG[k * LD + j] = 0.0; }
        __syncthreads();
        if (!drop) {
          const int rem = nb - k - 1;
          for (int e = tid; e < rem * rem; e += kTailThreads) {
            const int i = k + 1 + e / rem, j = k + 1 + e % rem;
            if (j <= i) {
              DA[i * LD + j] -= DA[i * LD + k] * DA[j * LD + k];
              if (!full) DG[i * LD + j] -= DG[i * LD + k] * DG[j * LD + k];
            }
          }
        }
      }
      __syncthreads();
      if (tid == 0) s_rank = rank_now;
      if (tid < nb) { DG[tid * LD + tid] = dgk[tid]; DA[tid * LD + tid] = dak[tid]; keep[kb + tid] = bkeep[tid]; }
      __syncthreads();
      for (int e = tid; e < nb * nb; e += kTailThreads) {
        const int i = e / nb, j = e % nb;
        if (j <= i) { A[(size_t)(kb + i) * ld + kb + j] = DA[i * LD + j]; if (!full) G[(size_t)(kb + i) * ld + kb + j] = DG[i * LD + j]; }
      }
    }
    stamp();  // diag block done (CTA 0)
    if (r0 >= n) break;  // last block: nothing below
    cluster.sync();
    // phase 2 (all CTAs): panel rows below the block, one row per thread
    if (crank != 0) {
      if (!full) tc_load_diag<NB>(DG, G, ld, kb, nb, tid);
      tc_load_diag<NB>(DA, A, ld, kb, nb, tid);
      if (tid < NB) bkeep[tid] = (tid < nb) ? keep[kb + tid] : 0;
    }
    __syncthreads();
    for (int row = r0 + gtid; row < n; row += gthreads) {
      tc_panel_row<NB>(A, ld, row, kb, nb, DA, bkeep);
      if (!full) tc_panel_row<NB>(G, ld, row, kb, nb, DG, bkeep);
    }
    cluster.sync();
    stamp();  // panel done
    // phase 3 (all CTAs): trailing update, 4x4 register tiles on the transposed panel
    const int nr = n - r0;
    for (int e = tid; e < nr * NB; e += kTailThreads) {
      const int i = e / NB, j = e % NB;
      PT_A[(size_t)j * ldt + i] = (j < nb) ? A[(size_t)(r0 + i) * ld + kb + j] : 0.0;
      if (!full) PT_G[(size_t)j * ldt + i] = (j < nb) ? G[(size_t)(r0 + i) * ld + kb + j] : 0.0;
    }
    {  // zero the tail of the padded rows so that partial tiles read zeros
      const int nrp = (nr + 3) & ~3;
      for (int e = tid; e < (nrp - nr) * NB; e += kTailThreads) {
        const int j = e / (nrp - nr), i = nr + e % (nrp - nr);
        PT_A[(size_t)j * ldt + i] = 0.0;
        if (!full) PT_G[(size_t)j * ldt + i] = 0.0;
      }
    }
    __syncthreads();
    const int nt = (nr + 3) / 4, ntile = nt * (nt + 1) / 2;
    for (int tl = gtid; tl < ntile; tl += gthreads) {
      int ti = (int)((sqrtf(8.0f * (float)tl + 1.0f) - 1.0f) * 0.5f);
      while (ti * (ti + 1) / 2 > tl) --ti;
      while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
      const int tj = tl - ti * (ti + 1) / 2;
      double acc[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
      tc_tile_4x4<NB>(PT_A, PT_A, ldt, 4 * ti, 4 * tj, acc);
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = 4 * ti + p, i2 = 4 * tj + q;
          if (i < nr && i2 <= i) A[(size_t)(r0 + i) * ld + r0 + i2] -= acc[p][q];
        }
      if (!full) {
        double acg[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
        tc_tile_4x4<NB>(PT_G, PT_G, ldt, 4 * ti, 4 * tj, acg);
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = 4 * ti + p, i2 = 4 * tj + q;
            if (i < nr && i2 <= i) G[(size_t)(r0 + i) * ld + r0 + i2] -= acg[p][q];
          }
      }
    }
    cluster.sync();
    stamp();  // trailing done
  }
  cluster.sync();
  if (gtid == 0) *rank_out = s_rank;  // CTA 0 thread 0 (gtid 0) owns s_rank
  // dropped rows: clear what earlier panels wrote left of the diagonal
  for (size_t e = gtid; e < (size_t)n * n; e += gthreads) {
    const int k = (int)(e / n), cc = (int)(e % n);
    if (cc < k && !keep[k]) A[(size_t)k * ld + cc] = 0.0;
  }
  cluster.sync();
  stamp();  // factorisation complete
  // ---------------------------------------------------------------- W = L^-1 [TP | r''], RHS columns sharded over the cluster
  {
    const int ncol = n + 1;
    // columns per chunk: an even share of the RHS, bounded by what fits in the (now free) panel buffers
    const int cwmax = max(1, min(kTailThreads, (int)((2 * (size_t)NB * ldt) / (size_t)n)));
    const int per = min(cwmax, (ncol + C - 1) / C);
    const int nchunk = (ncol + per - 1) / per;
    double* Ws = PT_A;                       // [n][per]
    double* Dblk = DG;
    int* skeep = reinterpret_cast<int*>(d0); // reuse
    for (int k = tid; k < n; k += kTailThreads) skeep[k] = keep[k];
    for (int q = crank; q < nchunk; q += C) {
      const int col0 = q * per, cw = min(per, ncol - col0);
      __syncthreads();
      for (int e = tid; e < n * cw; e += kTailThreads) {
        const int row = e / cw, cc = e % cw, col = col0 + cc;
        double v = 0.0;
        if (skeep[row]) v = (col < n) ? TP[(size_t)row * ld + col] : r2[row];
        Ws[(size_t)row * per + cc] = v;
      }
      for (int kb = 0; kb < n; kb += NB) {
        const int nb = min(NB, n - kb), r0 = kb + nb;
        __syncthreads();
        tc_load_diag<NB>(Dblk, A, ld, kb, nb, tid);
        __syncthreads();
        if (tid < cw) {
          double x[NB];
#pragma unroll
          for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? Ws[(size_t)(kb + j) * per + tid] : 0.0;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            double v = x[j];
#pragma unroll
            for (int cc = 0; cc < j; ++cc) v -= Dblk[j * LD + cc] * x[cc];
            x[j] = ((j < nb) && skeep[kb + j]) ? v / Dblk[j * LD + j] : 0.0;
          }
#pragma unroll
          for (int j = 0; j < NB; ++j)
            if (j < nb) Ws[(size_t)(kb + j) * per + tid] = x[j];
        }
        __syncthreads();
        // rows below: Ws[r][c] -= sum_j L[r][kb+j] * Ws[kb+j][c]   (L rows from global/L2, broadcast within a warp)
        for (int e = tid; e < (n - r0) * cw; e += kTailThreads) {
          const int i = e / cw, cc = e % cw;
          const double* lrow = A + (size_t)(r0 + i) * ld + kb;
          double s = 0.0;
#pragma unroll 8
          for (int j = 0; j < nb; ++j) s += lrow[j] * Ws[(size_t)(kb + j) * per + cc];
          Ws[(size_t)(r0 + i) * per + cc] -= s;
        }
      }
      __syncthreads();
      for (int e = tid; e < n * cw; e += kTailThreads) {
        const int row = e / cw, cc = e % cw, col = col0 + cc;
        const double v = Ws[(size_t)row * per + cc];
        if (col < n) Wm[(size_t)row * ld + col] = v; else yv[row] = v;
      }
    }
  }
  cluster.sync();
  stamp();  // substitution complete
  // ---------------------------------------------------------------- P <- P - W^T W (4x4 register tiles straight from L2), dx = W^T y
  {
    const int nt = (n + 3) / 4, ntile = nt * (nt + 1) / 2;
    for (int tl = gtid; tl < ntile; tl += gthreads) {
      int ti = (int)((sqrtf(8.0f * (float)tl + 1.0f) - 1.0f) * 0.5f);
      while (ti * (ti + 1) / 2 > tl) --ti;
      while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
      const int tj = tl - ti * (ti + 1) / 2;
      double acc[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
      for (int k = 0; k < n; ++k) {
        const double* wr = Wm + (size_t)k * ld;
        double a[4], b[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) { a[p] = (4 * ti + p < n) ? wr[4 * ti + p] : 0.0; b[p] = (4 * tj + p < n) ? wr[4 * tj + p] : 0.0; }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] += a[p] * b[q];
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int a_ = 4 * ti + p, b_ = 4 * tj + q;
          if (a_ < n && b_ < n && b_ <= a_) {
            const S v = (S)((double)P[(size_t)a_ * ldp + b_] - acc[p][q]);
            P[(size_t)a_ * ldp + b_] = v;
            P[(size_t)b_ * ldp + a_] = v;  // exactly symmetric by construction
          }
        }
    }
  }
  stamp();  // syrk (CTA 0's share) done
  if (crank == 0) {  // dx = W^T y and the state injection (msckf.h:1373-1391)
    double* sdx = PT_G;  // [n]
    __shared__ double part[16][33];
    const int al = tid & 31, kg = tid >> 5;  // 32 columns x 16 k-groups
    for (int a0 = 0; a0 < n; a0 += 32) {
      const int a = a0 + al;
      double s = 0.0;
      if (a < n)
        for (int k = kg; k < n; k += 16) s += Wm[(size_t)k * ld + a] * yv[k];
      part[kg][al] = s;
      __syncthreads();
      if (kg == 0 && a < n) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 16; ++g) t += part[g][al];
        sdx[a] = t;
        dx_out[a] = t;
      }
      __syncthreads();
    }
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
      st->last_m = m;
      st->last_rank = s_rank;
      double nn = 0.0;
      for (int a = 0; a < n; ++a) nn += sdx[a] * sdx[a];
      st->last_dx_norm = sqrt(nn);
    }
    for (int ci = tid; ci < M; ci += kTailThreads) {
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
  stamp();  // end
  if (prof && blockIdx.x == 0 && threadIdx.x == 0 && prof_i < 64) prof[prof_i] = 0ull;
}

}  // namespace mb
