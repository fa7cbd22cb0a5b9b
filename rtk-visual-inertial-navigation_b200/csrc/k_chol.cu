// K4: dense Cholesky of the reduced pose / speed-bias / GNSS-state system and the reduced solve,
// one CTA per window.  Replaces DenseSchurComplementSolver::SolveReducedLinearSystem
// (CERES/internal/ceres/schur_complement_solver.cc:223-268: Eigen LLT<Upper> + solve).
//
// S (n_f x n_f, row-major upper triangle, leading dimension ld) carries the rhs as column n_f, so
// the forward substitution U^T w = rhs is the same right-looking update as the factorisation.
// This is the one dense contraction of the path, and the only place tensor cores are used: the
// FP64 tensor-core instruction mma.sync.m8n8k4.f64 (DMMA) carries both the in-panel updates and
// the trailing update.  Blocked right-looking algorithm, NB-row panel staged in shared memory:
//   panel <- S[k0:k0+NB, k0:]                              (HBM/L2 -> shared)
//   for every 8-row block r of the panel:
//       8x8 Cholesky of the diagonal tile                  (one warp)
//       X = U_rr^-T P[r, :]   by substitution               (one thread per column)
//       P[s, :] -= U[r, s]^T X  for the row blocks s > r    (DMMA, 8x8 tiles in shared memory)
//   S[k0:k0+NB, k0:] <- panel                              (U rows are final)
//   S[i, j] -= sum_p U[p,i] U[p,j]  for i, j >= k0+NB       (DMMA, 32x32 tiles per warp, operands
//                                                           from the shared panel, C read-modify-
//                                                           written in HBM/L2 exactly once per panel)
// then the backward solve U z = w in 32-row blocks.  The factor stays in W_S: it is the
// `lhs_out2 = llt.matrixL()` export (schur_complement_solver.cc:253-258) transposed.
#include <cstdlib>
#include <mutex>

#include "dev_common.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

#ifndef SWGN_CHOL_CTAS
#define SWGN_CHOL_CTAS 2  // 3 (80 registers, spills) with 16-row panels measured 23 % slower, with 32-row panels 11 % slower
#endif
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// D(8x8) += A(8x4, row) * B(4x8, col); lane holds A[lane>>2][lane&3], B[lane&3][lane>>2],
// D[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ---- TMA (bulk async copy) staging of the panel rows: global -> shared, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n"
      "DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kThreads, SWGN_CHOL_CTAS) k_chol(DeviceBatch b, int only_window, int NB, int pw) {
  __shared__ WinDesc sd;
  __shared__ int s_fail;
  __shared__ double s_rdiag[2][8];
  __shared__ double s_diag[32][33];
  __shared__ __align__(8) uint64_t s_bar;
  extern __shared__ __align__(16) double dyn[];  // panel NB x pw, then wv, zv, rdg [max_nf each]
  const int w = only_window >= 0 ? only_window : blockIdx.x;
  TRState* st = b.state + w;
  if (only_window < 0 && !(st->active && st->need_solve)) return;
  if (b.params.export_mode) return;  // the reduced system is exported, not solved
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int la = lane & 3, lb = lane >> 2;
  const int nf = d.n_f, ld = d.ld, ncol = nf + 1;
  double* S = v.W(W_S);
  double* P = dyn;
  double* wv = dyn + (size_t)NB * pw;
  double* zv = wv + b.max_nf;
  double* rdg = zv + b.max_nf;  // 1 / U[i][i], filled as the pivots are computed
  if (tid == 0) {
    s_fail = 0;
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned bar_phase = 0;
  long long* dbg = b.debug ? b.debug + 8 * (size_t)(b.n_windows + w) : nullptr;  // second half: k_schur owns the first
  long long t_prev = dbg ? clock64() : 0, acc_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_begin = t_prev;
#define CHOL_STAMP(i) do { if (dbg && tid == 0) { const long long t_ = clock64(); acc_t[i] += t_ - t_prev; t_prev = t_; } } while (0)

  const int32_t* cmask = v.I(I_CHOL_MASK);
  for (int k0 = 0; k0 < nf; k0 += NB) {
    const int nb = min(NB, nf - k0);
    // 16-column groups (absolute columns 16 g ..) that can be non-zero in this panel's rows: everything else is
    // an exact zero that stays zero, so its TRSM columns, in-panel tiles and trailing-update tiles are skipped
    unsigned long long pmask = ~0ull;
    if (NB == 32 && nb == NB) pmask = (unsigned long long)(unsigned)cmask[2 * (k0 >> 5)] | ((unsigned long long)(unsigned)cmask[2 * (k0 >> 5) + 1] << 32);
    auto live16 = [&](int abs_col) { return (pmask >> (abs_col >> 4)) & 1ull; };
    const int Wm = nf - k0;              // matrix columns of the panel (panel col j = global col k0 + j)
    // panel column of the rhs: right after the matrix columns, except in a final partial panel,
    // where the identity padding of rows nb..NB-1 occupies columns nb..NB-1 and the rhs moves to NB
    const int jr = (nb < NB) ? NB : Wm;
    const int Wp = (jr + 1 + 7) & ~7;    // columns processed inside the panel
    const int Wz = min(pw - 4, max(Wp, (Wm + 1 + 31 + (NB == 16 ? 16 : 0)) & ~31));  // zero-filled extent (trailing tiles read up to here)
    // ---- stage the panel.  Full panels: one bulk async copy (TMA) per row, S[k0+r, k0:ld] ->
    // P[r, 0:ld-k0], issued by one thread and awaited on an mbarrier; entries left of the diagonal
    // come along (finite, never read), columns beyond ld-k0 are zero-filled by the other threads.
    // The final partial panel needs identity padding and the relocated rhs: staged by hand.
    if (nb == NB) {
      const int ncopy = ld - k0;  // doubles per row, a multiple of 4
      if (tid == 0) {
        mbar_expect_tx(&s_bar, (unsigned)(NB * ncopy * sizeof(double)));
        for (int r = 0; r < NB; ++r) bulk_g2s(P + (size_t)r * pw, S + (size_t)(k0 + r) * ld + k0, (unsigned)(ncopy * sizeof(double)), &s_bar);
      }
      for (int e = tid; e < NB * (Wz - ncopy); e += kThreads) {
        const int r = e / (Wz - ncopy), j = ncopy + e % (Wz - ncopy);
        P[(size_t)r * pw + j] = 0.0;
      }
      mbar_wait(&s_bar, bar_phase);
      bar_phase ^= 1;
    } else {
      for (int r = wid; r < NB; r += kWarps) {
        const double* src = S + (size_t)(k0 + r) * ld;
        double* dst = P + (size_t)r * pw;
        for (int j0 = lane; j0 < Wz; j0 += 128) {  // four independent loads in flight per lane
          double val[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + 32 * u;
            double t = 0.0;
            if (r < nb) {
              if (j >= r && j < Wm) t = src[k0 + j];
              else if (j == jr) t = src[nf];
            } else if (j == r) {
              t = 1.0;
            }
            val[u] = t;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j0 + 32 * u < Wz) dst[j0 + 32 * u] = val[u];
        }
      }
    }
    __syncthreads();
    CHOL_STAMP(0);
    for (int r8 = 0; r8 * 8 < NB; ++r8) {
      const int r0 = r8 * 8;
      // (1) 8x8 Cholesky of the diagonal tile (upper, in place), reciprocal pivots to s_rdiag
      if (wid == 0) {
        // lane j < 8 keeps column j of the tile in registers; pivots travel by shuffle.  1/sqrt
        // through rsqrt keeps the serial pivot chain short (346 pivots per window are a latency
        // floor of the whole factorisation).
        double* T = P + (size_t)r0 * pw + r0;
        const int j = lane & 7;
        double t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = T[(size_t)i * pw + j];
        bool bad = false;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const double piv = __shfl_sync(0xffffffffu, t[p], p);
          if (!(piv > 0.0)) bad = true;  // Eigen LLT: NumericalIssue
          const double ri = rsqrt(piv);
          if (lane == p) {
            s_rdiag[0][p] = ri;
            if (k0 + r0 + p < nf) rdg[k0 + r0 + p] = ri;
          }
          t[p] = (j == p) ? piv * ri : t[p] * ri;
#pragma unroll
          for (int q = p + 1; q < 8; ++q) {
            const double c = __shfl_sync(0xffffffffu, t[p], q);
            if (j >= q) t[q] -= c * t[p];
          }
        }
        __syncwarp();  // lanes 8..31 read the same tile above (duplicates of columns 0..7)
        if (lane < 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (j >= i) T[(size_t)i * pw + j] = t[i];
        }
        if (bad && lane == 0) s_fail = 1;
      }
      __syncthreads();
      CHOL_STAMP(5);
      // (2) X = U_rr^-T P[r0:r0+8, j] for every column right of the tile, one thread per column
      {
        const double* U = P + (size_t)r0 * pw + r0;
        for (int j = r0 + 8 + tid; j < Wp; j += kThreads) {
          if (j >= NB && j < Wm && !live16(k0 + j)) continue;  // structurally zero column of the panel
          double* col = P + (size_t)r0 * pw + j;
          double x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            double s = col[(size_t)i * pw];
#pragma unroll
            for (int p = 0; p < i; ++p) s -= U[(size_t)p * pw + i] * x[p];
            x[i] = s * s_rdiag[0][i];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) col[(size_t)i * pw] = x[i];
        }
      }
      __syncthreads();
      CHOL_STAMP(6);
      // (3) DMMA update of the panel's remaining row blocks: P[s0.., j0..] -= U[r0.., s0..]^T X[r0.., j0..]
      {
        const int nsb = NB / 8, ntj = Wp / 8;
        int idx = wid;
        for (int s8 = r8 + 1; s8 < nsb; ++s8) {
          const int len = ntj - s8;
          while (idx < len) {
            const int s0 = s8 * 8, j0 = (s8 + idx) * 8;
            if (j0 >= NB && j0 + 8 <= Wm && !live16(k0 + j0)) {
              idx += kWarps;
              continue;
            }
            double2* cp = reinterpret_cast<double2*>(P + (size_t)(s0 + lb) * pw + j0 + 2 * la);
            double2 c = *cp;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const double* row = P + (size_t)(r0 + 4 * kk + la) * pw;
              dmma884(c.x, c.y, -row[s0 + lb], row[j0 + lb]);
            }
            *cp = c;
            idx += kWarps;
          }
          idx -= len;
        }
      }
      __syncthreads();
      CHOL_STAMP(7);
    }
    CHOL_STAMP(1);
    if (s_fail) break;
    // ---- write the finished U rows back
    for (int r = wid; r < nb; r += kWarps) {
      double* dst = S + (size_t)(k0 + r) * ld;
      const double* src = P + (size_t)r * pw;
      for (int j = r + lane; j < Wm; j += 32) dst[k0 + j] = src[j];
      if (lane == 0) dst[nf] = src[jr];
    }
    CHOL_STAMP(2);
    // ---- trailing update with DMMA: 16x32 warp tiles of the block upper triangle of S[t0:, t0:],
    // k = NB.  The C tile of the NEXT work item is loaded (HBM/L2) while the current one runs on
    // the tensor pipe; the accumulators start from C and A is negated: D = C - U^T U.
    const int t0 = k0 + NB;
    if (t0 < nf) {
      const int th = nf - t0, tw = ncol - t0;
      const int TI = (th + 15) >> 4, TJ = (tw + 31) >> 5;
      // work items: (ti, tj) with 32*tj + 31 >= 16*ti, enumerated row by row
      auto first_tj = [](int ti) { return ti >> 1; };
      int ti = 0, idx = wid;
      auto advance = [&](int& ti_, int& idx_) {  // normalise (ti, idx) to a valid item or ti == TI
        while (ti_ < TI && idx_ >= TJ - first_tj(ti_)) {
          idx_ -= TJ - first_tj(ti_);
          ++ti_;
        }
      };
      auto load_tile = [&](int ti_, int tj_, double (&c)[2][4][2]) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const int gi = t0 + 16 * ti_ + 8 * mi + lb;
          const double* row = S + (size_t)gi * ld;
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int gj = t0 + 32 * tj_ + 8 * ni + 2 * la;
            const bool v0 = gi < nf && gj >= gi && gj <= nf, v1 = gi < nf && gj + 1 >= gi && gj + 1 <= nf;
            if (v0 && v1) {
              const double2 t = *reinterpret_cast<const double2*>(row + gj);
              c[mi][ni][0] = t.x;
              c[mi][ni][1] = t.y;
            } else {
              c[mi][ni][0] = v0 ? row[gj] : 0.0;
              c[mi][ni][1] = v1 ? row[gj + 1] : 0.0;
            }
          }
        }
      };
      // a tile (ti, tj) only changes when both of its column groups are live in the panel
      auto tile_live = [&](int ti_, int tj_) {
        const int ci = t0 + 16 * ti_, cj = t0 + 32 * tj_;
        return live16(ci) && (live16(cj) || live16(cj + 16));
      };
      auto next_live = [&](int& ti_, int& idx_) {
        for (;;) {
          advance(ti_, idx_);
          if (ti_ >= TI || tile_live(ti_, first_tj(ti_) + idx_)) return;
          idx_ += kWarps;
        }
      };
      next_live(ti, idx);
      double cn[2][4][2];
      if (ti < TI) load_tile(ti, first_tj(ti) + idx, cn);
      while (ti < TI) {
        const int tj = first_tj(ti) + idx;
        double acc[2][4][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            acc[mi][ni][0] = cn[mi][ni][0];
            acc[mi][ni][1] = cn[mi][ni][1];
          }
        int nti = ti, nidx = idx + kWarps;
        next_live(nti, nidx);
        if (nti < TI) load_tile(nti, first_tj(nti) + nidx, cn);
        const double* pa = P + NB + 16 * ti + lb;
        const double* pb = P + NB + 32 * tj + lb;
#pragma unroll 2
        for (int kk = 0; kk < NB / 4; ++kk) {
          const size_t ro = (size_t)(4 * kk + la) * pw;
          double a[2], bb[4];
#pragma unroll
          for (int q = 0; q < 2; ++q) a[q] = -pa[ro + 8 * q];
#pragma unroll
          for (int q = 0; q < 4; ++q) bb[q] = pb[ro + 8 * q];
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          const int gi = t0 + 16 * ti + 8 * mi + lb;
          double* row = S + (size_t)gi * ld;
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) {
            const int gj = t0 + 32 * tj + 8 * ni + 2 * la;
            const bool v0 = gi < nf && gj >= gi && gj <= nf, v1 = gi < nf && gj + 1 >= gi && gj + 1 <= nf;
            if (v0 && v1) *reinterpret_cast<double2*>(row + gj) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            else if (v0) row[gj] = acc[mi][ni][0];
            else if (v1) row[gj + 1] = acc[mi][ni][1];
          }
        }
        ti = nti;
        idx = nidx;
      }
    }
    __syncthreads();
    CHOL_STAMP(3);
  }

  const int fail = s_fail;
  if (!fail) {
    // ---- backward solve U z = w (w = column n_f), 32-row blocks from the bottom, row oriented:
    // t_i = w_i - U[i, k0+32:] z[k0+32:] for the 32 rows of the block (one warp per four rows,
    // lanes stride the columns: coalesced row segments, four rows in flight), then the 32x32
    // triangular solve by one warp with the stored reciprocal pivots
    for (int k0 = ((nf - 1) / 32) * 32; k0 >= 0; k0 -= 32) {
      const int nb = min(32, nf - k0);
      const int c0 = k0 + nb;  // first column right of the block
      for (int e = tid; e < nb * nb; e += kThreads) {
        const int r = e / nb, c = e - r * nb;
        s_diag[r][c] = (c >= r) ? S[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
      }
      {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const int rbase = 4 * wid;  // rows rbase..rbase+3 of the block
        const double* rowp[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) rowp[q] = S + (size_t)(k0 + min(rbase + q, nb - 1)) * ld;  // clamped: extra rows are discarded below
        int j = c0 + lane;
        for (; j + 96 < nf; j += 128) {  // sixteen independent loads in flight per lane
          double u[4][4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int q = 0; q < 4; ++q) u[t][q] = rowp[q][j + 32 * t];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const double zj = zv[j + 32 * t];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += u[t][q] * zj;
          }
        }
        for (; j < nf; j += 32) {
          const double zj = zv[j];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] += rowp[q][j] * zj;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double s = warp_sum(acc[q]);
          if (lane == 0 && rbase + q < nb) wv[rbase + q] = S[(size_t)(k0 + rbase + q) * ld + nf] - s;
        }
      }
      __syncthreads();
      if (wid == 0) {
        double wl = lane < nb ? wv[lane] : 0.0;
        for (int i = nb - 1; i >= 0; --i) {
          double zi = 0.0;
          if (lane == i) zi = wl * rdg[k0 + i];
          zi = __shfl_sync(0xffffffffu, zi, i);
          if (lane == i) wl = zi;
          if (lane < i) wl -= s_diag[lane][i] * zi;
        }
        if (lane < nb) zv[k0 + lane] = wl;
      }
      __syncthreads();
    }
    double* Y = v.W(W_Y) + d.n_e;
    for (int i = tid; i < nf; i += kThreads) Y[i] = zv[i];
  }
  CHOL_STAMP(4);
  if (dbg && tid == 0) {
    for (int i = 0; i < 5; ++i) dbg[i] = acc_t[i];
    dbg[1] = acc_t[5] + acc_t[6] + acc_t[7] + acc_t[1];
    dbg[5] = clock64() - t_begin;
    dbg[6] = acc_t[5];
    dbg[7] = acc_t[6];
  }
  if (tid == 0) {
    st->chol_ok = fail ? 0 : 1;
    st->have_factor = fail ? 0 : 1;
  }
}

// UpdateSchurHessianOnly (RVI/swf/swf_gnss.cpp:65-94): A = L_nn L_nn^T of the trailing n_tail rows,
// with L = U^T:  A_ij = sum_k U[m+k][m+i] U[m+k][m+j], k <= min(i,j)
__global__ void k_tail_information(DeviceBatch b, int window, int n_tail, double* A) {
  const WinDesc& d = b.desc[window];
  const double* S = b.wpool + d.woff[W_S];
  const int nf = d.n_f, ld = d.ld, m = nf - n_tail;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_tail * n_tail; e += gridDim.x * blockDim.x) {
    const int i = e / n_tail, j = e - i * n_tail;
    double s = 0.0;
    for (int k = 0; k < n_tail; ++k) {
      const double li = (k <= i) ? S[(size_t)(m + k) * ld + m + i] : 0.0;
      const double lj = (k <= j) ? S[(size_t)(m + k) * ld + m + j] : 0.0;
      s += li * lj;
    }
    A[e] = s;
  }
}

// ... for every window of the batch at once (one CTA per window), together with y = the current values of the n_tail
// scalar blocks at the end of the ordering (the float ambiguities) and a flag whether the window holds a factor
__global__ void k_tail_information_batch(DeviceBatch b, int n_tail, double* A_all, double* y_all, int32_t* have_A) {
  const int w = blockIdx.x;
  const WinDesc& d = b.desc[w];
  const TRState* st = b.state + w;
  const int ok = st->have_factor && d.n_f >= n_tail && d.n_cols - d.n_ecols >= n_tail;
  if (threadIdx.x == 0) have_A[w] = ok;
  if (!ok) return;
  const double* S = b.wpool + d.woff[W_S];
  const int nf = d.n_f, ld = d.ld, m = nf - n_tail;
  double* A = A_all + (size_t)w * n_tail * n_tail;
  for (int e = threadIdx.x; e < n_tail * n_tail; e += blockDim.x) {
    const int i = e / n_tail, j = e - i * n_tail;
    double s = 0.0;
    for (int k = 0; k < n_tail; ++k) {
      const double li = (k <= i) ? S[(size_t)(m + k) * ld + m + i] : 0.0;
      const double lj = (k <= j) ? S[(size_t)(m + k) * ld + m + j] : 0.0;
      s += li * lj;
    }
    A[e] = s;
  }
  const int32_t* col_state = b.ipool + d.ioff[I_COL_STATE];
  const double* x = b.wpool + d.woff[W_X];
  for (int k = threadIdx.x; k < n_tail; k += blockDim.x) y_all[(size_t)w * n_tail + k] = x[col_state[d.n_cols - n_tail + k]];
}

int chol_block(const DeviceBatch& b) {
  static const int forced = []() {  // experiment switch
    const char* e = std::getenv("SWGN_CHOL_NB");
    return e ? std::atoi(e) : 0;
  }();
  if (forced == 16 || forced == 32) return b.max_nf <= 760 ? forced : 16;
  return b.max_nf <= 760 ? 32 : 16;
}
int chol_pitch(const DeviceBatch& b) {
  const int NB = chol_block(b);
  const int need = (b.max_nf + 1 + (NB == 16 ? 16 : 0) + 31) & ~31;
  return (need > NB + 8 ? need : NB + 8) + 4;  // = 4 or 12 mod 16: conflict-free DMMA fragment loads
}
size_t chol_smem(const DeviceBatch& b) {
  return sizeof(double) * ((size_t)chol_block(b) * chol_pitch(b) + 3 * (size_t)b.max_nf);
}

}  // namespace

void launch_chol(const DeviceBatch& b, int only_window, cudaStream_t s) {
  const int grid = only_window >= 0 ? 1 : b.n_windows;
  k_chol<<<grid, kThreads, chol_smem(b), s>>>(b, only_window, chol_block(b), chol_pitch(b));
}

void launch_tail_information_batch(const DeviceBatch& b, int n_tail, double* A_all, double* y_all, int32_t* have_A, cudaStream_t s) {
  k_tail_information_batch<<<b.n_windows, 128, 0, s>>>(b, n_tail, A_all, y_all, have_A);
}

void launch_tail_information(const DeviceBatch& b, int window, int n_tail, double* A_dev, cudaStream_t s) {
  k_tail_information<<<8, 256, 0, s>>>(b, window, n_tail, A_dev);
}

// The opt-in dynamic shared memory limit is per function and per device: batches with different
// reduced-system sizes share it, so it only ever grows.
cudaError_t configure_chol(const DeviceBatch& b) {
  static std::mutex mu;  // batches are created from several host threads
  static size_t granted[64] = {0};
  std::lock_guard<std::mutex> lk(mu);
  const size_t dyn = chol_smem(b);
  if (dyn > 227 * 1024) return cudaErrorInvalidValue;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || dyn > granted[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) granted[dev] = dyn;
  }
  return cudaSuccess;
}

}  // namespace swgn
