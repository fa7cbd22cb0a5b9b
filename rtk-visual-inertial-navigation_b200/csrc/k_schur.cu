// K3 + K5: Schur elimination of the first ordering group and back-substitution, one CTA per
// window.  Replaces SchurEliminator<-1,-1,-1>::Eliminate / BackSubstitute
// (CERES/internal/ceres/schur_eliminator_impl.h:177-306, 309-375) for the predefined ordering of
// RVI/swf/swf_gnss.cpp:629-783 (landmarks 3x3, every second speed-bias 9x9, epoch clocks 1x1).
//
// Formulation (no atomics, deterministic -- the reference serialises S updates with per-cell
// mutexes, :552):
//   phase 1, per chunk c (rows sharing e-block e):   L L' = D_e^2 + sum E'E      (chunk factor)
//            W_f = L^-1 sum E'F_f  for every f-block of the chunk,  w_g = L^-1 sum E'b
//   phase 2, gather: S_pq = [p==q] D_p^2 + sum_rows F_p'F_q - sum_chunks W_p'W_q ,
//            rhs_p = sum_rows F_p'b - sum_chunks W_p'w_g   (extra column of the diagonal cells);
//            the 8x8 output tiles of all touched block cells are dealt to the warps by the planner
//            (plan.cpp, "gather streams"), every term is one FP64 tensor-core MMA and each element of S
//            is stored exactly once.
//   back-substitution: y_e = L^-T (w_g - sum_f W_f z_f).
// (E'E + D^2)^-1 of the reference (InvertPSDMatrix, invert_psd_matrix.h:62-67: LLT-solve-identity)
// is applied in factored form: buffer' inv buffer = (L^-1 buffer)'(L^-1 buffer).
#include <cstdlib>
#include <mutex>

#include "dev_common.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

constexpr int kThreads = 256;            // k_backsub
constexpr int kWarps = kThreads / 32;
// k_schur: one CTA of SCHUR_WARPS warps per window (the gather streams are dealt to exactly these warps).
// A thread-block cluster per window (8 CTAs, barrier.cluster between the phases) was measured and
// dropped: the barrier imbalance cost more than the parallelism gained.
constexpr int kSchurThreads = 256;
constexpr int kSchurWarps = kSchurThreads / 32;

// D(8x8) += A(8x4, row) * B(4x8, col); lane holds A[lane>>2][lane&3], B[lane&3][lane>>2],
// D[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// explicit global-space load (the window base pointer is laundered through an asm barrier below,
// which would otherwise demote the loads to generic LD)
__device__ __forceinline__ double ld_global(const double* p) {
  double v;
  asm("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// Bulk L2 prefetch (TMA engine, no registers, no completion tracking): pulls a byte range into L2
// ahead of the demand loads.  J and the residuals were written by k_eval for the whole batch and
// have long left L2 when k_schur starts, so every first touch would otherwise pay DRAM latency
// inside a chain of dependent index -> operand loads.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_range(const void* base, size_t bytes, int tid, int nthreads) {
  constexpr size_t kChunk = 4096;
  const char* p = reinterpret_cast<const char*>(base);
  const size_t aligned = bytes & ~size_t(15);
  for (size_t off = (size_t)tid * kChunk; off < aligned; off += (size_t)nthreads * kChunk)
    prefetch_l2_bulk(p + off, (unsigned)((aligned - off) < kChunk ? (aligned - off) : kChunk));
}

// ---- phase 2: per-warp gather streams with a descriptor ring in shared memory -----------------------
// Every warp owns one linear stream of stages (plan.cpp, "gather stream": a 16-byte header plus SCHUR_STAGE
// 16-byte term descriptors of one output tile) and consumes it stage by stage:
//   D(s): header + 8 descriptors of stage s, global -> shared ring by asynchronous copies (cp.async, one
//         16-byte copy per lane < 9, completion tracked per commit group), issued three stages ahead of use
//   O(s): the 16 operand fragments of stage s, predicated global loads, straight-line
//   C(s): 8 FP64 tensor-core MMAs into two alternating accumulator pairs; the tile is stored when the header
//         says it is complete (each element of S is written exactly once)
// The descriptors are already in shared memory when a stage starts, so the dependent descriptor -> operand ->
// MMA chain of the gather costs ONE memory round trip per stage instead of two.  Measured alternatives: 8-byte
// cp.async copies of the operands into shared rings (LDGSTS issues far below LDG rate at this granularity: 1.5x
// slower than the direct gather), and operand loads one stage ahead in registers (kRegPipeline below).
constexpr int kStageRecs = SCHUR_STAGE + 1;  // header + terms
#ifndef SWGN_REG_PIPELINE
#define SWGN_REG_PIPELINE 0
#endif
#ifndef SWGN_SCHUR_CTAS
#define SWGN_SCHUR_CTAS 3
#endif
// operand loads one stage ahead in registers: 128 registers -> 2 CTAs/SM, measured 4.65 ms per 2048-window launch
// against 3.56 ms for loads issued in the consuming stage at 80 registers / 3 CTAs/SM (and 4.31 ms before the ring)
constexpr bool kRegPipeline = SWGN_REG_PIPELINE != 0;
struct GatherRings {
  int4 dq[4][kStageRecs + 3];  // descriptor ring: stages k .. k+3 (padded to 12 records)
};
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// lane constants of one output tile, derived from the meta word of the stage header
struct TileLane {
  int meta, a_lo, b_lo;
  bool a_ok, b_any, b_rhs;
};
__device__ __forceinline__ void tile_lane(int meta, int la, int lb, TileLane& T) {
  T.meta = meta;
  const int ps = meta & 63, qs = (meta >> 6) & 63, ti = ((meta >> 12) & 7) * 8, tj = ((meta >> 15) & 7) * 8;
  const int diag = (meta >> 18) & 1;
  const int ai = ti + lb, bj = tj + lb;
  T.a_ok = ai < ps;
  const bool b_ok = bj < qs;
  T.b_rhs = diag && bj == qs;
  T.a_lo = T.a_ok ? la * ps + ai : 0;
  T.b_lo = T.b_rhs ? la : (b_ok ? la * qs + bj : 0);
  T.b_any = b_ok || T.b_rhs;
}

// operand fragments of one stage: plain global loads into registers, issued a full stage ahead of use;
// straight-line (padding entries and out-of-block lanes are predicated off and contribute zeros)
__device__ __forceinline__ void load_stage(const double* JW, const int4* dq, int la, int lb, TileLane& T, double (&av)[SCHUR_STAGE],
                                           double (&bv)[SCHUR_STAGE]) {
  const int meta = dq[0].w;
  if (meta != T.meta) tile_lane(meta, la, lb, T);  // warp-uniform, once per tile
#pragma unroll
  for (int e = 0; e < SCHUR_STAGE; ++e) {
    const int4 t = dq[1 + e];
    const bool ok = t.x >= 0 && la <= ((t.x >> 28) & 3);
    const double a = (ok && T.a_ok) ? ld_global(JW + (T.a_lo + (t.x & 0x0fffffff))) : 0.0;
    av[e] = (t.x & (1 << 30)) ? -a : a;
    bv[e] = (ok && T.b_any) ? ld_global(JW + (T.b_lo + (T.b_rhs ? t.z : t.y))) : 0.0;
  }
}

// ECELL = false: tiles of the reduced system -> S (out0), row stride ld, rhs in column nf, + D^2 on the diagonal;
// ECELL = true: raw chunk products -> W_EFAC (out0, [E'E] of diagonal cells) / W_EBUF (out1, E'F and E'b)
template <bool ECELL>
__device__ void gather_stream(const double* JW, const int4* gs, int n_stage, GatherRings& R, int lane, double* out0, double* out1, int ld,
                              int nf, const double* lmd_f /* LM diagonal of the f-blocks */) {
  const int la = lane & 3, lb = lane >> 2;
  TileLane T;
  T.meta = -1;
  T.a_lo = T.b_lo = 0;
  T.a_ok = T.b_any = T.b_rhs = false;
  auto issue_d = [&](int s) {
    if (s < n_stage && lane < kStageRecs) cp_async16(&R.dq[s & 3][lane], gs + (size_t)s * kStageRecs + lane);
    cp_async_commit();
  };
  if (n_stage <= 0) return;
  issue_d(0);
  issue_d(1);
  issue_d(2);
  cp_async_wait<2>();  // D(0)
  __syncwarp();
  double av[SCHUR_STAGE], bv[SCHUR_STAGE];
  if (kRegPipeline) load_stage(JW, R.dq[0], la, lb, T, av, bv);
  double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
  for (int k = 0; k < n_stage; ++k) {
    cp_async_wait<1>();  // D(k+1) has landed (D(k+2) may still be in flight)
    __syncwarp();        // ... for every lane; also: all lanes are done with ring slot (k+3)&3 = (k-1)&3
    issue_d(k + 3);
    double an[SCHUR_STAGE], bn[SCHUR_STAGE];
    if (kRegPipeline) {
      if (k + 1 < n_stage) load_stage(JW, R.dq[(k + 1) & 3], la, lb, T, an, bn);
    } else {
      load_stage(JW, R.dq[k & 3], la, lb, T, av, bv);
    }
    const int4 hdr = R.dq[k & 3][0];
#pragma unroll
    for (int e = 0; e < SCHUR_STAGE; e += 2) {
      dmma884(c0, c1, av[e], bv[e]);
      dmma884(d0, d1, av[e + 1], bv[e + 1]);
    }
    if (hdr.z & 1) {  // the tile is complete: store it (each element of S is written exactly once)
      const int meta = hdr.w;
      const int ps = meta & 63, qs = (meta >> 6) & 63, ti = ((meta >> 12) & 7) * 8, tj = ((meta >> 15) & 7) * 8;
      const int diag = (meta >> 18) & 1;
      const int soff = hdr.x, frow = hdr.y;
      const int i = ti + lb;
      if (i < ps) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = tj + 2 * la + h;
          double val = h ? (c1 + d1) : (c0 + d0);
          if (ECELL) {
            if (j < qs) (diag ? out0 : out1)[soff + i * qs + j] = val;
            else if (diag && j == qs) out1[frow + i] = val;
          } else if (j < qs) {
            if (diag && i == j) {  // + D^2  (schur_eliminator_impl.h:194-215)
              const double dd = lmd_f[frow + i];
              val += dd * dd;
            }
            out0[soff + (size_t)i * ld + j] = val;
          } else if (diag && j == qs) {
            out0[(size_t)(frow + i) * ld + nf] = val;
          }
        }
      }
      c0 = c1 = d0 = d1 = 0.0;
    }
    if (kRegPipeline) {
#pragma unroll
      for (int e = 0; e < SCHUR_STAGE; ++e) {
        av[e] = an[e];
        bv[e] = bn[e];
      }
    }
  }
  cp_async_wait<0>();
}

// ---- phase 1, small e-blocks (1..3): one thread per chunk, everything in registers -----------
template <int ES>
__device__ __forceinline__ void chunk_thread(const Win& v, int chunk, const double* lmd) {
  const int32_t* chunk_row = v.I(I_CHUNK_ROW);
  const int32_t* row_cell = v.I(I_ROW_CELL);
  const int32_t* row_res = v.I(I_ROW_RES);
  const int32_t* row_nres = v.I(I_ROW_NRES);
  const int32_t* cell_col = v.I(I_CELL_COL);
  const int32_t* cell_val = v.I(I_CELL_VAL);
  const int32_t* cell_slot = v.I(I_CELL_SLOT);
  const int32_t* cell_first = v.I(I_CELL_FIRST);
  const int32_t* col_size = v.I(I_COL_SIZE);
  const double* J = v.W(W_JAC);
  const double* R = v.W(W_RES);
  double* EB = v.W(W_EBUF);
  const int ecol = v.I(I_CHUNK_ECOL)[chunk];
  const int epos = v.I(I_COL_POS)[ecol];
  const int r0 = chunk_row[chunk], r1 = chunk_row[chunk + 1];
  double ete[ES][ES], g[ES];
#pragma unroll
  for (int i = 0; i < ES; ++i) {
    g[i] = 0.0;
#pragma unroll
    for (int j = 0; j < ES; ++j) ete[i][j] = 0.0;
    const double dd = lmd ? lmd[epos + i] : 0.0;
    ete[i][i] = dd * dd;
  }
  for (int r = r0; r < r1; ++r) {  // ChunkDiagonalBlockAndGradient :444-507
    const double* E = J + cell_val[row_cell[r]];
    const double* bb = R + row_res[r];
    const int nres = row_nres[r];
    for (int rr = 0; rr < nres; ++rr) {
      double e[ES];
#pragma unroll
      for (int i = 0; i < ES; ++i) e[i] = E[rr * ES + i];
      const double br = bb[rr];
#pragma unroll
      for (int i = 0; i < ES; ++i) {
        g[i] += e[i] * br;
#pragma unroll
        for (int j = i; j < ES; ++j) ete[i][j] += e[i] * e[j];
      }
    }
  }
  // L L' = ete (lower L, row-major); a non-positive pivot poisons the chunk with NaN so that the
  // reduced factorisation fails and the caller retries with a larger mu (dogleg_strategy.cc:589)
  double L[ES][ES];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < ES; ++j) {
    double dj = ete[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k];
    if (!(dj > 0.0)) ok = false;
    dj = sqrt(dj);
    L[j][j] = dj;
#pragma unroll
    for (int i = j + 1; i < ES; ++i) {
      double s = ete[j][i];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = s / dj;
    }
  }
  if (!ok) {
#pragma unroll
    for (int i = 0; i < ES; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) L[i][j] = nan("");
  }
  double* fac = v.W(W_EFAC) + v.I(I_CHUNK_FAC)[chunk];
#pragma unroll
  for (int i = 0; i < ES; ++i)
#pragma unroll
    for (int j = 0; j < ES; ++j) fac[i * ES + j] = (j <= i) ? L[i][j] : 0.0;
  {  // w_g = L^-1 g
    double wg[ES];
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      double s = g[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= L[i][k] * wg[k];
      wg[i] = s / L[i][i];
    }
    double* gp = EB + v.I(I_CHUNK_G)[chunk];
#pragma unroll
    for (int i = 0; i < ES; ++i) gp[i] = wg[i];
  }
  if (v.I(I_CHUNK_SIMPLE)[chunk]) return;  // the W blocks of simple chunks are built row-parallel (phase 1b)
  // W_f (+)= (L^-1 E_rr') F_rr for every residual row rr of every row block
  for (int r = r0; r < r1; ++r) {
    const int c0 = row_cell[r], c1 = row_cell[r + 1];
    if (c1 - c0 < 2) continue;
    const double* E = J + cell_val[c0];
    const int nres = row_nres[r];
    for (int rr = 0; rr < nres; ++rr) {
      double vv[ES];
#pragma unroll
      for (int i = 0; i < ES; ++i) {
        double s = E[rr * ES + i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i][k] * vv[k];
        vv[i] = s / L[i][i];
      }
      for (int c = c0 + 1; c < c1; ++c) {
        const int fs = col_size[cell_col[c]];
        const double* F = J + cell_val[c] + rr * fs;
        double* Wf = EB + cell_slot[c];
        const bool store = cell_first[c] && rr == 0;
        for (int j = 0; j < fs; ++j) {
          const double f = F[j];
#pragma unroll
          for (int i = 0; i < ES; ++i) {
            const double t = vv[i] * f;
            Wf[i * fs + j] = store ? t : Wf[i * fs + j] + t;
          }
        }
      }
    }
  }
}

// ---- phase 1a, small e-blocks whose slots are fed by several rows: one warp per chunk, lanes over the rows ----
// (planner: <= 32 rows, <= 2 residuals and <= 4 f-cells per row, <= 32 slots).  Same arithmetic as chunk_thread, but
// the rows' index -> Jacobian load chains run side by side instead of one after the other, and the slot sums
// are warp reductions instead of read-modify-write chains through W_EBUF.
template <int ES>
__device__ void chunk_small_warp(const Win& v, int chunk, const double* lmd) {
  const int lane = threadIdx.x & 31;
  const int32_t* row_cell = v.I(I_ROW_CELL);
  const int32_t* cell_col = v.I(I_CELL_COL);
  const int32_t* cell_val = v.I(I_CELL_VAL);
  const int32_t* cell_slot = v.I(I_CELL_SLOT);
  const int32_t* col_size = v.I(I_COL_SIZE);
  const double* J = v.W(W_JAC);
  const double* R = v.W(W_RES);
  double* EB = v.W(W_EBUF);
  const int ecol = v.I(I_CHUNK_ECOL)[chunk];
  const int epos = v.I(I_COL_POS)[ecol];
  const int r0 = v.I(I_CHUNK_ROW)[chunk], r1 = v.I(I_CHUNK_ROW)[chunk + 1];
  const int s0 = v.I(I_CHUNK_SLOT)[chunk], ns = v.I(I_CHUNK_SLOT)[chunk + 1] - s0;
  // slot table, one slot per lane
  int my_fs = 0, my_buf = -1;
  if (lane < ns) {
    my_fs = col_size[v.I(I_SLOT_COL)[s0 + lane]];
    my_buf = v.I(I_SLOT_BUF)[s0 + lane];
  }
  // my row
  const int r = r0 + lane;
  const bool have = r < r1;
  int nres = 0, nfc = 0, fval[4] = {0, 0, 0, 0}, fslot[4] = {-1, -1, -1, -1};
  double e[2][ES], br[2] = {0.0, 0.0};
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int i = 0; i < ES; ++i) e[q][i] = 0.0;
  if (have) {
    const int c0 = row_cell[r], c1 = row_cell[r + 1];
    nres = v.I(I_ROW_NRES)[r];
    nfc = c1 - c0 - 1;
    const double* E = J + cell_val[c0];
    const double* bb = R + v.I(I_ROW_RES)[r];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < nfc) {
        fval[q] = cell_val[c0 + 1 + q];
        fslot[q] = cell_slot[c0 + 1 + q];
      }
#pragma unroll
    for (int q = 0; q < 2; ++q)
      if (q < nres) {
#pragma unroll
        for (int i = 0; i < ES; ++i) e[q][i] = E[q * ES + i];
        br[q] = bb[q];
      }
  }
  double ete[ES][ES], g[ES];
#pragma unroll
  for (int i = 0; i < ES; ++i) {
    g[i] = warp_sum(e[0][i] * br[0] + e[1][i] * br[1]);
#pragma unroll
    for (int j = i; j < ES; ++j) ete[i][j] = warp_sum(e[0][i] * e[0][j] + e[1][i] * e[1][j]);
    const double dd = lmd ? lmd[epos + i] : 0.0;
    ete[i][i] += dd * dd;
  }
  double L[ES][ES];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < ES; ++j) {
    double dj = ete[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k];
    if (!(dj > 0.0)) ok = false;
    dj = sqrt(dj);
    L[j][j] = dj;
#pragma unroll
    for (int i = j + 1; i < ES; ++i) {
      double s = ete[j][i];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = s / dj;
    }
  }
  if (!ok) {
#pragma unroll
    for (int i = 0; i < ES; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) L[i][j] = nan("");
  }
  if (lane == 0) {
    double* fac = v.W(W_EFAC) + v.I(I_CHUNK_FAC)[chunk];
#pragma unroll
    for (int i = 0; i < ES; ++i)
#pragma unroll
      for (int j = 0; j < ES; ++j) fac[i * ES + j] = (j <= i) ? L[i][j] : 0.0;
    double wg[ES];
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      double s = g[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= L[i][k] * wg[k];
      wg[i] = s / L[i][i];
    }
    double* gp = EB + v.I(I_CHUNK_G)[chunk];
#pragma unroll
    for (int i = 0; i < ES; ++i) gp[i] = wg[i];
  }
  // vv_q = L^-1 E_q' for my residual rows
  double vv[2][ES];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      double s = e[q][i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= L[i][k] * vv[q][k];
      vv[q][i] = s / L[i][i];
    }
  // W_f = sum over the rows feeding slot f of vv' F_f: one warp reduction per element of the slot block
  for (int s = 0; s < ns; ++s) {
    const int fs = __shfl_sync(0xffffffffu, my_fs, s), sb = __shfl_sync(0xffffffffu, my_buf, s);
    int mine = -1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (fslot[q] == sb) mine = fval[q];
    for (int j = 0; j < fs; ++j) {
      double f0 = 0.0, f1 = 0.0;
      if (mine >= 0) {
        f0 = J[mine + j];
        if (nres > 1) f1 = J[mine + fs + j];
      }
#pragma unroll
      for (int i = 0; i < ES; ++i) {
        const double w = warp_sum(vv[0][i] * f0 + vv[1][i] * f1);
        if (lane == 0) EB[sb + i * fs + j] = w;
      }
    }
  }
}

// ---- phase 1b: one thread per row of a simple chunk: W_f = (L^-1 E') F_f, stored once ---------
// rec = the row's self-contained record (I_SROW): every data load below depends on that one index load only
template <int ES>
__device__ __forceinline__ void row_w(const Win& v, const int4 r0, const int4 r1) {
  const double* J = v.W(W_JAC);
  double* EB = v.W(W_EBUF);
  const double* Lp = v.W(W_EFAC) + r0.z;
  const double* E = J + r0.x;
  const int nres = r0.y & 0xff, n_fcells = r0.y >> 16;
  double L[ES][ES];
#pragma unroll
  for (int i = 0; i < ES; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) L[i][k] = Lp[i * ES + k];
  const int32_t* cell_col = v.I(I_CELL_COL);
  const int32_t* cell_val = v.I(I_CELL_VAL);
  const int32_t* cell_slot = v.I(I_CELL_SLOT);
  const int32_t* col_size = v.I(I_COL_SIZE);
  if (nres == 2) {  // the visual case: both residual rows at once, every W entry written once
    double v0[ES], v1[ES];
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      double s0 = E[i], s1 = E[ES + i];
#pragma unroll
      for (int k = 0; k < i; ++k) {
        s0 -= L[i][k] * v0[k];
        s1 -= L[i][k] * v1[k];
      }
      v0[i] = s0 / L[i][i];
      v1[i] = s1 / L[i][i];
    }
    for (int q = 0; q < n_fcells; ++q) {
      const int c = r0.w + q;
      const int fs = q == 0 ? r1.z : col_size[cell_col[c]];
      const double* F = J + (q == 0 ? r1.x : cell_val[c]);
      double* Wf = EB + (q == 0 ? r1.y : cell_slot[c]);
      for (int j = 0; j < fs; ++j) {
        const double f0 = F[j], f1 = F[fs + j];
#pragma unroll
        for (int i = 0; i < ES; ++i) Wf[i * fs + j] = v0[i] * f0 + v1[i] * f1;
      }
    }
    return;
  }
  for (int rr = 0; rr < nres; ++rr) {
    double vv[ES];
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      double s = E[rr * ES + i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= L[i][k] * vv[k];
      vv[i] = s / L[i][i];
    }
    for (int q = 0; q < n_fcells; ++q) {
      const int c = r0.w + q;
      const int fs = q == 0 ? r1.z : col_size[cell_col[c]];
      const double* F = J + (q == 0 ? r1.x : cell_val[c]) + rr * fs;
      double* Wf = EB + (q == 0 ? r1.y : cell_slot[c]);
      for (int j = 0; j < fs; ++j) {
        const double f = F[j];
#pragma unroll
        for (int i = 0; i < ES; ++i) Wf[i * fs + j] = (rr == 0) ? vv[i] * f : Wf[i * fs + j] + vv[i] * f;
      }
    }
  }
}

// ---- phase 1, larger e-blocks (4..16, the 9-dim speed-bias blocks): one warp per chunk --------
__device__ void chunk_warp(const Win& v, int chunk, const double* lmd, double* sm /* per-warp scratch */) {
  const int lane = threadIdx.x & 31;
  const int32_t* chunk_row = v.I(I_CHUNK_ROW);
  const int32_t* chunk_slot = v.I(I_CHUNK_SLOT);
  const int32_t* slot_col = v.I(I_SLOT_COL);
  const int32_t* slot_buf = v.I(I_SLOT_BUF);
  const int32_t* row_cell = v.I(I_ROW_CELL);
  const int32_t* row_res = v.I(I_ROW_RES);
  const int32_t* row_nres = v.I(I_ROW_NRES);
  const int32_t* cell_col = v.I(I_CELL_COL);
  const int32_t* cell_val = v.I(I_CELL_VAL);
  const int32_t* cell_slot = v.I(I_CELL_SLOT);
  const int32_t* col_size = v.I(I_COL_SIZE);
  const double* J = v.W(W_JAC);
  const double* R = v.W(W_RES);
  double* EB = v.W(W_EBUF);
  const int ecol = v.I(I_CHUNK_ECOL)[chunk];
  const int es = col_size[ecol];
  const int epos = v.I(I_COL_POS)[ecol];
  const int s0 = chunk_slot[chunk], s1 = chunk_slot[chunk + 1];
  const int ebase = (s1 > s0) ? slot_buf[s0] : v.I(I_CHUNK_G)[chunk];
  const int gofs = v.I(I_CHUNK_G)[chunk];
  // scratch: ete/L [16][16] then the buffer in the global slot layout (slot blocks es x fs, then g)
  double* ete = sm;
  double* buf = sm + MAX_WARP_E * MAX_WARP_E;
  const int nbuf = gofs + es - ebase;
  for (int k = lane; k < es * es; k += 32) {
    const int i = k / es, j = k - i * es;
    const double dd = (lmd && i == j) ? lmd[epos + i] : 0.0;
    ete[i * MAX_WARP_E + j] = dd * dd;
  }
  __syncwarp();
  // raw products [E'E | E'b] (W_EFAC / g slot) and E'F_f (slot blocks) were gathered by the tensor-core
  // pass of phase 1a (e-cells); bring them into the warp's scratch and add D^2
  {
    const double* rawfac = v.W(W_EFAC) + v.I(I_CHUNK_FAC)[chunk];
    for (int k = lane; k < es * es; k += 32) {
      const int i = k / es, j = k - i * es;
      if (j >= i) ete[i * MAX_WARP_E + j] += rawfac[k];
    }
    for (int k = lane; k < nbuf; k += 32) buf[k] = EB[ebase + k];
    __syncwarp();
  }
  // Cholesky of ete (upper part filled) -> lower L in place (row-major [i][j], j <= i)
  bool ok = true;
  for (int j = 0; j < es; ++j) {
    double dj = ete[j * MAX_WARP_E + j];
    for (int k = 0; k < j; ++k) dj -= ete[j * MAX_WARP_E + k] * ete[j * MAX_WARP_E + k];
    if (!(dj > 0.0)) ok = false;
    dj = sqrt(dj);
    __syncwarp();
    if (lane == 0) ete[j * MAX_WARP_E + j] = dj;
    const int i = j + 1 + lane;
    if (i < es) {
      double s = ete[j * MAX_WARP_E + i];
      for (int k = 0; k < j; ++k) s -= ete[i * MAX_WARP_E + k] * ete[j * MAX_WARP_E + k];
      ete[i * MAX_WARP_E + j] = s / dj;
    }
    __syncwarp();
  }
  if (!ok) {
    for (int k = lane; k < es * es; k += 32) ete[(k / es) * MAX_WARP_E + (k % es)] = nan("");
    __syncwarp();
  }
  double* fac = v.W(W_EFAC) + v.I(I_CHUNK_FAC)[chunk];
  for (int k = lane; k < es * es; k += 32) {
    const int i = k / es, j = k - i * es;
    fac[k] = (j <= i) ? ete[i * MAX_WARP_E + j] : 0.0;
  }
  // forward substitution L^-1 on every column of every slot block and on g.  The columns of all slots are
  // flattened over the lanes (a 9-dim e-block has ~40 columns in 6..8 slots: two passes instead of one
  // pass per slot with 6..9 busy lanes); the slot table is loaded once, one slot per lane
  const int ns1 = s1 - s0 + 1;  // slots + the g column
  if (ns1 <= 32) {
    int my_fs = 0, my_off = 0;
    if (lane < ns1) {
      const int s = s0 + lane;
      my_fs = (s < s1) ? col_size[slot_col[s]] : 1;
      my_off = ((s < s1) ? slot_buf[s] : gofs) - ebase;
    }
    int pre = my_fs;  // inclusive prefix sum of the slot widths
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, ns1 - 1);
    for (int c = lane; c < ((total + 31) & ~31); c += 32) {
      int fs = 1, off = 0, j = -1;
      for (int s = 0; s < ns1; ++s) {
        const int hi = __shfl_sync(0xffffffffu, pre, s), f = __shfl_sync(0xffffffffu, my_fs, s), o = __shfl_sync(0xffffffffu, my_off, s);
        if (c < total && c >= hi - f && c < hi) { fs = f; off = o; j = c - (hi - f); }
      }
      if (j >= 0) {
        for (int i = 0; i < es; ++i) {
          double t = buf[off + i * fs + j];
          for (int k = 0; k < i; ++k) t -= ete[i * MAX_WARP_E + k] * buf[off + k * fs + j];
          buf[off + i * fs + j] = t / ete[i * MAX_WARP_E + i];
        }
      }
    }
  } else {
    for (int s = s0; s <= s1; ++s) {
      const int fs = (s < s1) ? col_size[slot_col[s]] : 1;
      const int off = ((s < s1) ? slot_buf[s] : gofs) - ebase;
      for (int j = lane; j < fs; j += 32) {
        for (int i = 0; i < es; ++i) {
          double t = buf[off + i * fs + j];
          for (int k = 0; k < i; ++k) t -= ete[i * MAX_WARP_E + k] * buf[off + k * fs + j];
          buf[off + i * fs + j] = t / ete[i * MAX_WARP_E + i];
        }
      }
    }
  }
  __syncwarp();
  for (int k = lane; k < nbuf; k += 32) EB[ebase + k] = buf[k];
  __syncwarp();
}

__device__ __forceinline__ void chunk_dispatch(const Win& v, int chunk, const double* lmd) {
  const int es = v.I(I_COL_SIZE)[v.I(I_CHUNK_ECOL)[chunk]];
  if (es == 3) chunk_thread<3>(v, chunk, lmd);
  else if (es == 1) chunk_thread<1>(v, chunk, lmd);
  else chunk_thread<2>(v, chunk, lmd);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSchurThreads, SWGN_SCHUR_CTAS) k_schur(DeviceBatch b, int only_window) {
  __shared__ WinDesc sd;
  extern __shared__ __align__(16) double dyn[];  // phase 1: kSchurWarps * max_wbuf doubles; phase 2: the gather rings
  const int w = only_window >= 0 ? only_window : blockIdx.x;
  TRState* st = b.state + w;
  if (only_window < 0 && !(st->active && st->need_solve)) return;
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  if (d.sb_ok) return;  // this window runs the streamed kernel (k_schur_stream.cu)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gtid = tid, gwid = wid;
  const double* lmd = v.W(W_LMD);
  double* S = v.W(W_S);
  const int nf = d.n_f, ld = d.ld;
  long long* dbg = b.debug ? b.debug + 8 * (size_t)w : nullptr;
#define SWGN_STAMP(i) do { if (dbg && gtid == 0) dbg[i] = clock64(); } while (0)
  SWGN_STAMP(0);
  // start the DRAM -> L2 transfer of everything this CTA will read: Jacobian + residuals, the
  // index tables of the window (contiguous in ipool), the LM diagonal
  prefetch_range(v.W(W_JAC), sizeof(double) * (size_t)d.n_jac, gtid, kSchurThreads);
  prefetch_range(v.W(W_RES), sizeof(double) * (size_t)d.n_res, gtid, kSchurThreads);
  prefetch_range(v.I(I_ROW_RES), sizeof(int32_t) * (size_t)(d.ioff[I_PROJ] - d.ioff[I_ROW_RES]), gtid, kSchurThreads);
  prefetch_range(lmd, sizeof(double) * (size_t)d.n_t, gtid, kSchurThreads);

  // phase 0: clear the upper triangle and the rhs column (cells never touched must read as 0).  Measured: leaving the
  // clear out shortens a 4 096-window launch from 6.8 to 5.8 ms; handing the same bytes to the TMA engine as asynchronous
  // shared -> global copies of a zeroed line (no LSU instructions, overlapped with phase 1) gains nothing (6.8 ms): the cost
  // is the 0.48 MB per window of write traffic competing with phase 1, not instruction issue.  What would help is writing
  // fewer bytes: 27 % of the block cells are live (stored in full by phase 2) and k_chol never reads the column groups its
  // symbolic masks mark dead -- a planner clear-list of the remainder is the open item.
  for (int i = gwid; i < nf; i += kSchurWarps) {
    double* row = S + (size_t)i * ld;
    for (int j = i + lane; j <= nf; j += 32) row[j] = 0.0;
  }
  SWGN_STAMP(1);
  // phase 1a: factors of the small chunks (one thread each; W buffers too unless row-parallel) and,
  // on the tensor pipe, the raw products E'[E | b | F] of the larger e-blocks (one warp per e-cell)
  {
    const int32_t* tcw = v.I(I_TCHUNK_W);
    for (int k = gwid; k < d.n_tchunks_w; k += kSchurWarps) {
      const int chunk = tcw[k];
      const int es = v.I(I_COL_SIZE)[v.I(I_CHUNK_ECOL)[chunk]];
      if (es == 1) chunk_small_warp<1>(v, chunk, lmd);
      else if (es == 3) chunk_small_warp<3>(v, chunk, lmd);
      else chunk_small_warp<2>(v, chunk, lmd);
    }
    const int32_t* tch = v.I(I_TCHUNK_T);
    for (int k = gtid; k < d.n_tchunks_t; k += kSchurThreads) chunk_dispatch(v, tch[k], lmd);
    const int32_t* eptr = v.I(I_ESTREAM_PTR);
    const int4* gs = reinterpret_cast<const int4*>(v.I(I_ESTREAM)) + (size_t)eptr[wid] * kStageRecs;
    const double* JWc = v.W(W_JAC);
    asm volatile("" : "+l"(JWc));
    GatherRings* rings = reinterpret_cast<GatherRings*>(dyn);
    gather_stream<true>(JWc, gs, eptr[wid + 1] - eptr[wid], rings[wid], lane, v.W(W_EFAC), v.W(W_EBUF), 0, 0, nullptr);
  }
  SWGN_STAMP(2);
  __syncthreads();
  SWGN_STAMP(3);
  // phase 1b: factor + forward substitution of the larger e-blocks (one warp per chunk, shared-memory
  // scratch), and the W blocks of the simple chunks, one thread per row (consecutive threads read
  // consecutive row blocks of J)
  {
    const int32_t* wch = v.I(I_WCHUNK);
    double* sm = dyn + (size_t)wid * b.max_wbuf;
    for (int k = gwid; k < d.n_wchunks; k += kSchurWarps) chunk_warp(v, wch[k], lmd, sm);
  }
  {
    const int4* srow = reinterpret_cast<const int4*>(v.I(I_SROW));
    for (int k = gtid; k < d.n_srows; k += kSchurThreads) {
      const int4 r0 = srow[2 * k], r1 = srow[2 * k + 1];
      const int es = (r0.y >> 8) & 0xff;
      if (es == 3) row_w<3>(v, r0, r1);
      else if (es == 1) row_w<1>(v, r0, r1);
      else row_w<2>(v, r0, r1);
    }
  }
  __syncthreads();
  SWGN_STAMP(4);
  // phase 2: one warp per block cell.  S_pq = sum_t (+/-) A_t' B_t is a skinny GEMM whose K
  // dimension is the stack of the gathered blocks; every term feeds one FP64 tensor-core MMA
  // (m8n8k4: A_t' is the 8x4 operand, B_t the 4x8 operand, rows beyond m masked to zero), lanes
  // load their fragment element straight from the window's JW region, so a warp touches one or
  // two cache lines per operand.  Diagonal cells carry the rhs as column qs of the B operand.
  // Terms come in runs of equal shape (plan.cpp) and are processed four at a time: eight operand
  // loads in flight per lane before the four MMAs.
  {
    const int32_t* wptr = v.I(I_WSTREAM_PTR);
    const int4* gs = reinterpret_cast<const int4*>(v.I(I_WSTREAM)) + (size_t)wptr[wid] * kStageRecs;
    const int n_stage = wptr[wid + 1] - wptr[wid];
    const double* JW = v.W(W_JAC);
    asm volatile("" : "+l"(JW));
    GatherRings* rings = reinterpret_cast<GatherRings*>(dyn);  // phase-1 scratch is dead after the barrier
    gather_stream<false>(JW, gs, n_stage, rings[wid], lane, S, nullptr, ld, nf, lmd + d.n_e);
  }
  SWGN_STAMP(5);
  if (dbg) {
    __syncthreads();
    SWGN_STAMP(6);
    if (gtid == 0) {
      unsigned smid;
      asm("mov.u32 %0, %smid;" : "=r"(smid));
      dbg[7] = smid;
    }
  }
  if (b.keep_copy) {
    __syncthreads();
    double* SC = v.W(W_SCOPY);
    for (int k = gtid; k < nf * ld; k += kSchurThreads) SC[k] = S[k];
  }
  if (gtid == 0) {
    st->num_linear_solves += 1;
    st->have_factor = 0;
    st->have_reduced = b.params.export_mode ? 1 : 0;
    st->chol_ok = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Back-substitution + Gauss-Newton step in the scaled space: gn = -diag .* [y_e ; z]
// (dogleg_strategy.cc:606).  Also closes the linear-solve attempt of this tick: a failed
// factorisation or a non-finite step multiplies mu by 10 and leaves need_solve set (:589-595).
// ---------------------------------------------------------------------------------------------
template <int ES>
__device__ __forceinline__ void backsub_thread(const Win& v, int chunk, const double* Y, double* Yout) {
  const int32_t* chunk_slot = v.I(I_CHUNK_SLOT);
  const int32_t* slot_col = v.I(I_SLOT_COL);
  const int32_t* slot_buf = v.I(I_SLOT_BUF);
  const int32_t* col_size = v.I(I_COL_SIZE);
  const int32_t* col_pos = v.I(I_COL_POS);
  const double* EB = v.W(W_EBUF);
  const double* L = v.W(W_EFAC) + v.I(I_CHUNK_FAC)[chunk];
  const double* wg = EB + v.I(I_CHUNK_G)[chunk];
  double t[ES];
#pragma unroll
  for (int i = 0; i < ES; ++i) t[i] = wg[i];
  for (int s = chunk_slot[chunk]; s < chunk_slot[chunk + 1]; ++s) {
    const int fc = slot_col[s], fs = col_size[fc];
    const double* z = Y + col_pos[fc];
    const double* Wf = EB + slot_buf[s];
    for (int j = 0; j < fs; ++j) {
      const double zj = z[j];
#pragma unroll
      for (int i = 0; i < ES; ++i) t[i] -= Wf[i * fs + j] * zj;
    }
  }
  double y[ES];
#pragma unroll
  for (int i = ES - 1; i >= 0; --i) {
    double s = t[i];
#pragma unroll
    for (int k = i + 1; k < ES; ++k) s -= L[k * ES + i] * y[k];
    y[i] = s / L[i * ES + i];
  }
  double* out = Yout + col_pos[v.I(I_CHUNK_ECOL)[chunk]];
#pragma unroll
  for (int i = 0; i < ES; ++i) out[i] = y[i];
}

__device__ void backsub_warp(const Win& v, int chunk, const double* Y, double* Yout) {
  const int lane = threadIdx.x & 31;
  const int32_t* chunk_slot = v.I(I_CHUNK_SLOT);
  const int32_t* slot_col = v.I(I_SLOT_COL);
  const int32_t* slot_buf = v.I(I_SLOT_BUF);
  const int32_t* col_size = v.I(I_COL_SIZE);
  const int32_t* col_pos = v.I(I_COL_POS);
  const double* EB = v.W(W_EBUF);
  const int ecol = v.I(I_CHUNK_ECOL)[chunk];
  const int es = col_size[ecol];
  const double* L = v.W(W_EFAC) + v.I(I_CHUNK_FAC)[chunk];
  double t = 0.0;
  if (lane < es) {
    t = EB[v.I(I_CHUNK_G)[chunk] + lane];
    for (int s = chunk_slot[chunk]; s < chunk_slot[chunk + 1]; ++s) {
      const int fc = slot_col[s], fs = col_size[fc];
      const double* z = Y + col_pos[fc];
      const double* Wf = EB + slot_buf[s] + lane * fs;
      for (int j = 0; j < fs; ++j) t -= Wf[j] * z[j];
    }
  }
  // L' y = t, bottom up
  double y = 0.0;
  for (int i = es - 1; i >= 0; --i) {
    double yi = 0.0;
    if (lane == i) yi = t / L[i * es + i];
    yi = __shfl_sync(0xffffffffu, yi, i);
    if (lane == i) y = yi;
    if (lane < i) t -= L[i * es + lane] * yi;
  }
  if (lane < es) Yout[col_pos[ecol] + lane] = y;
}

__global__ void __launch_bounds__(kThreads) k_backsub(DeviceBatch b, int only_window) {
  __shared__ WinDesc sd;
  const int w = only_window >= 0 ? only_window : blockIdx.x;
  TRState* st = b.state + w;
  if (only_window < 0 && !(st->active && st->need_solve)) return;
  const int tid = threadIdx.x, wid = tid >> 5;
  const bool lm = b.params.strategy == SWGN_LEVENBERG_MARQUARDT;
  if (!b.params.export_mode && !st->chol_ok) {  // LINEAR_SOLVER_FAILURE
    if (tid == 0) {
      if (lm) {  // no retry loop in LevenbergMarquardtStrategy::ComputeStep: the step is invalid (k_step shrinks the radius)
        st->need_solve = 0;
        st->solve_ok = 0;
      } else {
        st->mu *= 10.0;  // dogleg: retry with mu * 10
      }
    }
    return;
  }
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  double* Y = v.W(W_Y);
  double* GN = v.W(W_GN);
  const double* DG = v.W(W_DIAG);
  int bad = 0;
  if (b.params.export_mode) {
    // schur_complement_solver.cc:172-188: the reduced system is exported and the solve returns
    // success with a zero step
    for (int k = tid; k < d.n_t; k += kThreads) { Y[k] = 0.0; GN[k] = 0.0; }
  } else {
    const int32_t* tch = v.I(I_TCHUNK);
    for (int k = tid; k < d.n_tchunks; k += kThreads) {
      const int chunk = tch[k];
      const int es = v.I(I_COL_SIZE)[v.I(I_CHUNK_ECOL)[chunk]];
      if (es == 3) backsub_thread<3>(v, chunk, Y, Y);
      else if (es == 1) backsub_thread<1>(v, chunk, Y, Y);
      else backsub_thread<2>(v, chunk, Y, Y);
    }
    const int32_t* wch = v.I(I_WCHUNK);
    for (int k = wid; k < d.n_wchunks; k += kWarps) backsub_warp(v, wch[k], Y, Y);
    __syncthreads();
    for (int k = tid; k < d.n_t; k += kThreads) {
      const double y = Y[k];
      if (!finite_d(y)) bad = 1;
      GN[k] = -DG[k] * y;
    }
  }
  bad = block_any(bad);
  if (tid == 0 && only_window < 0) {
    if (bad && lm) {
      st->need_solve = 0;
      st->solve_ok = 0;
    } else if (bad) {
      st->mu *= 10.0;  // stays in need_solve: retried in the next tick
    } else {
      st->need_solve = 0;
      st->solve_ok = 1;
    }
  }
}

static size_t schur_dyn_bytes(const DeviceBatch& b) {
  size_t dyn = sizeof(double) * (size_t)kSchurWarps * (size_t)(b.max_wbuf > 0 ? b.max_wbuf : 1);
  dyn = dyn > sizeof(GatherRings) * kSchurWarps ? dyn : sizeof(GatherRings) * kSchurWarps;
  // experiment switch: a floor on the dynamic shared memory caps the CTAs resident per SM (>= 76 KB: 2, >= 114 KB: 1),
  // i.e. the W / EFAC working set competing for the 126 MB L2
  static const size_t floor_bytes = []() {
    const char* e = std::getenv("SWGN_SCHUR_SMEM_FLOOR");
    return e ? (size_t)std::atol(e) : (size_t)0;
  }();
  return dyn > floor_bytes ? dyn : floor_bytes;
}

void launch_schur_gather(const DeviceBatch& b, int only_window, cudaStream_t s) {
  const int grid = only_window >= 0 ? 1 : b.n_windows;
  const size_t dyn = schur_dyn_bytes(b);
  k_schur<<<grid, kSchurThreads, dyn, s>>>(b, only_window);
}
// every window is run by exactly one of the two kernels (WinDesc::sb_ok); a kernel is only launched when the batch
// holds windows of its kind
void launch_schur(const DeviceBatch& b, int only_window, cudaStream_t s) {
  if (b.sb_windows > 0) launch_schur_stream(b, only_window, s);
  if (b.gather_windows > 0) launch_schur_gather(b, only_window, s);
}
void launch_backsub(const DeviceBatch& b, int only_window, cudaStream_t s) {
  const int grid = only_window >= 0 ? 1 : b.n_windows;
  k_backsub<<<grid, kThreads, 0, s>>>(b, only_window);
}

cudaError_t configure_schur(const DeviceBatch& b) {
  static std::mutex mu;  // batches are created from several host threads
  static size_t granted[64] = {0};
  std::lock_guard<std::mutex> lk(mu);
  static_assert(kSchurWarps == SCHUR_WARPS, "the gather streams are dealt to the warps of one CTA");
  const size_t dyn = schur_dyn_bytes(b);
  if (dyn > 227 * 1024) return cudaErrorInvalidValue;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dyn > 48 * 1024 && (dev < 0 || dev >= 64 || dyn > granted[dev])) {
    cudaError_t e = cudaFuncSetAttribute(k_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) granted[dev] = dyn;
  }
  return cudaSuccess;
}

}  // namespace swgn
