// K3, streamed: Schur elimination of the first ordering group with the window's Jacobian staged through shared
// memory by TMA and the reduced system accumulated on chip.  Replaces SchurEliminator<-1,-1,-1>::Eliminate
// (CERES/internal/ceres/schur_eliminator_impl.h:177-306) for the predefined ordering of RVI/swf/swf_gnss.cpp:629-783.
//
// One CTA of SB_WARPS warps per window, one CTA per SM.  The host planner (plan_stream.cpp) cuts the rows, in
// Jacobian order, into batches of whole chunks; a batch's J segment, residual segment and record sections are
// contiguous in HBM and arrive by three bulk copies (cp.async.bulk ... mbarrier::complete_tx) one batch ahead of their
// use.  Per batch:
//   A  chunk products.  Landmark-like chunks (e-size <= 3, every f-block fed by one row): one thread per chunk,
//      E'E + D^2, E'b, L L' and w_g = L^-1 E'b in registers.  All other chunks (speed-bias blocks, epoch clocks):
//      the raw products E'[E | b | F] as runs of FP64 tensor-core MMAs from shared memory.
//   B  W_f = L^-1 E'F_f: one thread per row for the landmark-like chunks, one warp per chunk (in-place Cholesky and
//      forward substitution in shared memory) for the others.  The batch's E-buffer / factor segments leave for HBM
//      by bulk stores (k_backsub reads them), overlapped with C.
//   C  S_pq += F_p'F_q (rows), S_pq -= W_p'W_q (chunks): runs of MMAs per 8x8 tile of a block cell, operands from
//      shared memory, accumulators = the COMPACT block cells of S, resident in shared memory for the whole window;
//      the planner deals the tiles of every batch to the warps, a tile is owned by one warp per batch: no atomics,
//      deterministic (the reference serialises with per-cell mutexes, :552).
// The Jacobian is read from HBM exactly once and S is written exactly once, rows at a time, zeros included (no clear
// pass); the only other traffic is the record stream (8 bytes per MMA term, read through per-warp cp.async rings).
#include "dev_common.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

constexpr int kThreads = SB_WARPS * 32;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
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
      "SB_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SB_DONE;\n\t"
      "bra SB_WAIT_LOOP;\n"
      "SB_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- run streams ---------------------------------------------------------------------------------------------
// A warp's stream is a sequence of 16-byte units in global memory (L2): a run header followed by n/2 units of two
// terms each.  Units arrive in a per-warp ring of 4 chunks x 8 units by cp.async, three chunks ahead of their use.
struct Ring {
  int4 u[4][8];
};
static_assert(sizeof(Ring) == SB_RING_BYTES, "ring size is part of the planner's shared-memory budget");

struct Reader {
  const int4* gs;
  Ring* R;
  int n_units, issued, ready, lane;
  __device__ __forceinline__ void issue() {
    const int u = issued * 8 + lane;
    if (lane < 8 && u < n_units) cp_async16(&R->u[issued & 3][lane], gs + u);
    cp_async_commit();
    ++issued;
  }
  __device__ __forceinline__ void start(const int4* g, int n, Ring* r, int ln) {
    gs = g;
    n_units = n;
    R = r;
    lane = ln;
    issued = ready = 0;
    __syncwarp();  // every lane is done with the ring contents of the previous stream
    issue();
    issue();
    issue();
  }
  // unit u (and everything before it) has landed
  __device__ __forceinline__ void ensure(int u) {
    while (u >= ready * 8) {
      cp_async_wait<2>();
      __syncwarp();  // ... for every lane; also: all lanes are past chunk ready - 1, whose slot the next issue reuses
      ++ready;
      issue();
    }
  }
  __device__ __forceinline__ int4 unit(int u) const { return R->u[(u >> 3) & 3][u & 7]; }
  __device__ __forceinline__ void finish() { cp_async_wait<0>(); }
};

struct TileLane {
  int a_lo, b_lo;
  bool a_ok, b_any, b_rhs;
};

// Runs of one phase of one batch.  OA = operand area, ACC = accumulators.  ecell runs (phase A) store the raw chunk
// products into the operand area; the others (phase C) accumulate into the compact block cells of S.
__device__ void run_stream(const int4* gs, int n_units, Ring* ring, int lane, double* OA, double* ACC) {
  if (n_units <= 0) return;
  const int la = lane & 3, lb = lane >> 2;
  Reader rd;
  rd.start(gs, n_units, ring, lane);
  int u = 0;
  while (u < n_units) {
    rd.ensure(u);
    const int4 h = rd.unit(u);
    ++u;
    const int n = h.z & 0xffff, first = (h.z >> 16) & 1, ecell = (h.z >> 17) & 1;
    const int meta = h.w;
    const int ps = meta & 63, qs = (meta >> 6) & 63, ti = ((meta >> 12) & 7) * 8, tj = ((meta >> 15) & 7) * 8;
    const int diag = (meta >> 18) & 1;
    TileLane T;
    {
      const int ai = ti + lb, bj = tj + lb;
      T.a_ok = ai < ps;
      const bool b_ok = bj < qs;
      T.b_rhs = diag && bj == qs;
      T.a_lo = T.a_ok ? la * ps + ai : 0;
      T.b_lo = T.b_rhs ? la : (b_ok ? la * qs + bj : 0);
      T.b_any = b_ok || T.b_rhs;
    }
    // my two accumulator elements: (i, j) and (i, j + 1)
    const int i = ti + lb, j = tj + 2 * la;
    double* p0 = nullptr;
    double* p1 = nullptr;
    if (i < ps) {
      if (ecell) {
        if (j < qs) p0 = OA + h.x + i * qs + j;
        else if (diag && j == qs) p0 = OA + h.y + i;
        if (j + 1 < qs) p1 = OA + h.x + i * qs + j + 1;
        else if (diag && j + 1 == qs) p1 = OA + h.y + i;
      } else {
        const int stride = qs + diag;
        if (j < stride) p0 = ACC + h.x + i * stride + j;
        if (j + 1 < stride) p1 = ACC + h.x + i * stride + j + 1;
      }
    }
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    if (!first) {
      if (p0) c0 = *p0;
      if (p1) c1 = *p1;
    }
    for (int t = 0; t < n; t += 4) {
      rd.ensure(u + 1);
      const int4 r0 = rd.unit(u), r1 = rd.unit(u + 1);
      u += 2;
      const int w0[4] = {r0.x, r0.z, r1.x, r1.z}, w1[4] = {r0.y, r0.w, r1.y, r1.w};
      double av[4], bv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = w1[e] >= 0 && la <= ((w1[e] >> 16) & 3);
        const double a = (ok && T.a_ok) ? OA[T.a_lo + (w0[e] & 0xffff)] : 0.0;
        av[e] = (w1[e] & (1 << 18)) ? -a : a;
        bv[e] = (ok && T.b_any) ? OA[T.b_lo + (T.b_rhs ? (w1[e] & 0xffff) : ((unsigned)w0[e] >> 16))] : 0.0;
      }
      dmma884(c0, c1, av[0], bv[0]);
      dmma884(d0, d1, av[1], bv[1]);
      dmma884(c0, c1, av[2], bv[2]);
      dmma884(d0, d1, av[3], bv[3]);
    }
    if (p0) *p0 = c0 + d0;
    if (p1) *p1 = c1 + d1;
  }
  rd.finish();
}

// ---- phase A, landmark-like chunks: one thread per chunk ---------------------------------------------------------
// L L' = D^2 + sum E'E (lower L, row-major), w_g = L^-1 sum E'b.  A non-positive pivot poisons the chunk with NaN so
// that the reduced factorisation fails and the caller retries with a larger mu (dogleg_strategy.cc:589).
template <int ES>
__device__ __forceinline__ void tchunk_factor(const int4 rec, const int32_t* pkg, double* OA, const double* lmd) {
  const int n_rows = rec.y & 0xffff;
  const int2* crow = reinterpret_cast<const int2*>(pkg + rec.x);
  double ete[ES][ES], g[ES];
#pragma unroll
  for (int i = 0; i < ES; ++i) {
    g[i] = 0.0;
#pragma unroll
    for (int j = 0; j < ES; ++j) ete[i][j] = 0.0;
    const double dd = lmd[rec.w + i];
    ete[i][i] = dd * dd;
  }
  for (int r = 0; r < n_rows; ++r) {  // ChunkDiagonalBlockAndGradient, schur_eliminator_impl.h:444-507
    const int2 cr = crow[r];
    const double* E = OA + (cr.x & 0xffff);
    const double* bb = OA + cr.y;
    const int nres = cr.x >> 16;
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
  double* fac = OA + (rec.z & 0xffff);
#pragma unroll
  for (int i = 0; i < ES; ++i)
#pragma unroll
    for (int j = 0; j < ES; ++j) fac[i * ES + j] = (j <= i) ? L[i][j] : 0.0;
  double* gp = OA + ((unsigned)rec.z >> 16);
  double wg[ES];
#pragma unroll
  for (int i = 0; i < ES; ++i) {
    double s = g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i][k] * wg[k];
    wg[i] = s / L[i][i];
    gp[i] = wg[i];
  }
}

// ---- phase B, landmark-like chunks: one thread per row, W_f = (L^-1 E') F_f, written once --------------------------
template <int ES>
__device__ __forceinline__ void trow_w(const int4 r0, const int4 r1, const int32_t* pkg, double* OA) {
  const double* E = OA + r0.x;
  const double* Lp = OA + r0.z;
  const int nres = r0.y & 0xff, n_fcells = r0.y >> 16;
  double L[ES][ES];
#pragma unroll
  for (int i = 0; i < ES; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) L[i][k] = Lp[i * ES + k];
  double v0[ES], v1[ES];  // L^-1 E' for the (<= 2) residual rows
#pragma unroll
  for (int i = 0; i < ES; ++i) {
    double s0 = E[i], s1 = nres > 1 ? E[ES + i] : 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) {
      s0 -= L[i][k] * v0[k];
      s1 -= L[i][k] * v1[k];
    }
    v0[i] = s0 / L[i][i];
    v1[i] = s1 / L[i][i];
  }
  const int4* extra = reinterpret_cast<const int4*>(pkg + r0.w);
  for (int q = 0; q < n_fcells; ++q) {
    const int4 fc = q == 0 ? r1 : extra[q - 1];
    const double* F = OA + fc.x;
    double* Wf = OA + fc.y;
    const int fs = fc.z;
    for (int j = 0; j < fs; ++j) {
      const double f0 = F[j], f1 = nres > 1 ? F[fs + j] : 0.0;
#pragma unroll
      for (int i = 0; i < ES; ++i) Wf[i * fs + j] = v0[i] * f0 + v1[i] * f1;
    }
  }
}

// ---- phase B, all other chunks: one warp per chunk, in place in the operand area -----------------------------------
// in: raw E'E (upper part) at fac, raw E'F_f in the slot blocks, raw E'b in the g slot.  out: lower L at fac (upper
// part zero), W_f = L^-1 E'F_f, w_g = L^-1 E'b.
__device__ void mchunk_warp(const int4 m0, const int4 m1, const int32_t* pkg, double* OA, const double* lmd) {
  const int lane = threadIdx.x & 31;
  const int es = m0.x, epos = m0.y, ns1 = m0.w;
  double* ete = OA + m0.z;  // es x es row-major
  const int2* slots = reinterpret_cast<const int2*>(pkg + m1.x);
  if (lane < es) {
    const double dd = lmd[epos + lane];
    ete[lane * es + lane] += dd * dd;
  }
  __syncwarp();
  bool ok = true;
  for (int j = 0; j < es; ++j) {
    double dj = ete[j * es + j];
    for (int k = 0; k < j; ++k) dj -= ete[j * es + k] * ete[j * es + k];
    if (!(dj > 0.0)) ok = false;
    dj = sqrt(dj);
    __syncwarp();
    if (lane == 0) ete[j * es + j] = dj;
    for (int i = j + 1 + lane; i < es; i += 32) {
      double s = ete[j * es + i];
      for (int k = 0; k < j; ++k) s -= ete[i * es + k] * ete[j * es + k];
      ete[i * es + j] = s / dj;
    }
    __syncwarp();
  }
  for (int k = lane; k < es * es; k += 32) {
    const int i = k / es, j = k - i * es;
    if (!ok) ete[k] = nan("");
    else if (j > i) ete[k] = 0.0;
  }
  __syncwarp();
  // forward substitution on every column of every slot block and on g, columns flattened over the lanes
  int total = 0;
  for (int s = 0; s < ns1; ++s) total += slots[s].y;
  for (int c = lane; c < total; c += 32) {
    int s = 0, c0 = 0;
    while (c >= c0 + slots[s].y) {
      c0 += slots[s].y;
      ++s;
    }
    const int2 sl = slots[s];
    const int fs = sl.y, j = c - c0;
    double* B = OA + sl.x;
    for (int i = 0; i < es; ++i) {
      double t = B[i * fs + j];
      for (int k = 0; k < i; ++k) t -= ete[i * es + k] * B[k * fs + j];
      B[i * fs + j] = t / ete[i * es + i];
    }
  }
  __syncwarp();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) k_schur_stream(DeviceBatch b, int only_window) {
  __shared__ WinDesc sd;
  extern __shared__ __align__(16) unsigned char dyn_raw[];
  const int w = only_window >= 0 ? only_window : blockIdx.x;
  TRState* st = b.state + w;
  if (only_window < 0 && !(st->active && st->need_solve)) return;
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  if (!d.sb_ok) return;  // this window runs the gather kernel
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double* lmd = v.W(W_LMD);
  const int nf = d.n_f, ld = d.ld;
  long long* dbg = b.debug ? b.debug + 8 * (size_t)w : nullptr;
#define SWGN_STAMP(i) do { if (dbg && tid == 0) dbg[i] = clock64(); } while (0)
  SWGN_STAMP(0);

  // ---- shared-memory layout: [mbarriers | rings | headers | sections x 2 | operand area | accumulators]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_raw);
  Ring* rings = reinterpret_cast<Ring*>(dyn_raw + 16);
  int32_t* hdr = reinterpret_cast<int32_t*>(dyn_raw + 16 + SB_WARPS * SB_RING_BYTES);
  const int seccap = (d.sb_reccap + 3) & ~3;
  int32_t* sec0 = hdr + SB_HDR_INTS * d.sb_nbatch;
  const int scap = d.sb_jcap + d.sb_rcap;
  double* OA = reinterpret_cast<double*>(sec0 + 2 * seccap);
  double* ACC = OA + 2 * scap + d.sb_ecap + d.sb_fcap;
  const int32_t* ghdr = v.I(I_SB_HDR);
  const int32_t* grec = v.I(I_SB_REC);
  for (int k = tid; k < SB_HDR_INTS * d.sb_nbatch; k += kThreads) hdr[k] = ghdr[k];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double* gJ = v.W(W_JAC);
  const double* gR = v.W(W_RES);
  double* gEB = v.W(W_EBUF);
  double* gEF = v.W(W_EFAC);
  auto load_batch = [&](int k) {  // thread 0
    const int32_t* h = hdr + SB_HDR_INTS * k;
    const int s = k & 1;
    const unsigned bj = 8u * (unsigned)h[3], br = 8u * (unsigned)h[5], bs = 4u * (unsigned)h[15];
    mbar_expect_tx(&bars[s], bj + br + bs);
    if (bj) bulk_g2s(OA + s * scap, gJ + h[2], bj, &bars[s]);
    if (br) bulk_g2s(OA + s * scap + d.sb_jcap, gR + h[4], br, &bars[s]);
    if (bs) bulk_g2s(sec0 + s * seccap, grec + h[0], bs, &bars[s]);
    // the run streams of the batch are read from L2: start their way there now
    const unsigned runs = 4u * (unsigned)(h[1] - h[15]);
    if (runs) prefetch_l2_bulk(grec + h[0] + h[15], runs);
  };
  if (tid == 0) load_batch(0);

  for (int k = 0; k < d.sb_nbatch; ++k) {
    const int s = k & 1;
    const int32_t* h = hdr + SB_HDR_INTS * k;
    const int32_t* pkg = sec0 + s * seccap;
    if (tid == 0 && k + 1 < d.sb_nbatch) load_batch(k + 1);  // stage (k + 1) & 1 was released by the barrier closing batch k - 1
    mbar_wait(&bars[s], (unsigned)((k >> 1) & 1));
    const int4* gruns = reinterpret_cast<const int4*>(grec + h[0]);
    // ---- A
    {
      const int n_tchunk = h[10];
      const int4* tch = reinterpret_cast<const int4*>(pkg + (h[13] & 0xffff));
      for (int c = tid; c < n_tchunk; c += kThreads) {
        const int4 rec = tch[c];
        const int es = rec.y >> 16;
        if (es == 3) tchunk_factor<3>(rec, pkg, OA, lmd);
        else if (es == 1) tchunk_factor<1>(rec, pkg, OA, lmd);
        else tchunk_factor<2>(rec, pkg, OA, lmd);
      }
      const int p0 = pkg[wid], p1 = pkg[wid + 1];
      run_stream(gruns + (p0 >> 2), (p1 - p0) >> 2, rings + wid, lane, OA, ACC);
    }
    __syncthreads();
    // ---- B
    {
      const int n_mchunk = h[12];
      const int4* mch = reinterpret_cast<const int4*>(pkg + h[14]);
      for (int c = wid; c < n_mchunk; c += SB_WARPS) mchunk_warp(mch[2 * c], mch[2 * c + 1], pkg, OA, lmd);
      const int n_trow = h[11];
      const int4* trw = reinterpret_cast<const int4*>(pkg + ((unsigned)h[13] >> 16));
      // rows from the far end: the warps busy with the chunks above get the fewest
      for (int c = kThreads - 1 - tid; c < n_trow; c += kThreads) {
        const int4 r0 = trw[2 * c], r1 = trw[2 * c + 1];
        const int es = (r0.y >> 8) & 0xff;
        if (es == 3) trow_w<3>(r0, r1, pkg, OA);
        else if (es == 1) trow_w<1>(r0, r1, pkg, OA);
        else trow_w<2>(r0, r1, pkg, OA);
      }
    }
    fence_async_smem();  // the W / factor segments written above are read by the bulk stores below
    __syncthreads();
    // ---- E-buffer and chunk-factor segments to HBM (k_backsub), overlapped with C
    if (tid == 0) {
      double* wb = OA + 2 * scap;
      if (h[7]) bulk_s2g(gEB + h[6], wb, 8u * (unsigned)h[7]);
      if (h[9]) bulk_s2g(gEF + h[8], wb + d.sb_ecap, 8u * (unsigned)h[9]);
      bulk_commit();
    }
    // ---- C
    {
      const int p0 = pkg[SB_WARPS + 1 + wid], p1 = pkg[SB_WARPS + 2 + wid];
      run_stream(gruns + (p0 >> 2), (p1 - p0) >> 2, rings + wid, lane, OA, ACC);
    }
    if (tid == 0) bulk_wait_read0();  // the next batch overwrites the W / factor segments
    __syncthreads();
  }
  SWGN_STAMP(1);
  // ---- write S: rows of the upper triangle + rhs column, zeros where no block cell exists.  Lookup tables in the
  // (now idle) section / operand area: block and offset-in-block of every f tangent index, block sizes, cell map.
  {
    int32_t* tab = sec0;
    const int n_fb = d.n_fb, n_ecols = d.n_ecols, n_e = d.n_e;
    int32_t* t_blk = tab;                       // [nf]  block (0-based among the retained blocks) | offset in block << 16
    int32_t* t_size = t_blk + ((nf + 3) & ~3);  // [n_fb]
    int32_t* t_map = t_size + ((n_fb + 3) & ~3);  // [n_fb * n_fb]
    const int32_t* tcol = v.I(I_TCOL);
    const int32_t* col_pos = v.I(I_COL_POS);
    const int32_t* col_size = v.I(I_COL_SIZE);
    const int32_t* amap = v.I(I_ACC_MAP);
    for (int k = tid; k < nf; k += kThreads) {
      const int c = tcol[n_e + k];
      t_blk[k] = (c - n_ecols) | ((n_e + k - col_pos[c]) << 16);
    }
    for (int k = tid; k < n_fb; k += kThreads) t_size[k] = col_size[n_ecols + k];
    for (int k = tid; k < n_fb * n_fb; k += kThreads) t_map[k] = amap[k];
    __syncthreads();
    double* S = v.W(W_S);
    const double* lmd_f = lmd + n_e;
    for (int i = wid; i < nf; i += SB_WARPS) {
      const int bi = t_blk[i], p = bi & 0xffff, li = bi >> 16;
      const int ps = t_size[p];
      double* row = S + (size_t)i * ld;
      const int dcell = t_map[p * n_fb + p];
      for (int j = i + lane; j <= nf; j += 32) {
        double val = 0.0;
        if (j == nf) {
          val = ACC[dcell + li * (ps + 1) + ps];
        } else {
          const int bj = t_blk[j], q = bj & 0xffff, lj = bj >> 16;
          const int off = t_map[p * n_fb + q];
          if (off >= 0) val = ACC[off + li * (t_size[q] + (p == q ? 1 : 0)) + lj];
          if (j == i) {  // + D^2  (schur_eliminator_impl.h:194-215)
            const double dd = lmd_f[i];
            val += dd * dd;
          }
        }
        row[j] = val;
      }
    }
  }
  SWGN_STAMP(5);
  if (dbg) {
    __syncthreads();
    SWGN_STAMP(6);
    if (tid == 0) {
      unsigned smid;
      asm("mov.u32 %0, %smid;" : "=r"(smid));
      dbg[7] = smid;
    }
  }
  if (b.keep_copy) {
    __syncthreads();
    double* S = v.W(W_S);
    double* SC = v.W(W_SCOPY);
    for (int k = tid; k < nf * ld; k += kThreads) SC[k] = S[k];
  }
  if (tid == 0) {
    st->num_linear_solves += 1;
    st->have_factor = 0;
    st->have_reduced = b.params.export_mode ? 1 : 0;
    st->chol_ok = 0;
  }
#undef SWGN_STAMP
}

void launch_schur_stream(const DeviceBatch& b, int only_window, cudaStream_t s) {
  const int grid = only_window >= 0 ? 1 : b.n_windows;
  k_schur_stream<<<grid, kThreads, b.sb_smem, s>>>(b, only_window);
}

cudaError_t configure_schur_stream(const DeviceBatch& b) {
  if (b.sb_windows <= 0) return cudaSuccess;
  if (b.sb_smem > 227 * 1024) return cudaErrorInvalidValue;
  return cudaFuncSetAttribute(k_schur_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.sb_smem);
}

}  // namespace swgn
