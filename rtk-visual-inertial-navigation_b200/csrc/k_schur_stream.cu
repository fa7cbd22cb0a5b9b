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
#include <mutex>

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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- run streams ---------------------------------------------------------------------------------------------
// A warp's stream is a sequence of 16-byte units in global memory (L2): a run header followed by n/2 units of two
// terms each.  Units arrive in a per-warp ring of 4 chunks x 8 units by cp.async, two chunks ahead of their use.
struct Ring {
  int4 u[32];  // 4 chunks x 8 units
};
static_assert(sizeof(Ring) == SB_RING_BYTES, "ring size is part of the planner's shared-memory budget");

struct Reader {
  const int4* gs;
  uint32_t ring;  // shared-space address
  int n_units, issued, ready, lane;
  __device__ __forceinline__ void issue() {
    const int u = issued * 8 + lane;
    if (lane < 8 && u < n_units) {
      const uint32_t dst = ring + 16u * (unsigned)(u & 31);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gs + u) : "memory");
    }
    cp_async_commit();
    ++issued;
  }
  __device__ __forceinline__ void start(const int4* g, int n, Ring* r, int ln) {
    gs = g;
    n_units = n;
    ring = smem_u32(r);
    lane = ln;
    issued = ready = 0;
    __syncwarp();  // every lane is done with the ring contents of the previous stream
    issue();
    issue();
  }
  // unit u (and everything before it) has landed.  Two chunks are in flight in a ring of four: the chunk issued here
  // (ready + 2) reuses the slot of chunk ready - 2, which every lane has left behind -- a group of two units may still
  // straddle chunks ready - 1 and ready when this is called for its second unit.
  __device__ __forceinline__ void ensure(int u) {
    while (u >= ready * 8) {
      cp_async_wait<1>();
      __syncwarp();  // ... for every lane
      ++ready;
      issue();
    }
  }
  __device__ __forceinline__ int4 unit(int u) const {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ring + 16u * (unsigned)(u & 31)));
    return v;
  }
  __device__ __forceinline__ void finish() { cp_async_wait<0>(); }
};

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// Runs of one phase of one batch.  oa / acc = shared-space byte addresses of the operand area and of the accumulators.
// ecell runs (phase A) store the raw chunk products into the operand area; the others (phase C) accumulate into the
// compact block cells of S.  The inner loop is branch-free: the only predicate is the row mask of the A operand (rows
// beyond the slab contribute exact zeros; everything the B operand or out-of-block lanes read is finite memory of the
// operand area -- zero-filled at kernel entry -- and lands in accumulator elements that are never stored).
__device__ __noinline__ void run_stream(const int4* gs, int n_units, Ring* ring, int lane, uint32_t oa, uint32_t acc) {
  if (n_units <= 0) return;
  const int la = lane & 3, lb = lane >> 2;
  const unsigned lmask = 1u << (16 + la);
  Reader rd;
  rd.start(gs, n_units, ring, lane);
  int u = 0;
  while (u < n_units) {
    rd.ensure(u);
    const int4 h = rd.unit(u);
    ++u;
    const int n_pos = h.z & 0xfff, n_neg = (h.z >> 12) & 0xfff, first = (h.z >> 24) & 1, ecell = (h.z >> 25) & 1;
    const int meta = h.w;
    const int ps = meta & 63, qs = (meta >> 6) & 63, ti = ((meta >> 12) & 7) * 8, tj = ((meta >> 15) & 7) * 8;
    const int diag = (meta >> 18) & 1;
    const bool b_rhs = diag && tj + lb == qs;
    const uint32_t a_lane = oa + 8u * (unsigned)(la * ps + ti + lb);
    const uint32_t b_lane = oa + 8u * (unsigned)(b_rhs ? la : la * qs + tj + lb);
    // my two accumulator elements: (i, j) and (i, j + 1); address 0 = not stored
    const int i = ti + lb, j = tj + 2 * la;
    uint32_t p0 = 0, p1 = 0;
    if (i < ps) {
      if (ecell) {
        if (j < qs) p0 = oa + 8u * (unsigned)(h.x + i * qs + j);
        else if (diag && j == qs) p0 = oa + 8u * (unsigned)(h.y + i);
        if (j + 1 < qs) p1 = oa + 8u * (unsigned)(h.x + i * qs + j + 1);
        else if (diag && j + 1 == qs) p1 = oa + 8u * (unsigned)(h.y + i);
      } else {
        const int stride = qs + diag;
        if (j < stride) p0 = acc + 8u * (unsigned)(h.x + i * stride + j);
        if (j + 1 < stride) p1 = acc + 8u * (unsigned)(h.x + i * stride + j + 1);
      }
    }
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;  // + terms
    double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;  // - terms
    if (!first) {
      if (p0) c0 = lds_f64(p0);
      if (p1) c1 = lds_f64(p1);
    }
#define SWGN_GROUP(X0, X1, Y0, Y1)                                                                         \
  {                                                                                                        \
    rd.ensure(u + 1);                                                                                      \
    const int4 r0 = rd.unit(u), r1 = rd.unit(u + 1);                                                       \
    u += 2;                                                                                                \
    const int w0[4] = {r0.x, r0.z, r1.x, r1.z}, w1[4] = {r0.y, r0.w, r1.y, r1.w};                          \
    double av[4], bv[4];                                                                                   \
    _Pragma("unroll") for (int e = 0; e < 4; ++e) {                                                        \
      const double a = lds_f64(a_lane + 8u * ((unsigned)w0[e] & 0xffffu));                                 \
      av[e] = (w1[e] & lmask) ? a : 0.0;                                                                   \
      bv[e] = lds_f64(b_lane + 8u * (b_rhs ? ((unsigned)w1[e] & 0xffffu) : ((unsigned)w0[e] >> 16)));      \
    }                                                                                                      \
    dmma884(X0, X1, av[0], bv[0]);                                                                         \
    dmma884(Y0, Y1, av[1], bv[1]);                                                                         \
    dmma884(X0, X1, av[2], bv[2]);                                                                         \
    dmma884(Y0, Y1, av[3], bv[3]);                                                                         \
  }
    for (int t = 0; t < n_pos; t += 4) SWGN_GROUP(c0, c1, d0, d1)
    for (int t = 0; t < n_neg; t += 4) SWGN_GROUP(e0, e1, f0, f1)
#undef SWGN_GROUP
    if (p0) sts_f64(p0, (c0 + d0) - (e0 + f0));
    if (p1) sts_f64(p1, (c1 + d1) - (e1 + f1));
  }
  rd.finish();
}

// ---- landmark-like chunks (e-size <= 3, every f-block fed by one row, <= 2 residuals per row): 8 lanes per chunk ----
// Lanes stride over the chunk's rows: partial E'E / E'b in registers, a 3-step shuffle reduction inside the group,
// then every lane holds L L' = D^2 + sum E'E and w_g = L^-1 sum E'b (computed redundantly) and writes
// W_f = (L^-1 E') F_f for its own rows.  A non-positive pivot poisons the chunk with NaN so that the reduced
// factorisation fails and the caller retries with a larger mu (dogleg_strategy.cc:589).
template <int ES>
__device__ __noinline__ void tchunk_group(bool active, const int4 rec, const int32_t* pkg, const int32_t* textra, double* OA,
                                             const double* lmd, int l8) {
  constexpr int NT = ES * (ES + 1) / 2;
  const int n_rows = active ? (rec.y & 0xffff) : 0;
  const int4* rows = reinterpret_cast<const int4*>(pkg + rec.x);
  double ete[NT], g[ES];
#pragma unroll
  for (int k = 0; k < NT; ++k) ete[k] = 0.0;
#pragma unroll
  for (int i = 0; i < ES; ++i) g[i] = 0.0;
  for (int r = l8; r < n_rows; r += 8) {  // ChunkDiagonalBlockAndGradient, schur_eliminator_impl.h:444-507
    const int4 rr = rows[r];
    const double* E = OA + (rr.x & 0xffff);
    const double* bb = OA + rr.y;
    const int nres = rr.x >> 16;
    for (int q = 0; q < nres; ++q) {
      double e[ES];
#pragma unroll
      for (int i = 0; i < ES; ++i) e[i] = E[q * ES + i];
      const double br = bb[q];
      int k = 0;
#pragma unroll
      for (int i = 0; i < ES; ++i) {
        g[i] += e[i] * br;
#pragma unroll
        for (int j = i; j < ES; ++j) ete[k++] += e[i] * e[j];
      }
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
    for (int k = 0; k < NT; ++k) ete[k] += __shfl_xor_sync(0xffffffffu, ete[k], o);
#pragma unroll
    for (int i = 0; i < ES; ++i) g[i] += __shfl_xor_sync(0xffffffffu, g[i], o);
  }
  if (!active) return;
  double A[ES][ES];
  {
    int k = 0;
#pragma unroll
    for (int i = 0; i < ES; ++i) {
#pragma unroll
      for (int j = i; j < ES; ++j) A[i][j] = ete[k++];
      const double dd = lmd[rec.w + i];
      A[i][i] += dd * dd;
    }
  }
  double L[ES][ES], inv[ES];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < ES; ++j) {
    double dj = A[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k];
    if (!(dj > 0.0)) ok = false;
    inv[j] = rsqrt(dj);  // no divide / square root in the dependent chain: L_jj = d * rsqrt(d), the rest multiplies by 1 / L_jj
    L[j][j] = dj * inv[j];
#pragma unroll
    for (int i = j + 1; i < ES; ++i) {
      double t = A[j][i];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
      L[i][j] = t * inv[j];
    }
  }
  if (!ok) {
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      inv[i] = nan("");
#pragma unroll
      for (int j = 0; j <= i; ++j) L[i][j] = nan("");
    }
  }
  if (l8 == 0) {
    double* fac = OA + (rec.z & 0xffff);
#pragma unroll
    for (int i = 0; i < ES; ++i)
#pragma unroll
      for (int j = 0; j < ES; ++j) fac[i * ES + j] = (j <= i) ? L[i][j] : 0.0;
    double* gp = OA + ((unsigned)rec.z >> 16);
    double wg[ES];
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      double t = g[i];
#pragma unroll
      for (int k = 0; k < i; ++k) t -= L[i][k] * wg[k];
      wg[i] = t * inv[i];
      gp[i] = wg[i];
    }
  }
  for (int r = l8; r < n_rows; r += 8) {
    const int4 rr = rows[r];
    const int n_fcells = (rr.w >> 8) & 0xff;
    if (n_fcells == 0) continue;
    const double* E = OA + (rr.x & 0xffff);
    const int nres = rr.x >> 16;
    double v0[ES], v1[ES];  // L^-1 E' for the (<= 2) residual rows
#pragma unroll
    for (int i = 0; i < ES; ++i) {
      double s0 = E[i], s1 = nres > 1 ? E[ES + i] : 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) {
        s0 -= L[i][k] * v0[k];
        s1 -= L[i][k] * v1[k];
      }
      v0[i] = s0 * inv[i];
      v1[i] = s1 * inv[i];
    }
    for (int q = 0; q < n_fcells; ++q) {
      int f_oa, w_oa, fs;
      if (q == 0) {
        f_oa = rr.z & 0xffff;
        w_oa = (unsigned)rr.z >> 16;
        fs = rr.w & 0xff;
      } else {
        const int4 x = reinterpret_cast<const int4*>(textra)[((unsigned)rr.w >> 16) + q - 1];
        f_oa = x.x;
        w_oa = x.y;
        fs = x.z;
      }
      const double* F = OA + f_oa;
      double* Wf = OA + w_oa;
      for (int j = 0; j < fs; ++j) {
        const double f0 = F[j], f1 = nres > 1 ? F[fs + j] : 0.0;
#pragma unroll
        for (int i = 0; i < ES; ++i) Wf[i * fs + j] = v0[i] * f0 + v1[i] * f1;
      }
    }
  }
}

// ---- all other chunks (speed-bias blocks, epoch clocks ...): one warp per chunk, in place in the operand area --------
// in: raw E'E (upper part) at fac, raw E'F_f in the slot blocks, raw E'b in the g slot (phase A, tensor pipe).
// out: lower L at fac (upper part zero), W_f = L^-1 E'F_f, w_g = L^-1 E'b.
// Right-looking Cholesky on the upper triangle (lanes over the trailing elements, rsqrt pivots: no divide or square root
// in the dependent chain), then the lanes own the columns of [E'F ... | E'b] and forward-substitute them.  Compact loops
// on purpose: this runs on one or two warps per batch and must stay resident in the instruction cache.
__device__ __noinline__ void mchunk_warp(const int4 m0, const int4 m1, const int32_t* pkg, double* OA, const double* lmd) {
  const int lane = threadIdx.x & 31;
  const int es = m0.x, epos = m0.y, ns1 = m0.w;
  double* U = OA + m0.z;  // es x es row-major; upper part = E'E
  const int2* slots = reinterpret_cast<const int2*>(pkg + m1.x);
  if (lane < es) {
    const double dd = lmd[epos + lane];
    U[lane * es + lane] += dd * dd;
  }
  __syncwarp();
  bool ok = true;
  for (int j = 0; j < es; ++j) {
    const double dj = U[j * es + j];
    if (!(dj > 0.0)) ok = false;
    const double dinv = rsqrt(dj);
    __syncwarp();
    // row j of U' (= column j of L): scale by 1 / L_jj; the diagonal keeps 1 / L_jj until the end
    for (int i = j + lane; i < es; i += 32) U[j * es + i] = i == j ? dinv : U[j * es + i] * dinv;
    __syncwarp();
    // trailing update of the upper triangle: u_ik -= l_ij l_kj, j < i <= k
    const int m = es - j - 1;
    for (int e = lane; e < m * m; e += 32) {
      const int a = e / m, c = e - a * m;
      if (c < a) continue;
      const int i = j + 1 + a, k = j + 1 + c;
      U[i * es + k] -= U[j * es + i] * U[j * es + k];
    }
    __syncwarp();
  }
  // forward substitution, one column of [E'F ... | E'b] per lane: t_i = (b_i - sum_k<i l_ik t_k) / l_ii, l_ik = U[k][i]
  int total = 0;
  for (int s = 0; s < ns1; ++s) total += slots[s].y;
  for (int c = lane; c < total; c += 32) {
    int s = 0, c0 = 0;
    while (c >= c0 + slots[s].y) {
      c0 += slots[s].y;
      ++s;
    }
    const int2 sl = slots[s];
    const int fs = sl.y;
    double* B = OA + sl.x + (c - c0);
    for (int i = 0; i < es; ++i) {
      double t = B[i * fs];
      for (int k = 0; k < i; ++k) t -= U[k * es + i] * B[k * fs];
      B[i * fs] = ok ? t * U[i * es + i] : nan("");
    }
  }
  __syncwarp();
  // U' -> L in place: transpose the strict upper part into the lower part, diagonal back to L_ii, upper part zero
  for (int e = lane; e < es * es; e += 32) {
    const int i = e / es, k = e - i * es;
    if (k > i) U[k * es + i] = U[i * es + k];
  }
  __syncwarp();
  for (int e = lane; e < es * es; e += 32) {
    const int i = e / es, k = e - i * es;
    if (!ok) U[e] = nan("");
    else if (k > i) U[e] = 0.0;
    else if (k == i) U[e] = 1.0 / U[e];
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
  const uint32_t oa_s = smem_u32(OA), acc_s = smem_u32(ACC);
  // operand loads of the MMA runs are not predicated on block shape: whatever they read beyond a block must be finite
  for (int k = tid; k < (2 * scap + d.sb_ecap + d.sb_fcap + d.sb_acc) / 2; k += kThreads) reinterpret_cast<double2*>(OA)[k] = make_double2(0.0, 0.0);
  const int32_t* ghdr = v.I(I_SB_HDR);
  const int32_t* grec = v.I(I_SB_REC);
  for (int k = tid; k < SB_HDR_INTS * d.sb_nbatch; k += kThreads) hdr[k] = ghdr[k];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();  // the zero fill above is ordered before the bulk copies that land in the same memory
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
  long long t_wait = 0, t_a = 0, t_b = 0, t_c = 0, t_prev = dbg ? clock64() : 0;
#define SWGN_LAP(acc) do { if (dbg && tid == 0) { const long long t_ = clock64(); acc += t_ - t_prev; t_prev = t_; } } while (0)

  for (int k = 0; k < d.sb_nbatch; ++k) {
    const int s = k & 1;
    const int32_t* h = hdr + SB_HDR_INTS * k;
    const int32_t* pkg = sec0 + s * seccap;
    if (tid == 0 && k + 1 < d.sb_nbatch) load_batch(k + 1);  // stage (k + 1) & 1 was released by the barrier closing batch k - 1
    mbar_wait(&bars[s], (unsigned)((k >> 1) & 1));
    SWGN_LAP(t_wait);
    const int4* gruns = reinterpret_cast<const int4*>(grec + h[0]);
    // ---- A: landmark-like chunks start to finish (8 lanes each); raw products of the other chunks on the tensor pipe
    {
      const int n_tchunk = h[10];
      const int4* tch = reinterpret_cast<const int4*>(pkg + h[13]);
      const int32_t* textra = pkg + h[11];
      for (int c0 = 4 * wid; c0 < n_tchunk; c0 += 4 * SB_WARPS) {
        const int c = c0 + (lane >> 3);
        const bool active = c < n_tchunk;
        const int4 rec = active ? tch[c] : make_int4(0, 3 << 16, 0, 0);
        // the shuffles of the group reduction need the whole warp on one path: the e-sizes present in the warp's
        // four chunks take turns
        for (int es = 3; es >= 1; --es) {
          const bool mine = active && (rec.y >> 16) == es;
          if (!__any_sync(0xffffffffu, mine)) continue;
          if (es == 3) tchunk_group<3>(mine, rec, pkg, textra, OA, lmd, lane & 7);
          else if (es == 1) tchunk_group<1>(mine, rec, pkg, textra, OA, lmd, lane & 7);
          else tchunk_group<2>(mine, rec, pkg, textra, OA, lmd, lane & 7);
        }
      }
      const int p0 = pkg[wid], p1 = pkg[wid + 1];
      run_stream(gruns + (p0 >> 2), (p1 - p0) >> 2, rings + wid, lane, oa_s, acc_s);
    }
    const int n_mchunk = h[12];
    if (n_mchunk > 0) {  // (uniform over the CTA)
      __syncthreads();
      if (dbg && tid == 0) {
        const long long t_ = clock64();
        long long* pb = b.debug + 8 * (size_t)(2 * b.n_windows + w);
        if (k == 2) pb[4] = t_ - t_prev;
        if (k == 10) pb[5] = t_ - t_prev;
      }
      SWGN_LAP(t_a);
      // ---- B: factor + forward substitution of the other chunks
      const int4* mch = reinterpret_cast<const int4*>(pkg + h[14]);
      for (int c = wid; c < n_mchunk; c += SB_WARPS) mchunk_warp(mch[2 * c], mch[2 * c + 1], pkg, OA, lmd);
    }
    fence_async_smem();  // the W / factor segments written above are read by the bulk stores below
    __syncthreads();
    if (dbg && tid == 0) {
      const long long t_ = clock64();
      long long* pb = b.debug + 8 * (size_t)(2 * b.n_windows + w);
      if (k == 2) pb[6] = t_ - t_prev;
      if (k == 10) pb[7] = t_ - t_prev;
    }
    SWGN_LAP(t_b);
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
      run_stream(gruns + (p0 >> 2), (p1 - p0) >> 2, rings + wid, lane, oa_s, acc_s);
    }
    if (tid == 0) bulk_wait_read0();  // the next batch overwrites the W / factor segments
    __syncthreads();
    if (dbg && tid == 0) {  // per-batch phase times of four sample batches (development aid)
      const long long t_ = clock64();
      long long* pb = b.debug + 8 * (size_t)(2 * b.n_windows + w);
      if (k == 2) pb[0] = t_ - t_prev;
      if (k == 10) pb[1] = t_ - t_prev;
      if (k == d.sb_nbatch - 3) pb[2] = t_ - t_prev;
      if (k == d.sb_nbatch - 2) pb[3] = t_ - t_prev;
    }
    SWGN_LAP(t_c);
  }
  SWGN_STAMP(1);
  if (dbg && tid == 0) {
    dbg[2] = t_wait;
    dbg[3] = t_a;
    dbg[4] = t_b;
    dbg[5] = t_c;
  }
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
#undef SWGN_LAP
}

void launch_schur_stream(const DeviceBatch& b, int only_window, cudaStream_t s) {
  const int grid = only_window >= 0 ? 1 : b.n_windows;
  k_schur_stream<<<grid, kThreads, b.sb_smem, s>>>(b, only_window);
}

cudaError_t configure_schur_stream(const DeviceBatch& b) {
  if (b.sb_windows <= 0) return cudaSuccess;
  if (b.sb_smem > 227 * 1024) return cudaErrorInvalidValue;
  // the attribute is per function and device: batches of every size share it, so it only ever grows
  static std::mutex mu;
  static unsigned granted[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 0 && dev < 64 && b.sb_smem <= granted[dev]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(k_schur_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.sb_smem);
  if (e == cudaSuccess && dev >= 0 && dev < 64) granted[dev] = b.sb_smem;
  return e;
}

}  // namespace swgn
