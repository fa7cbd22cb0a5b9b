// C ABI of the solver (include/swgn.h): window planning on the host, pools in HBM, and the tick
// loop that drives the per-window trust-region state machines (k_tr.cu).  There is no CPU compute
// path in here: without a CUDA device swgn_batch_create fails with SWGN_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/swgn.h"
#include "kernels.cuh"
#include "plan.h"

namespace swgn {
cudaError_t configure_schur(const DeviceBatch& b);
cudaError_t configure_chol(const DeviceBatch& b);
void launch_gather_states(const DeviceBatch& b, double* dst, const int64_t* offs, int to_device, cudaStream_t s);
}  // namespace swgn

using namespace swgn;

namespace {
thread_local std::string g_err;
swgn_status fail(swgn_status st, const std::string& m) {
  g_err = m;
  return st;
}
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(SWGN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)
}  // namespace
namespace swgn {
// other translation units of the library (gnss_epoch.cpp) report through the same thread-local message
swgn_status set_error(swgn_status st, const std::string& m) { return fail(st, m); }
// The device's default memory pool releases what is unused at every synchronisation (release threshold 0), so a call that
// takes its scratch with cudaMallocAsync re-grows the pool each time: 50-70 ms per 80 MB, and now and then a stall of
// hundreds of ms when the release coincides with the next allocation.  Keep the pool's memory once it has grown.
void keep_pool_memory(int device) {
  static std::mutex mu;
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[device] = true;
}
}  // namespace swgn
namespace {

// ---- device / pinned slabs.  A batch takes ONE device allocation and ONE pinned allocation and carves its
// arrays out of them; destroyed batches park their slabs in a small process-wide cache so that the
// create-solve-destroy cycle of a per-frame host (one ceres::Solve per window, as the reference runs it) costs
// no cudaMalloc / cudaFree / cudaMallocHost after the first frame.  Big slabs are returned to the driver.
struct SlabCache {
  struct Slab {
    void* p;
    size_t cap;
    int device;
  };
  std::mutex mu;
  std::vector<Slab> slabs[2];  // [0] device, [1] pinned host
  size_t held[2] = {0, 0};
  // per kind: largest slab worth keeping, most bytes held (device memory is plentiful, pinned host memory is not)
  static constexpr size_t kMaxSlab[2] = {(size_t)2 << 30, (size_t)256 << 20}, kMaxHeld[2] = {(size_t)4 << 30, (size_t)1 << 30};
  static constexpr size_t kMaxCount = 16, kHeadRoomBelow = (size_t)128 << 20;
};
SlabCache g_slabs;
}  // namespace
extern "C" int64_t swgn_release_cached_memory(void);
namespace {

cudaError_t slab_alloc(int kind, int device, size_t bytes, void** out, size_t* cap) {
  bytes = std::max<size_t>(bytes, 256);
  {
    std::lock_guard<std::mutex> lk(g_slabs.mu);
    auto& v = g_slabs.slabs[kind];
    int best = -1;
    for (int i = 0; i < (int)v.size(); ++i)
      if (v[i].device == device && v[i].cap >= bytes && v[i].cap <= 2 * bytes + ((size_t)1 << 20) && (best < 0 || v[i].cap < v[best].cap))
        best = i;
    if (best >= 0) {
      *out = v[best].p;
      *cap = v[best].cap;
      g_slabs.held[kind] -= v[best].cap;
      v.erase(v.begin() + best);
      return cudaSuccess;
    }
  }
  // cacheable sizes get 25 % head room (64 KB granules): the next frame's window is rarely the same size, and
  // pinning a fresh buffer costs ~0.1 s
  if (bytes <= SlabCache::kHeadRoomBelow) bytes = (bytes + bytes / 4 + 65535) & ~(size_t)65535;
  *cap = bytes;
  cudaError_t e = kind == 0 ? cudaMalloc(out, bytes) : cudaMallocHost(out, bytes);
  if (e == cudaErrorMemoryAllocation) {
    // the parked slabs of destroyed batches (other size classes, other devices) may be what is in the way: hand them
    // back to the driver and try once more
    cudaGetLastError();
    swgn_release_cached_memory();
    cudaSetDevice(device);
    e = kind == 0 ? cudaMalloc(out, bytes) : cudaMallocHost(out, bytes);
  }
  return e;
}
void slab_free(int kind, int device, void* p, size_t cap) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_slabs.mu);
    auto& v = g_slabs.slabs[kind];
    if (cap <= SlabCache::kMaxSlab[kind] && g_slabs.held[kind] + cap <= SlabCache::kMaxHeld[kind] && v.size() < SlabCache::kMaxCount) {
      v.push_back({p, cap, device});
      g_slabs.held[kind] += cap;
      return;
    }
  }
  if (kind == 0) cudaFree(p);
  else cudaFreeHost(p);
}
}  // namespace

// Everything of a graph that the planner turned into index tables: block table, ordering, constness, every factor's
// block list, prior shapes, chain shapes, program order and is_use masks.  swgn_batch_update_inputs re-packs constants
// against the tables of the batch, so a graph is only accepted when this fingerprint is the one it was planned with.
static uint64_t structure_fingerprint(const swgn_graph* g) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t bytes) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < bytes; ++i) h = (h ^ c[i]) * 1099511628211ull;
  };
  auto arr = [&](const int32_t* p, size_t n) {
    const uint64_t tag = p ? n : ~0ull;
    mix(&tag, sizeof(tag));
    if (p && n) mix(p, sizeof(int32_t) * n);
  };
  const int32_t counts[10] = {g->n_blocks, g->n_state, g->n_proj, g->n_imu, g->n_gnss, g->n_prior, g->n_unit, g->n_order, g->n_chain,
                              g->proj_cauchy_a > 0 ? 1 : 0};
  mix(counts, sizeof(counts));
  const size_t nb = (size_t)std::max(0, g->n_blocks);
  arr(g->block_size, nb);
  arr(g->block_manifold, nb);
  arr(g->block_const, nb);
  arr(g->block_group, nb);
  arr(g->block_offset, nb);
  arr(g->proj_blocks, 3 * (size_t)std::max(0, g->n_proj));
  arr(g->imu_blocks, 4 * (size_t)std::max(0, g->n_imu));
  arr(g->gnss_kind, (size_t)std::max(0, g->n_gnss));
  arr(g->gnss_blocks, 3 * (size_t)std::max(0, g->n_gnss));
  const size_t np = (size_t)std::max(0, g->n_prior);
  arr(g->prior_n, np);
  arr(g->prior_blk_begin, np ? np + 1 : 0);
  const size_t npb = np && g->prior_blk_begin ? (size_t)g->prior_blk_begin[np] : 0;
  arr(g->prior_blocks, npb);
  arr(g->prior_blk_idx, npb);
  arr(g->unit_block, (size_t)std::max(0, g->n_unit));
  arr(reinterpret_cast<const int32_t*>(g->order), g->order ? (size_t)std::max(0, g->n_order) : 0);
  const size_t nc = (size_t)std::max(0, g->n_chain);
  arr(g->chain_blk_begin, nc ? nc + 1 : 0);
  arr(g->chain_blocks, nc && g->chain_blk_begin ? (size_t)g->chain_blk_begin[nc] : 0);
  arr(g->chain_frame_begin, nc ? nc + 1 : 0);
  const size_t nh = (size_t)std::max(0, g->n_host);
  arr(g->host_nres, nh);
  arr(g->host_blk_begin, nh ? nh + 1 : 0);
  arr(g->host_blocks, nh && g->host_blk_begin ? (size_t)g->host_blk_begin[nh] : 0);
  const size_t nf = (size_t)std::max(0, g->n_proj) + std::max(0, g->n_imu) + std::max(0, g->n_gnss) + np + std::max(0, g->n_unit) + nc + nh;
  const uint64_t use_tag = g->is_use ? nf : ~0ull;
  mix(&use_tag, sizeof(use_tag));
  if (g->is_use) mix(g->is_use, nf);
  return h;
}

// host-evaluated residual blocks of one window (swgn_graph.host_*): what swgn_batch_solve needs to call back
struct HostFactors {
  swgn_host_eval_fn fn = nullptr;
  void* user = nullptr;
  std::vector<int32_t> nres, blk_begin, blk_state, blk_size;  // per factor / per block: state offset and global size
  int64_t buf_doubles = 0;
};

struct swgn_batch {
  std::vector<HostFactors> host;       // per window (empty vectors when the window has none)
  bool any_host = false;
  std::vector<double> h_hoststate, h_hostbuf;
  std::vector<uint64_t> fingerprint;  // per window: structure_fingerprint of the graph it was planned from
  int n = 0, device = 0;
  swgn_options opt;
  cudaStream_t stream = nullptr;
  std::vector<WinDesc> desc;
  std::vector<int64_t> schur_doubles;
  std::vector<int64_t> state_off;  // prefix offsets of the packed state staging buffer
  WinDesc* d_desc = nullptr;
  int32_t* d_ipool = nullptr;
  double* d_cpool = nullptr;
  double* d_wpool = nullptr;
  TRState* d_state = nullptr;
  int32_t* d_counters = nullptr;
  int32_t* h_counters = nullptr;  // pinned
  double* d_stage = nullptr;      // packed states of all windows
  int64_t* d_state_off = nullptr;
  double* h_stage = nullptr;      // pinned
  double* h_cpool = nullptr;      // pinned staging of the factor constants (update_inputs)
  // double-buffered inputs (swgn_batch_prefetch_inputs / _commit_inputs)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_prefetch = nullptr;
  double* blocks[2] = {nullptr, nullptr};  // [constants | packed states]; they alternate as live / shadow
  int live_block = -1;            // -1: the constants inside the slab are live (until the first commit)
  int prefetched = -1;            // block holding prefetched, not yet committed inputs
  double* h_stage2 = nullptr;     // pinned staging of the prefetched states
  void* d_slab = nullptr;         // every d_* array above is carved out of this one allocation,
  void* h_slab = nullptr;         // h_counters and h_stage out of this one (slab_alloc / slab_free)
  size_t d_slab_cap = 0, h_slab_cap = 0;
  long long* d_debug = nullptr;   // SWGN_DEBUG_TIMELINE=1: per-window phase timestamps of k_schur
  size_t ipool_n = 0, cpool_n = 0, wpool_n = 0;
  DeviceBatch db;
  std::vector<TRState> h_state;
  bool has_scopy = false;
  // timing of the last solve
  std::vector<cudaEvent_t> ev;  // [0] start, [1] stop, then pairs around every Schur launch
  cudaEvent_t ev_count[2] = {nullptr, nullptr};  // the active-window counter of a tick has reached the host
  double total_ms = 0, schur_ms = 0;
  int schur_launches = 0, kernel_launches = 0;
  bool solved = false;
};

extern "C" {

void swgn_default_options(swgn_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->max_num_iterations = 8;
  o->max_num_consecutive_invalid_steps = 5;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->dogleg_min_mu = 1e-12;
  o->is_optimize = 1;
  o->n_parameter_head = 0;
  o->device = 0;
}

const char* swgn_last_error(void) { return g_err.c_str(); }
const char* swgn_version(void) { return "swgn 0.1 (sm_100a)"; }
int32_t swgn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int64_t swgn_release_cached_memory(void) {
  std::vector<SlabCache::Slab> out[2];
  {
    std::lock_guard<std::mutex> lk(g_slabs.mu);
    for (int kind = 0; kind < 2; ++kind) {
      out[kind].swap(g_slabs.slabs[kind]);
      g_slabs.held[kind] = 0;
    }
  }
  int64_t bytes = 0;
  int prev = -1;
  if (!out[0].empty() || !out[1].empty()) cudaGetDevice(&prev);
  for (int kind = 0; kind < 2; ++kind)
    for (const SlabCache::Slab& s : out[kind]) {
      cudaSetDevice(s.device);
      if (kind == 0) cudaFree(s.p);
      else cudaFreeHost(s.p);
      bytes += (int64_t)s.cap;
    }
  if (prev >= 0) cudaSetDevice(prev);
  return bytes;
}

void swgn_batch_destroy(swgn_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (cudaEvent_t e : b->ev)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : b->ev_count)
    if (e) cudaEventDestroy(e);
  slab_free(0, b->device, b->d_slab, b->d_slab_cap);
  slab_free(1, b->device, b->h_slab, b->h_slab_cap);
  cudaFree(b->d_debug);
  if (b->h_cpool) cudaFreeHost(b->h_cpool);
  if (b->copy_stream) {
    cudaStreamSynchronize(b->copy_stream);
    cudaStreamDestroy(b->copy_stream);
  }
  if (b->ev_prefetch) cudaEventDestroy(b->ev_prefetch);
  if (b->h_stage2) cudaFreeHost(b->h_stage2);
  for (double* blk : b->blocks)
    if (blk) cudaFree(blk);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

swgn_status swgn_batch_create(const swgn_options* options, int32_t n_windows, const swgn_graph* const* graphs,
                              swgn_batch** out) {
  if (!options || !graphs || !out || n_windows <= 0) return fail(SWGN_ERR_INVALID, "bad arguments");
  *out = nullptr;
  if (options->trust_region_strategy != SWGN_DOGLEG && options->trust_region_strategy != SWGN_LEVENBERG_MARQUARDT)
    return fail(SWGN_ERR_UNSUPPORTED, "unknown trust-region strategy");
  if (options->jacobi_scaling && options->trust_region_strategy != SWGN_LEVENBERG_MARQUARDT)
    return fail(SWGN_ERR_UNSUPPORTED, "jacobi_scaling is implemented for the LEVENBERG_MARQUARDT strategy only (the reference's DOGLEG solve sets it to false)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || options->device < 0 || options->device >= ndev)
    return fail(SWGN_ERR_NO_DEVICE, "no usable CUDA device (the solver has no CPU fallback)");
  CU(cudaSetDevice(options->device));

  const bool dbg_t = std::getenv("SWGN_DEBUG_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_start = now();
  double t_mark = t_start;
  auto mark = [&](const char* what) {
    if (!dbg_t) return;
    const double t = now();
    std::fprintf(stderr, "[swgn_batch_create] %-18s %8.1f ms\n", what, 1e3 * (t - t_mark));
    t_mark = t;
  };
  // ---- plan every window (host, structural work only), in parallel
  std::vector<WindowPlan> plans(n_windows);
  std::vector<swgn_status> sts(n_windows, SWGN_OK);
  std::vector<std::string> errs(n_windows);
  std::vector<uint64_t> fingerprints(n_windows);
  // per-window host work on all host threads (batches of thousands of tiny graphs -- the per-epoch GNSS problems -- spend
  // as long in the constructors, destructors and fingerprints of their plans as in the planning itself)
  auto on_all_threads = [&](auto&& per_window) {
    const int nt = std::max(1, std::min<int>(n_windows / 4, (int)std::thread::hardware_concurrency()));
    std::atomic<int> next(0);
    auto work = [&]() {
      for (;;) {
        const int w0 = next.fetch_add(8);
        if (w0 >= n_windows) break;
        for (int w = w0; w < std::min(n_windows, w0 + 8); ++w) per_window(w);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  };
  on_all_threads([&](int w) {
    sts[w] = build_plan(graphs[w], options->n_parameter_head, &plans[w], &errs[w]);
    if (sts[w] == SWGN_OK) fingerprints[w] = structure_fingerprint(graphs[w]);
  });
  for (int w = 0; w < n_windows; ++w)
    if (sts[w] != SWGN_OK) return fail(sts[w], "window " + std::to_string(w) + ": " + errs[w]);

  mark("plan");
  swgn_batch* b = new swgn_batch();
  b->n = n_windows;
  b->device = options->device;
  b->opt = *options;
  b->desc.resize(n_windows);
  b->schur_doubles.resize(n_windows);
  b->state_off.resize(n_windows + 1);
  b->fingerprint = std::move(fingerprints);
  b->host.resize(n_windows);
  for (int w = 0; w < n_windows; ++w) {
    const swgn_graph* g = graphs[w];
    if (g->n_host <= 0) continue;
    HostFactors& h = b->host[w];
    b->any_host = true;
    h.fn = g->host_eval;
    h.user = g->host_user;
    h.nres.assign(g->host_nres, g->host_nres + g->n_host);
    h.blk_begin.assign(g->host_blk_begin, g->host_blk_begin + g->n_host + 1);
    for (int k = 0; k < g->host_blk_begin[g->n_host]; ++k) {
      h.blk_state.push_back(g->block_offset[g->host_blocks[k]]);
      h.blk_size.push_back(g->block_size[g->host_blocks[k]]);
    }
    h.buf_doubles = plans[w].d.n_hostbuf;
  }
  b->has_scopy = n_windows <= 256;
  size_t io = 0, co = 0, wo = 0;
  int max_wbuf = 0, max_nf = 0, max_prior_n = 0, max_chain = 0, max_chain_k = 0, sb_windows = 0;
  size_t sb_smem = 0;
  auto al = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
  int64_t so = 0;
  for (int w = 0; w < n_windows; ++w) {
    WindowPlan& p = plans[w];
    WinDesc& d = p.d;
    for (int a = 0; a < NUM_IARR; ++a) {
      d.ioff[a] = (int64_t)io;
      io = al(io + p.iarr[a].size(), 4);
    }
    for (int a = 0; a < NUM_CARR; ++a) {
      d.coff[a] = (int64_t)co;
      co = al(co + p.carr[a].size(), 2);
    }
    for (int a = 0; a < NUM_WARR; ++a) {
      d.woff[a] = (int64_t)wo;
      if (a == W_SCOPY && !b->has_scopy) continue;
      wo += (size_t)p.wsize[a];  // sizes are even: 16-byte alignment is preserved, JW stays contiguous
    }
    b->desc[w] = d;
    b->schur_doubles[w] = p.schur_doubles;
    b->state_off[w] = so;
    so += d.n_state;
    max_wbuf = std::max(max_wbuf, d.max_wbuf);
    max_nf = std::max(max_nf, d.n_f);
    max_prior_n = std::max(max_prior_n, d.max_prior_n);
    max_chain = std::max(max_chain, d.n_chain);
    max_chain_k = std::max(max_chain_k, d.max_chain_k);
    if (d.sb_ok) {
      ++sb_windows;
      sb_smem = std::max(sb_smem, stream_smem_bytes(d.sb_nbatch, d.sb_acc, d.sb_jcap, d.sb_rcap, d.sb_ecap, d.sb_fcap, d.sb_reccap));
    }
    if (options->n_parameter_head > 0 && d.n_f >= 1024) {
      swgn_batch_destroy(b);
      return fail(SWGN_ERR_TOO_LARGE, "reduced system has >= 1024 rows with exports requested");
    }
  }
  b->state_off[n_windows] = so;
  b->ipool_n = io;
  b->cpool_n = co;
  b->wpool_n = wo;

  auto bail = [&](cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    swgn_batch_destroy(b);
    return fail(SWGN_ERR_CUDA, m);
  };
#define CB(call)                                  \
  do {                                            \
    cudaError_t e_ = (call);                      \
    if (e_ != cudaSuccess) return bail(e_, #call); \
  } while (0)
  CB(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  {
    // one device slab: [desc | ipool | cpool | wpool | state | counters | stage | state_off], 256-byte aligned parts
    const size_t sz[8] = {sizeof(WinDesc) * (size_t)n_windows,      sizeof(int32_t) * std::max<size_t>(io, 4),
                          sizeof(double) * std::max<size_t>(co, 2), sizeof(double) * std::max<size_t>(wo, 2),
                          sizeof(TRState) * (size_t)n_windows,      sizeof(int32_t) * 4,
                          sizeof(double) * (size_t)std::max<int64_t>(so, 2), sizeof(int64_t) * ((size_t)n_windows + 1)};
    size_t off[9] = {0};
    for (int k = 0; k < 8; ++k) off[k + 1] = al(off[k] + sz[k], 256);
    CB(slab_alloc(0, b->device, off[8], &b->d_slab, &b->d_slab_cap));
    char* base = static_cast<char*>(b->d_slab);
    b->d_desc = reinterpret_cast<WinDesc*>(base + off[0]);
    b->d_ipool = reinterpret_cast<int32_t*>(base + off[1]);
    b->d_cpool = reinterpret_cast<double*>(base + off[2]);
    b->d_wpool = reinterpret_cast<double*>(base + off[3]);
    b->d_state = reinterpret_cast<TRState*>(base + off[4]);
    b->d_counters = reinterpret_cast<int32_t*>(base + off[5]);
    b->d_stage = reinterpret_cast<double*>(base + off[6]);
    b->d_state_off = reinterpret_cast<int64_t*>(base + off[7]);
    const size_t hs = sizeof(double) * (size_t)std::max<int64_t>(so, 2);
    CB(slab_alloc(1, b->device, 256 + hs, &b->h_slab, &b->h_slab_cap));
    b->h_counters = static_cast<int32_t*>(b->h_slab);
    b->h_stage = reinterpret_cast<double*>(static_cast<char*>(b->h_slab) + 256);
  }
  mark("alloc");
  CB(cudaMemsetAsync(b->d_wpool, 0, sizeof(double) * std::max<size_t>(wo, 2), b->stream));
  CB(cudaMemsetAsync(b->d_state, 0, sizeof(TRState) * n_windows, b->stream));
  {
    // stream the pools through two pinned staging buffers: worker threads pack one chunk of consecutive
    // windows (the pools are window-major) while the previous chunk is in flight to the device
    auto i_begin = [&](int w) { return w < n_windows ? (size_t)b->desc[w].ioff[0] : io; };
    auto c_begin = [&](int w) { return w < n_windows ? (size_t)b->desc[w].coff[0] : co; };
    auto chunk_bytes = [&](int w0, int w1) {
      return sizeof(int32_t) * (i_begin(w1) - i_begin(w0)) + 16 + sizeof(double) * (c_begin(w1) - c_begin(w0));
    };
    size_t stage_bytes = std::min<size_t>((size_t)128 << 20, chunk_bytes(0, n_windows));
    for (int w = 0; w < n_windows; ++w) stage_bytes = std::max(stage_bytes, chunk_bytes(w, w + 1));
    // Pinning memory costs ~0.1 s per allocation, so the two pinned staging buffers are process-wide, created
    // by the first large batch and reused by every later create (a real replay creates a batch per frame).
    // Small batches, or a create that finds the pool busy in another thread, stage through one pageable
    // buffer (the copy is then synchronous with respect to the host).
    static std::mutex pool_mutex;
    static unsigned char* pool[2] = {nullptr, nullptr};
    static size_t pool_bytes = 0;
    std::unique_lock<std::mutex> pool_lock(pool_mutex, std::defer_lock);
    bool pinned = chunk_bytes(0, n_windows) >= ((size_t)64 << 20) && pool_lock.try_lock();
    std::vector<unsigned char> pageable;
    unsigned char* stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaError_t ce = cudaSuccess;
    if (pinned && pool_bytes < stage_bytes) {
      for (int q = 0; q < 2; ++q) {
        if (pool[q]) cudaFreeHost(pool[q]);
        pool[q] = nullptr;
      }
      pool_bytes = 0;
      if (cudaMallocHost((void**)&pool[0], stage_bytes) == cudaSuccess && cudaMallocHost((void**)&pool[1], stage_bytes) == cudaSuccess) {
        pool_bytes = stage_bytes;
      } else {
        for (int q = 0; q < 2; ++q) {
          if (pool[q]) cudaFreeHost(pool[q]);
          pool[q] = nullptr;
        }
        cudaGetLastError();
        pinned = false;
        pool_lock.unlock();
      }
    }
    if (pinned) {
      stage[0] = pool[0];
      stage[1] = pool[1];
    } else {
      pageable.resize(stage_bytes);
      stage[0] = stage[1] = pageable.data();
    }
    for (int q = 0; q < 2 && ce == cudaSuccess; ++q) ce = cudaEventCreateWithFlags(&ev[q], cudaEventDisableTiming);
    const int hw = std::max(1, (int)std::thread::hardware_concurrency());
    int w0 = 0;
    for (int c = 0; w0 < n_windows && ce == cudaSuccess; ++c) {
      int w1 = w0 + 1;
      while (w1 < n_windows && chunk_bytes(w0, w1 + 1) <= stage_bytes) ++w1;
      unsigned char* buf = stage[c & 1];
      if (c >= (pinned ? 2 : 1)) ce = cudaEventSynchronize(ev[(c + (pinned ? 0 : 1)) & 1]);  // the buffer is free again
      if (ce != cudaSuccess) break;
      const size_t ib = i_begin(w0), cb = c_begin(w0);
      const size_t ibytes = (sizeof(int32_t) * (i_begin(w1) - ib) + 15) & ~(size_t)15;
      int32_t* hi = reinterpret_cast<int32_t*>(buf);
      double* hc = reinterpret_cast<double*>(buf + ibytes);
      std::atomic<int> next(w0);
      auto work = [&]() {
        for (;;) {
          const int w = next.fetch_add(1);
          if (w >= w1) break;
          const WindowPlan& p = plans[w];
          std::memset(hi + (i_begin(w) - ib), 0, sizeof(int32_t) * (i_begin(w + 1) - i_begin(w)));
          std::memset(hc + (c_begin(w) - cb), 0, sizeof(double) * (c_begin(w + 1) - c_begin(w)));
          for (int a = 0; a < NUM_IARR; ++a)
            if (!p.iarr[a].empty()) std::memcpy(hi + (b->desc[w].ioff[a] - ib), p.iarr[a].data(), sizeof(int32_t) * p.iarr[a].size());
          for (int a = 0; a < NUM_CARR; ++a)
            if (!p.carr[a].empty()) std::memcpy(hc + (b->desc[w].coff[a] - cb), p.carr[a].data(), sizeof(double) * p.carr[a].size());
          std::memcpy(b->h_stage + b->state_off[w], p.state.data(), sizeof(double) * p.state.size());
        }
      };
      {
        const int nt = std::max(1, std::min(hw, (w1 - w0) / 4));
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
      }
      if (i_begin(w1) > ib)
        ce = cudaMemcpyAsync(b->d_ipool + ib, hi, sizeof(int32_t) * (i_begin(w1) - ib), cudaMemcpyHostToDevice, b->stream);
      if (ce == cudaSuccess && c_begin(w1) > cb)
        ce = cudaMemcpyAsync(b->d_cpool + cb, hc, sizeof(double) * (c_begin(w1) - cb), cudaMemcpyHostToDevice, b->stream);
      if (ce == cudaSuccess) ce = cudaEventRecord(ev[c & 1], b->stream);
      w0 = w1;
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(b->d_desc, b->desc.data(), sizeof(WinDesc) * n_windows, cudaMemcpyHostToDevice, b->stream);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(b->d_state_off, b->state_off.data(), sizeof(int64_t) * (n_windows + 1), cudaMemcpyHostToDevice, b->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(b->d_stage, b->h_stage, sizeof(double) * so, cudaMemcpyHostToDevice, b->stream);
    cudaError_t cs = cudaStreamSynchronize(b->stream);  // the staging buffers go away
    if (ce == cudaSuccess) ce = cs;
    for (int q = 0; q < 2; ++q)
      if (ev[q]) cudaEventDestroy(ev[q]);
    if (ce != cudaSuccess) return bail(ce, "pool upload");
  }
  on_all_threads([&](int w) { plans[w] = WindowPlan(); });  // the tables are on the device: release them in parallel
  mark("pack + upload");
  DeviceBatch& db = b->db;
  std::memset(&db, 0, sizeof(db));
  db.n_windows = n_windows;
  db.desc = b->d_desc;
  db.ipool = b->d_ipool;
  db.cpool = b->d_cpool;
  db.wpool = b->d_wpool;
  db.state = b->d_state;
  db.counters = b->d_counters;
  db.max_wbuf = max_wbuf;
  db.max_nf = max_nf;
  db.max_prior_n = max_prior_n;
  db.max_chain = max_chain;
  db.max_chain_k = max_chain_k;
  db.chain_epoch = 1;
  db.keep_copy = 0;
  db.sb_windows = sb_windows;
  db.gather_windows = n_windows - sb_windows;
  db.sb_smem = (unsigned)sb_smem;
  if (std::getenv("SWGN_DEBUG_TIMELINE")) {
    CB(cudaMalloc(&b->d_debug, sizeof(long long) * 24 * n_windows));
    CB(cudaMemset(b->d_debug, 0, sizeof(long long) * 24 * n_windows));
  }
  db.debug = b->d_debug;
  SolverParams& P = db.params;
  P.max_num_iterations = options->max_num_iterations;
  P.max_num_consecutive_invalid_steps = options->max_num_consecutive_invalid_steps;
  P.initial_radius = options->initial_trust_region_radius;
  P.max_radius = options->max_trust_region_radius;
  P.min_radius = options->min_trust_region_radius;
  P.min_relative_decrease = options->min_relative_decrease;
  P.min_lm_diagonal = options->min_lm_diagonal;
  P.max_lm_diagonal = options->max_lm_diagonal;
  P.function_tolerance = options->function_tolerance;
  P.gradient_tolerance = options->gradient_tolerance;
  P.parameter_tolerance = options->parameter_tolerance;
  P.min_mu = options->dogleg_min_mu;
  P.is_optimize = options->is_optimize;
  P.n_parameter_head = options->n_parameter_head;
  P.export_mode = (options->n_parameter_head > 0 && !options->is_optimize) ? 1 : 0;
  P.strategy = options->trust_region_strategy;
  P.jacobi_scaling = options->jacobi_scaling ? 1 : 0;
  P.max_radius_lm = options->max_trust_region_radius;
  {
    // a window whose per-CTA working set exceeds the 227 KB of shared memory (reduced system too wide for the Cholesky
    // panel, chunk scratch too large) is a size limit of this implementation, not a CUDA failure
    auto too_large = [&](cudaError_t e, const char* what) {
      if (e != cudaErrorInvalidValue) return bail(e, what);
      const std::string m = std::string(what) + ": a window needs more shared memory per CTA than the device offers (largest reduced system " +
                            std::to_string(db.max_nf) + " rows)";
      swgn_batch_destroy(b);
      return fail(SWGN_ERR_TOO_LARGE, m);
    };
    cudaError_t e = configure_schur(db);
    if (e != cudaSuccess) return too_large(e, "configure_schur");
    e = configure_schur_stream(db);
    if (e != cudaSuccess) return too_large(e, "configure_schur_stream");
    e = configure_chol(db);
    if (e != cudaSuccess) return too_large(e, "configure_chol");
    e = configure_chain(db);
    if (e != cudaSuccess) return too_large(e, "configure_chain");
  }
  launch_gather_states(db, b->d_stage, b->d_state_off, 1, b->stream);
  CB(cudaGetLastError());
  CB(cudaStreamSynchronize(b->stream));
  b->h_state.resize(n_windows);
  *out = b;
  mark("configure");
  return SWGN_OK;
#undef CB
}

// Host-only planning probe: runs the preprocessing of one window (no device needed) and reports
// info[0..11] = n_cols, n_ecols, n_e, n_f, n_t, n_res, n_rows, n_chunks, n_jac, n_scells, n_sterms,
// n_srows; info[12..13] = algorithmic Schur bytes (low, high 32 bits); info[14] = MMAs per gather pass.
swgn_status swgn_plan_probe(const swgn_graph* g, int32_t n_parameter_head, int32_t* info) {
  if (!g || !info) return fail(SWGN_ERR_INVALID, "bad arguments");
  WindowPlan p;
  std::string err;
  swgn_status st = build_plan(g, n_parameter_head, &p, &err);
  if (st != SWGN_OK) return fail(st, err);
  const WinDesc& d = p.d;
  const int32_t v[12] = {d.n_cols, d.n_ecols, d.n_e, d.n_f, d.n_t, d.n_res, d.n_rows, d.n_chunks, d.n_jac, d.n_scells, d.n_sterms, d.n_srows};
  std::memcpy(info, v, sizeof(v));
  const int64_t bytes = 8 * p.schur_doubles;
  info[12] = (int32_t)(bytes & 0xffffffff);
  info[13] = (int32_t)(bytes >> 32);
  info[14] = (int32_t)p.n_mma;
  info[15] = 0;
  return SWGN_OK;
}

// Host-only: the symbolic fill-in masks k_chol uses (one 64-bit mask of live 16-column groups per 32-row
// panel of the reduced system); n_panels may be queried with masks NULL.
swgn_status swgn_plan_chol_masks(const swgn_graph* g, int32_t n_parameter_head, int32_t* n_panels, uint64_t* masks) {
  if (!g || !n_panels) return fail(SWGN_ERR_INVALID, "bad arguments");
  WindowPlan p;
  std::string err;
  swgn_status st = build_plan(g, n_parameter_head, &p, &err);
  if (st != SWGN_OK) return fail(st, err);
  const std::vector<int32_t>& m = p.iarr[I_CHOL_MASK];
  *n_panels = (int32_t)(m.size() / 2);
  if (masks)
    for (size_t k = 0; k < m.size() / 2; ++k) masks[k] = (uint64_t)(uint32_t)m[2 * k] | ((uint64_t)(uint32_t)m[2 * k + 1] << 32);
  return SWGN_OK;
}

// Host-only: decode the per-warp gather streams of one window and check their invariants (every operand
// offset inside the window's J | EBUF | residual range, every store inside S / W_EFAC / W_EBUF, one `end` per
// tile).  out[0..7] = stages, live terms, padding terms, tiles (end flags) of the reduced-system streams, min and
// max stages per warp, stages and live terms of the e-cell streams.
swgn_status swgn_plan_order(const swgn_graph* g, int32_t n_parameter_head, int32_t* n_cols, int32_t* col_block, int32_t* n_rows,
                            int32_t* row_factor) {
  if (!g) return fail(SWGN_ERR_INVALID, "bad arguments");
  WindowPlan p;
  std::string err;
  swgn_status st = build_plan(g, n_parameter_head, &p, &err);
  if (st != SWGN_OK) return fail(st, err);
  if (n_cols) *n_cols = p.d.n_cols;
  if (n_rows) *n_rows = p.d.n_rows;
  if (col_block) std::copy(p.iarr[I_COL_BLOCK].begin(), p.iarr[I_COL_BLOCK].begin() + p.d.n_cols, col_block);
  if (row_factor) std::copy(p.iarr[I_ROW_FACTOR].begin(), p.iarr[I_ROW_FACTOR].begin() + p.d.n_rows, row_factor);
  return SWGN_OK;
}

swgn_status swgn_plan_stream_check(const swgn_graph* g, int32_t n_parameter_head, int64_t* out) {
  if (!g || !out) return fail(SWGN_ERR_INVALID, "bad arguments");
  WindowPlan p;
  std::string err;
  swgn_status st = build_plan(g, n_parameter_head, &p, &err);
  if (st != SWGN_OK) return fail(st, err);
  const WinDesc& d = p.d;
  const int64_t jw = p.wsize[W_JAC] + p.wsize[W_EBUF] + p.wsize[W_RES];
  std::memset(out, 0, sizeof(int64_t) * 8);
  out[4] = INT64_MAX;
  for (int pass = 0; pass < 2; ++pass) {
    const std::vector<int32_t>& ws = p.iarr[pass == 0 ? I_WSTREAM : I_ESTREAM];
    const std::vector<int32_t>& ptr = p.iarr[pass == 0 ? I_WSTREAM_PTR : I_ESTREAM_PTR];
    if ((int)ptr.size() != SCHUR_WARPS + 1 || ws.size() % (4 * (SCHUR_STAGE + 1)) != 0 ||
        (size_t)ptr.back() * 4 * (SCHUR_STAGE + 1) != ws.size())
      return fail(SWGN_ERR_INVALID, "stream table sizes are inconsistent");
    for (int wv = 0; wv < SCHUR_WARPS; ++wv) {
      const int64_t n_st = ptr[wv + 1] - ptr[wv];
      if (n_st < 0) return fail(SWGN_ERR_INVALID, "stream pointers are not monotone");
      if (pass == 0) {
        out[4] = std::min(out[4], n_st);
        out[5] = std::max(out[5], n_st);
      }
      bool open = false;
      int cur_meta = -1;
      for (int64_t s_ = ptr[wv]; s_ < ptr[wv + 1]; ++s_) {
        const int32_t* h = ws.data() + (size_t)s_ * 4 * (SCHUR_STAGE + 1);
        const int meta = h[3], ps = meta & 63, qs = (meta >> 6) & 63, ti = ((meta >> 12) & 7) * 8, tj = ((meta >> 15) & 7) * 8;
        const int diag = (meta >> 18) & 1, ecell = (h[2] >> 1) & 1;
        if (ecell != pass || ps < 1 || qs < 1 || ti >= ps || tj >= qs + diag) return fail(SWGN_ERR_INVALID, "bad stage header");
        if (open && meta != cur_meta) return fail(SWGN_ERR_INVALID, "a tile changed before its end flag");
        cur_meta = meta;
        open = true;
        out[pass == 0 ? 0 : 6] += 1;
        for (int e = 0; e < SCHUR_STAGE; ++e) {
          const int32_t* t = h + 4 * (1 + e);
          if (t[0] < 0) {
            if (pass == 0) out[2] += 1;
            continue;
          }
          const int rows = ((t[0] >> 28) & 3) + 1;
          const int64_t a = t[0] & 0x0fffffff, bb = t[1], b2 = t[2];
          if (a + (int64_t)rows * ps > jw || bb + (int64_t)rows * qs > jw || (diag && b2 + rows > jw)) return fail(SWGN_ERR_INVALID, "operand outside the window");
          out[pass == 0 ? 1 : 7] += 1;
        }
        if (h[2] & 1) {
          open = false;
          if (pass == 0) {
            out[3] += 1;
            const int64_t last = (int64_t)h[0] + (int64_t)(std::min(ps, ti + 8) - 1) * d.ld + std::min(qs, tj + 8) - 1;
            if (h[0] < 0 || last >= (int64_t)d.n_f * d.ld || h[1] < 0 || h[1] + ps > d.n_f) return fail(SWGN_ERR_INVALID, "tile store outside S");
          } else {
            const int64_t lim = diag ? p.wsize[W_EFAC] : p.wsize[W_EBUF];
            if (h[0] < 0 || (int64_t)h[0] + (int64_t)ps * qs > lim) return fail(SWGN_ERR_INVALID, "e-cell store outside its buffer");
          }
        }
      }
      if (open) return fail(SWGN_ERR_INVALID, "a stream ends inside a tile");
    }
  }
  if (out[4] == INT64_MAX) out[4] = 0;
  return SWGN_OK;
}

swgn_status swgn_plan_stream_info(const swgn_graph* g, int32_t n_parameter_head, int32_t* info) {
  if (!g || !info) return fail(SWGN_ERR_INVALID, "bad arguments");
  WindowPlan p;
  p.want_stream_plan = true;
  std::string err;
  swgn_status st = build_plan(g, n_parameter_head, &p, &err);
  if (st != SWGN_OK) return fail(st, err);
  const StreamPlanInfo& sb = p.sb;
  std::memset(info, 0, sizeof(int32_t) * 16);
  const int32_t v[10] = {sb.fits, sb.nbatch, sb.acc, sb.jcap, sb.rcap, sb.ecap, sb.fcap, sb.reccap, sb.n_fb,
                         (int32_t)stream_smem_bytes(sb.nbatch, sb.acc, sb.jcap, sb.rcap, sb.ecap, sb.fcap, sb.reccap)};
  std::memcpy(info, v, sizeof(v));
  info[10] = sb.ok;  // fits AND enabled (SWGN_SCHUR_STREAM=1)
  return SWGN_OK;
}

swgn_status swgn_plan_array(const swgn_graph* g, int32_t n_parameter_head, int32_t array, int32_t* out, int64_t* n) {
  if (!g || !n || array < 0 || array >= NUM_IARR) return fail(SWGN_ERR_INVALID, "bad arguments");
  WindowPlan p;
  p.want_stream_plan = true;
  std::string err;
  swgn_status st = build_plan(g, n_parameter_head, &p, &err);
  if (st != SWGN_OK) return fail(st, err);
  *n = (int64_t)p.iarr[array].size();
  if (out) std::copy(p.iarr[array].begin(), p.iarr[array].end(), out);
  return SWGN_OK;
}

int32_t swgn_batch_size(const swgn_batch* b) { return b ? b->n : 0; }

swgn_status swgn_batch_set_state(swgn_batch* b, int32_t w, const double* state) {
  if (!b || w < 0 || w >= b->n || !state) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const WinDesc& d = b->desc[w];
  CU(cudaMemcpyAsync(b->d_wpool + d.woff[W_X], state, sizeof(double) * d.n_state, cudaMemcpyHostToDevice, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  return SWGN_OK;
}

swgn_status swgn_batch_get_state(swgn_batch* b, int32_t w, double* state) {
  if (!b || w < 0 || w >= b->n || !state) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const WinDesc& d = b->desc[w];
  CU(cudaMemcpyAsync(state, b->d_wpool + d.woff[W_X], sizeof(double) * d.n_state, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  return SWGN_OK;
}

// New measurements and initial states on an unchanged structure (the replayed-sequence case):
// repack the factor constants and states of every window from the caller's graphs and upload them
// chunk by chunk from pinned staging into (d_cpool, d_stage) on `stream`.
static swgn_status pack_and_upload(swgn_batch* b, const swgn_graph* const* graphs, double* h_cpool, double* h_stage, double* d_cpool,
                                   double* d_stage, cudaStream_t stream) {
  // windows are packed by worker threads in index order while this thread uploads every chunk of
  // consecutive windows as soon as all of its windows are packed (the pools are window-major, so a chunk is
  // one contiguous range): packing and the H2D copy overlap
  const int n_chunks = b->n >= 64 ? 8 : 1;
  const int per_chunk = (b->n + n_chunks - 1) / n_chunks;
  std::vector<std::atomic<int>> chunk_done(n_chunks);
  for (auto& c : chunk_done) c.store(0);
  std::atomic<int> next(0), bad(0);
  auto work = [&]() {
    for (;;) {
      const int w = next.fetch_add(1);
      if (w >= b->n) break;
      const swgn_graph* g = graphs[w];
      const WinDesc& d = b->desc[w];
      int64_t sizes[NUM_CARR];
      if (g) constant_sizes(g, sizes);
      if (!g) {
        bad.store(1);
        chunk_done[w / per_chunk].fetch_add(1);
        continue;
      }
      bool ok = g->n_state == d.n_state && g->n_proj == d.n_proj && g->n_imu == d.n_imu && g->n_gnss == d.n_gnss &&
                g->n_prior == d.n_prior && g->n_unit == d.n_unit && g->n_chain == d.n_chain;
      if (ok) ok = structure_fingerprint(g) == b->fingerprint[w];
      if (ok && g->n_chain > 0) ok = g->chain_frame_begin[g->n_chain] == d.n_chain_frames;
      for (int a = 0; ok && a + 1 < NUM_CARR; ++a) ok = d.coff[a] + sizes[a] <= d.coff[a + 1];
      if (ok && w + 1 < b->n) ok = d.coff[NUM_CARR - 1] + sizes[NUM_CARR - 1] <= b->desc[w + 1].coff[0];
      if (ok && w + 1 == b->n) ok = d.coff[NUM_CARR - 1] + sizes[NUM_CARR - 1] <= (int64_t)b->cpool_n;
      if (!ok) {
        bad.store(1);
        chunk_done[w / per_chunk].fetch_add(1);
        continue;
      }
      double* ptr[NUM_CARR];
      for (int a = 0; a < NUM_CARR; ++a) ptr[a] = h_cpool + d.coff[a];
      pack_constants(g, ptr);
      std::memcpy(h_stage + b->state_off[w], g->state, sizeof(double) * d.n_state);
      chunk_done[w / per_chunk].fetch_add(1, std::memory_order_release);
    }
  };
  {
    const int nt = std::max(1, std::min<int>(b->n / 16, (int)std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work);
    cudaError_t ce = cudaSuccess;
    for (int c = 0; c < n_chunks; ++c) {
      const int w0 = c * per_chunk, w1 = std::min(b->n, w0 + per_chunk);
      if (w0 >= w1) break;
      while (chunk_done[c].load(std::memory_order_acquire) < w1 - w0) std::this_thread::yield();
      if (bad.load() || ce != cudaSuccess) continue;
      const size_t c0 = (size_t)b->desc[w0].coff[0], c1 = w1 < b->n ? (size_t)b->desc[w1].coff[0] : b->cpool_n;
      ce = cudaMemcpyAsync(d_cpool + c0, h_cpool + c0, sizeof(double) * (c1 - c0), cudaMemcpyHostToDevice, stream);
      const int64_t s0 = b->state_off[w0], s1 = b->state_off[w1];
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_stage + s0, h_stage + s0, sizeof(double) * (size_t)(s1 - s0), cudaMemcpyHostToDevice, stream);
    }
    for (auto& t : th) t.join();
    if (ce != cudaSuccess) {
      cudaStreamSynchronize(stream);
      CU(ce);
    }
  }
  if (bad.load()) {
    cudaStreamSynchronize(stream);
    return fail(SWGN_ERR_INVALID, "graph structure differs from the one the batch was created with");
  }
  return SWGN_OK;
}

swgn_status swgn_batch_update_inputs(swgn_batch* b, const swgn_graph* const* graphs, int64_t* bytes_h2d) {
  if (!b || !graphs) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  if (!b->h_cpool) CU(cudaMallocHost(&b->h_cpool, sizeof(double) * std::max<size_t>(b->cpool_n, 2)));
  const swgn_status st = pack_and_upload(b, graphs, b->h_cpool, b->h_stage, b->d_cpool, b->d_stage, b->stream);
  if (st != SWGN_OK) return st;
  const int64_t ns = b->state_off[b->n];
  launch_gather_states(b->db, b->d_stage, b->d_state_off, 1, b->stream);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(b->stream));
  b->db.chain_epoch += 1;  // IMUGNSSFactor chains reload their hidden states and forget their history
  if (bytes_h2d) *bytes_h2d = (int64_t)(sizeof(double) * (b->cpool_n + (size_t)ns));
  return SWGN_OK;
}

// Double-buffered inputs for a replayed sequence: the next step's constants and states are packed and uploaded into a
// SHADOW block [constants | packed states] on a private copy stream while the current solve runs on the batch's stream;
// commit makes that block the live one (same layout, WinDesc::coff) and scatters its states.  Two blocks alternate; the
// pool inside the batch's slab is only live until the first commit.
swgn_status swgn_batch_prefetch_inputs(swgn_batch* b, const swgn_graph* const* graphs, int64_t* bytes_h2d) {
  if (!b || !graphs) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const size_t nc = std::max<size_t>(b->cpool_n, 2), ns = (size_t)std::max<int64_t>(b->state_off[b->n], 2);
  if (!b->copy_stream) {
    CU(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&b->ev_prefetch, cudaEventDisableTiming));
    CU(cudaMallocHost(&b->h_stage2, sizeof(double) * ns));
  }
  if (!b->h_cpool) CU(cudaMallocHost(&b->h_cpool, sizeof(double) * nc));
  const int target = b->live_block == 0 ? 1 : 0;
  if (!b->blocks[target]) CU(cudaMalloc(&b->blocks[target], sizeof(double) * (nc + ns)));
  CU(cudaStreamSynchronize(b->copy_stream));  // the pinned staging of the previous prefetch is free again
  b->prefetched = -1;
  const swgn_status st = pack_and_upload(b, graphs, b->h_cpool, b->h_stage2, b->blocks[target], b->blocks[target] + nc, b->copy_stream);
  if (st != SWGN_OK) return st;
  CU(cudaEventRecord(b->ev_prefetch, b->copy_stream));
  b->prefetched = target;
  if (bytes_h2d) *bytes_h2d = (int64_t)(sizeof(double) * (b->cpool_n + (size_t)b->state_off[b->n]));
  return SWGN_OK;
}

swgn_status swgn_batch_commit_inputs(swgn_batch* b) {
  if (!b) return fail(SWGN_ERR_INVALID, "bad arguments");
  if (b->prefetched < 0) return fail(SWGN_ERR_INVALID, "no prefetched inputs: call swgn_batch_prefetch_inputs first");
  CU(cudaSetDevice(b->device));
  const size_t nc = std::max<size_t>(b->cpool_n, 2);
  CU(cudaStreamWaitEvent(b->stream, b->ev_prefetch, 0));
  b->live_block = b->prefetched;
  b->prefetched = -1;
  b->d_cpool = b->blocks[b->live_block];
  b->db.cpool = b->d_cpool;
  launch_gather_states(b->db, b->d_cpool + nc, b->d_state_off, 1, b->stream);
  CU(cudaGetLastError());
  b->db.chain_epoch += 1;  // IMUGNSSFactor chains reload their hidden states and forget their history
  return SWGN_OK;
}

// development aid: phase timestamps (clock64) of the last k_schur launch (8 per window), then the
// accumulated phase cycles of the last k_chol launch (8 per window)
swgn_status swgn_batch_debug_timeline(swgn_batch* b, int64_t* out) {
  if (!b || !b->d_debug || !out) return fail(SWGN_ERR_INVALID, "timeline not enabled (SWGN_DEBUG_TIMELINE=1)");
  CU(cudaMemcpy(out, b->d_debug, sizeof(long long) * 24 * b->n, cudaMemcpyDeviceToHost));
  return SWGN_OK;
}

int64_t swgn_batch_states_size(const swgn_batch* b) { return b ? b->state_off[b->n] : 0; }

swgn_status swgn_batch_set_states(swgn_batch* b, const double* states) {
  if (!b || !states) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const int64_t n = b->state_off[b->n];
  std::memcpy(b->h_stage, states, sizeof(double) * n);
  CU(cudaMemcpyAsync(b->d_stage, b->h_stage, sizeof(double) * n, cudaMemcpyHostToDevice, b->stream));
  launch_gather_states(b->db, b->d_stage, b->d_state_off, 1, b->stream);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(b->stream));
  return SWGN_OK;
}

swgn_status swgn_batch_get_states(swgn_batch* b, double* states) {
  if (!b || !states) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const int64_t n = b->state_off[b->n];
  launch_gather_states(b->db, b->d_stage, b->d_state_off, 0, b->stream);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(b->h_stage, b->d_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  std::memcpy(states, b->h_stage, sizeof(double) * n);
  return SWGN_OK;
}

static cudaEvent_t get_event(swgn_batch* b, size_t i) {
  while (b->ev.size() <= i) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;  // (recording on a null event fails with a CUDA error the caller reports)
    b->ev.push_back(e);
  }
  return b->ev[i];
}

// Evaluate the host-evaluated residual blocks of every window at the point the next k_eval launch will use (W_X, or
// W_XCAND for a candidate evaluation) with the application's own cost functions and upload the records.  Two host round
// trips per trust-region iteration: the price of cost functions the device has no kind for.
static swgn_status host_evaluate(swgn_batch* b, int state_arr, int only_window) {
  if (!b->any_host) return SWGN_OK;
  cudaStream_t s = b->stream;
  for (int w = 0; w < b->n; ++w) {
    if (only_window >= 0 && w != only_window) continue;
    HostFactors& h = b->host[w];
    if (h.nres.empty()) continue;
    const WinDesc& d = b->desc[w];
    b->h_hoststate.resize((size_t)d.n_state);
    CU(cudaMemcpyAsync(b->h_hoststate.data(), b->d_wpool + d.woff[state_arr], sizeof(double) * d.n_state, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    b->h_hostbuf.assign((size_t)h.buf_doubles, 0.0);
    int64_t off = 0;
    std::vector<const double*> params;
    std::vector<double*> jac;
    for (size_t i = 0; i < h.nres.size(); ++i) {
      params.clear();
      jac.clear();
      double* r = b->h_hostbuf.data() + off;
      int64_t jo = off + h.nres[i];
      for (int k = h.blk_begin[i]; k < h.blk_begin[i + 1]; ++k) {
        params.push_back(b->h_hoststate.data() + h.blk_state[k]);
        jac.push_back(b->h_hostbuf.data() + jo);
        jo += (int64_t)h.nres[i] * h.blk_size[k];
      }
      if (h.fn(h.user, (int32_t)i, params.data(), r, jac.data()) != 0) {
        const double nanv = std::nan("");  // an evaluation failure reads as a non-finite residual: Ceres' "evaluation failed"
        r[0] = nanv;
      }
      off = jo;
    }
    CU(cudaMemcpyAsync(b->d_wpool + d.woff[W_HOSTBUF], b->h_hostbuf.data(), sizeof(double) * (size_t)h.buf_doubles, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
  }
  return SWGN_OK;
}

swgn_status swgn_batch_solve(swgn_batch* b, swgn_summary* summaries) {
  if (!b) return fail(SWGN_ERR_INVALID, "null batch");
  CU(cudaSetDevice(b->device));
  cudaStream_t s = b->stream;
  const DeviceBatch& db = b->db;
  const int max_iter = b->opt.max_num_iterations;
  // every iteration can retry its linear solve while mu < 1 (mu starts at 1e-12, x10 per retry)
  const int tick_limit = (max_iter + 2) * 16;
  int launches = 0, n_schur = 0;
  CU(cudaEventRecord(get_event(b, 0), s));
  const int eval_launches = db.max_chain > 0 ? 2 : 1;  // k_chain runs ahead of k_eval when chains exist
  {
    const swgn_status hs = host_evaluate(b, W_X, -1);
    if (hs != SWGN_OK) return hs;
  }
  launch_eval(db, EVAL_INIT, RUN_STATE_MACHINE, s);
  launches += eval_launches;
  b->h_counters[0] = b->h_counters[1] = -1;
  bool done = false;
  int tick = 0;
  for (int q = 0; q < 2; ++q)
    if (!b->ev_count[q]) CU(cudaEventCreateWithFlags(&b->ev_count[q], cudaEventDisableTiming));
  cudaEvent_t* ev_count = b->ev_count;
  for (; tick < tick_limit && !done; ++tick) {
    const int slot = tick & 1;
    // the count of active windows of tick - 2 has usually reached the host by now: when it is zero, no later tick has
    // anything to do (checked without waiting; a solve that converges early stops issuing launches)
    if (tick >= 2 && tick < max_iter && cudaEventQuery(ev_count[slot]) == cudaSuccess && b->h_counters[slot] == 0) {
      done = true;
      break;
    }
    CU(cudaMemsetAsync(b->d_counters + slot, 0, sizeof(int32_t), s));
    launch_begin(db, tick, s);
    ++launches;
    CU(cudaMemcpyAsync(b->h_counters + slot, b->d_counters + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(ev_count[slot], s));
    if (tick >= max_iter) {
      // no window can still be running before tick max_iter unless it converged early; from
      // here on wait for the count of active windows
      CU(cudaStreamSynchronize(s));
      if (b->h_counters[slot] == 0) {
        done = true;
        break;
      }
    }
    CU(cudaEventRecord(get_event(b, 2 + 2 * n_schur), s));
    launch_schur(db, RUN_STATE_MACHINE, s);
    CU(cudaEventRecord(get_event(b, 3 + 2 * n_schur), s));
    ++n_schur;
    launch_chol(db, RUN_STATE_MACHINE, s);
    launch_backsub(db, RUN_STATE_MACHINE, s);
    launch_step(db, s);
    if (b->any_host) {
      const swgn_status hs = host_evaluate(b, W_XCAND, -1);
      if (hs != SWGN_OK) return hs;
    }
    launch_eval(db, EVAL_CANDIDATE, RUN_STATE_MACHINE, s);
    launch_end(db, s);
    if (b->any_host) {
      const swgn_status hs = host_evaluate(b, W_X, -1);  // (a rejected step left W_X where it was: same values, same records)
      if (hs != SWGN_OK) return hs;
    }
    launch_eval(db, EVAL_ACCEPTED, RUN_STATE_MACHINE, s);
    launches += 5 + 2 * eval_launches;
  }
  launch_finish(db, s);
  ++launches;
  CU(cudaEventRecord(get_event(b, 1), s));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(b->h_state.data(), b->d_state, sizeof(TRState) * b->n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (!done && tick >= tick_limit) return fail(SWGN_ERR_CUDA, "tick limit reached with active windows");
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]));
  b->total_ms = ms;
  b->schur_ms = 0;
  for (int i = 0; i < n_schur; ++i) {
    CU(cudaEventElapsedTime(&ms, b->ev[2 + 2 * i], b->ev[3 + 2 * i]));
    b->schur_ms += ms;
  }
  b->schur_launches = n_schur;
  b->kernel_launches = launches;
  b->solved = true;
  if (summaries) {
    for (int w = 0; w < b->n; ++w) {
      const TRState& t = b->h_state[w];
      swgn_summary& m = summaries[w];
      m.initial_cost = t.initial_cost;
      m.final_cost = t.final_cost;
      m.fixed_cost = t.fixed_cost;
      m.num_successful_steps = t.num_successful;
      m.num_unsuccessful_steps = t.num_unsuccessful;
      m.num_iterations = t.iteration;
      m.num_linear_solves = t.num_linear_solves;
      m.termination_type = t.termination;
      m.n_e = b->desc[w].n_e;
      m.n_f = b->desc[w].n_f;
      m.n_residuals = b->desc[w].n_res;
    }
  }
  return SWGN_OK;
}

swgn_status swgn_batch_last_timing(const swgn_batch* b, double* total_ms, double* schur_ms, int32_t* schur_launches,
                                   int32_t* kernel_launches) {
  if (!b || !b->solved) return fail(SWGN_ERR_INVALID, "no solve has run");
  if (total_ms) *total_ms = b->total_ms;
  if (schur_ms) *schur_ms = b->schur_ms;
  if (schur_launches) *schur_launches = b->schur_launches;
  if (kernel_launches) *kernel_launches = b->kernel_launches;
  return SWGN_OK;
}

// algorithmic bytes of one Schur elimination of window w (SURVEY.md 8d formula)
int64_t swgn_batch_schur_bytes(const swgn_batch* b, int32_t w) {
  if (!b || w < 0 || w >= b->n) return 0;
  return 8 * b->schur_doubles[w];
}

swgn_status swgn_batch_get_reduced(swgn_batch* b, int32_t w, double* S, double* r, int32_t* n) {
  if (!b || w < 0 || w >= b->n || !n) return fail(SWGN_ERR_INVALID, "bad arguments");
  const WinDesc& d = b->desc[w];
  *n = d.n_f;
  if (!S && !r) return SWGN_OK;
  CU(cudaSetDevice(b->device));
  TRState t;
  CU(cudaMemcpy(&t, b->d_state + w, sizeof(t), cudaMemcpyDeviceToHost));
  int src = -1;
  if (t.have_reduced) src = W_S;                  // export-mode solve: S was not factorised
  else if (b->has_scopy && b->db.keep_copy) src = W_SCOPY;  // staged linear solve kept a copy
  if (src < 0) return fail(SWGN_ERR_INVALID, "no reduced system available: run an export-mode solve (is_optimize = 0) or swgn_batch_linear_solve");
  std::vector<double> h((size_t)d.n_f * d.ld);
  CU(cudaMemcpy(h.data(), b->d_wpool + d.woff[src], sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  for (int i = 0; i < d.n_f; ++i) {
    if (S)
      for (int j = 0; j < d.n_f; ++j) S[(size_t)i * d.n_f + j] = (j >= i) ? h[(size_t)i * d.ld + j] : 0.0;
    if (r) r[i] = h[(size_t)i * d.ld + d.n_f];
  }
  return SWGN_OK;
}

swgn_status swgn_batch_get_cholesky(swgn_batch* b, int32_t w, double* L, int32_t* n) {
  if (!b || w < 0 || w >= b->n || !n) return fail(SWGN_ERR_INVALID, "bad arguments");
  const WinDesc& d = b->desc[w];
  *n = d.n_f;
  if (!L) return SWGN_OK;
  CU(cudaSetDevice(b->device));
  TRState t;
  CU(cudaMemcpy(&t, b->d_state + w, sizeof(t), cudaMemcpyDeviceToHost));
  if (!t.have_factor) return fail(SWGN_ERR_INVALID, "no Cholesky factor available (last reduced solve failed or export mode)");
  std::vector<double> h((size_t)d.n_f * d.ld);
  CU(cudaMemcpy(h.data(), b->d_wpool + d.woff[W_S], sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  for (int i = 0; i < d.n_f; ++i)
    for (int j = 0; j < d.n_f; ++j) L[(size_t)i * d.n_f + j] = (j <= i) ? h[(size_t)j * d.ld + i] : 0.0;
  return SWGN_OK;
}

swgn_status swgn_batch_get_tail_information(swgn_batch* b, int32_t w, int32_t n_tail, double* A) {
  if (!b || w < 0 || w >= b->n || !A || n_tail <= 0 || n_tail > b->desc[w].n_f) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  TRState t;
  CU(cudaMemcpy(&t, b->d_state + w, sizeof(t), cudaMemcpyDeviceToHost));
  if (!t.have_factor) return fail(SWGN_ERR_INVALID, "no Cholesky factor available");
  double* dA = nullptr;
  CU(cudaMalloc(&dA, sizeof(double) * n_tail * n_tail));
  launch_tail_information(b->db, w, n_tail, dA, b->stream);
  cudaError_t e = cudaMemcpyAsync(A, dA, sizeof(double) * n_tail * n_tail, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  cudaFree(dA);
  CU(e);
  return SWGN_OK;
}

swgn_status swgn_batch_ambiguity_fix(swgn_batch* b, int32_t n_tail, const int32_t* win_epoch, const int32_t* epoch_begin,
                                     const int32_t* obs_amb, const int32_t* obs_sysfreq, const int32_t* last_fix,
                                     swgn_fix_result* results, int32_t* dd_pairs, double* F) {
  if (!b || n_tail <= 0 || !win_epoch || !epoch_begin || !obs_amb || !obs_sysfreq || !results) return fail(SWGN_ERR_INVALID, "bad arguments");
  const int nw = b->n;
  if (win_epoch[0] < 0) return fail(SWGN_ERR_INVALID, "bad epoch list");
  for (int w = 0; w < nw; ++w)
    if (win_epoch[w + 1] < win_epoch[w] + 1) return fail(SWGN_ERR_INVALID, "every window needs at least one epoch");
  const int n_ep = win_epoch[nw];
  for (int e = 0; e < n_ep; ++e)
    if (epoch_begin[e + 1] < epoch_begin[e] || epoch_begin[e] < 0) return fail(SWGN_ERR_INVALID, "epoch offsets are not monotone");
  const int n_obs = epoch_begin[n_ep];
  CU(cudaSetDevice(b->device));
  cudaStream_t s = b->stream;
  const size_t n = (size_t)n_tail, nwk = fix_work_doubles(n_tail), niw = fix_work_ints(n_tail);
  // one device block: A | y | F | work (doubles), then results, then the int arrays
  const size_t n_dbl = (size_t)nw * (n * n + n + 2 * n + nwk);
  const size_t n_int = (size_t)nw * (1 + 2 * n + niw + 1) + (size_t)(nw + 1) + (size_t)(n_ep + 1) + 2 * (size_t)(n_obs + 1);
  const size_t bytes = sizeof(double) * n_dbl + sizeof(swgn_fix_result) * (size_t)nw + sizeof(int32_t) * n_int + 64;
  char* dbuf = nullptr;
  swgn::keep_pool_memory(b->device);
  CU(cudaMallocAsync((void**)&dbuf, bytes, s));  // stream-ordered: no device-wide synchronisation per call
  double* dA = reinterpret_cast<double*>(dbuf);
  double* dy = dA + (size_t)nw * n * n;
  double* dF = dy + (size_t)nw * n;
  double* dwork = dF + (size_t)nw * 2 * n;
  swgn_fix_result* dres = reinterpret_cast<swgn_fix_result*>(dwork + (size_t)nw * nwk);
  int32_t* dhave = reinterpret_cast<int32_t*>(dres + nw);
  int32_t* dpairs = dhave + nw;
  int32_t* diwork = dpairs + (size_t)nw * 2 * n;
  int32_t* dlast = diwork + (size_t)nw * niw;
  int32_t* dwin = dlast + nw;
  int32_t* deb = dwin + (nw + 1);
  int32_t* doa = deb + (n_ep + 1);
  int32_t* dsf = doa + (n_obs + 1);
  cudaError_t e = cudaMemcpyAsync(dwin, win_epoch, sizeof(int32_t) * (nw + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(deb, epoch_begin, sizeof(int32_t) * (n_ep + 1), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && n_obs > 0) e = cudaMemcpyAsync(doa, obs_amb, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && n_obs > 0) e = cudaMemcpyAsync(dsf, obs_sysfreq, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = last_fix ? cudaMemcpyAsync(dlast, last_fix, sizeof(int32_t) * nw, cudaMemcpyHostToDevice, s)
                                     : cudaMemsetAsync(dlast, 0, sizeof(int32_t) * nw, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(dF, 0, sizeof(double) * (size_t)nw * 2 * n, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(dpairs, 0, sizeof(int32_t) * (size_t)nw * 2 * n, s);
  if (e == cudaSuccess) {
    launch_tail_information_batch(b->db, n_tail, dA, dy, dhave, s);
    launch_ambiguity_fix_batch(nw, n_tail, dA, dy, dwin, deb, doa, dsf, dlast, dhave, dpairs, dF, dres, dwork, diwork, s);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(results, dres, sizeof(swgn_fix_result) * (size_t)nw, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && dd_pairs) e = cudaMemcpyAsync(dd_pairs, dpairs, sizeof(int32_t) * (size_t)nw * 2 * n, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && F) e = cudaMemcpyAsync(F, dF, sizeof(double) * (size_t)nw * 2 * n, cudaMemcpyDeviceToHost, s);
  cudaFreeAsync(dbuf, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  CU(e);
  return SWGN_OK;
}

swgn_status swgn_preintegrate_batch(int32_t device, int32_t n_factors, const int32_t* sample_begin, const double* samples,
                                    const double* bias, const double noise[4], double* records, int32_t* info) {
  if (n_factors <= 0 || !sample_begin || !samples || !bias || !noise || !records) return fail(SWGN_ERR_INVALID, "bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(SWGN_ERR_NO_DEVICE, "no usable CUDA device (the solver has no CPU fallback)");
  for (int f = 0; f < n_factors; ++f)
    if (sample_begin[f + 1] - sample_begin[f] < 1) return fail(SWGN_ERR_INVALID, "every factor needs at least its initial sample");
  CU(cudaSetDevice(device));
  const size_t ns = (size_t)sample_begin[n_factors];
  int32_t *d_begin = nullptr, *d_info = nullptr;
  double *d_samples = nullptr, *d_bias = nullptr, *d_rec = nullptr;
  cudaError_t e = cudaMalloc(&d_begin, sizeof(int32_t) * (n_factors + 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_info, sizeof(int32_t) * n_factors);
  if (e == cudaSuccess) e = cudaMalloc(&d_samples, sizeof(double) * 7 * ns);
  if (e == cudaSuccess) e = cudaMalloc(&d_bias, sizeof(double) * 6 * n_factors);
  if (e == cudaSuccess) e = cudaMalloc(&d_rec, sizeof(double) * (size_t)SWGN_IMU_STRIDE * n_factors);
  if (e == cudaSuccess) e = cudaMemcpy(d_begin, sample_begin, sizeof(int32_t) * (n_factors + 1), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_samples, samples, sizeof(double) * 7 * ns, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_bias, bias, sizeof(double) * 6 * n_factors, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_preintegrate(n_factors, d_begin, d_samples, d_bias, noise, d_rec, d_info, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(records, d_rec, sizeof(double) * (size_t)SWGN_IMU_STRIDE * n_factors, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && info) e = cudaMemcpy(info, d_info, sizeof(int32_t) * n_factors, cudaMemcpyDeviceToHost);
  cudaFree(d_begin); cudaFree(d_info); cudaFree(d_samples); cudaFree(d_bias); cudaFree(d_rec);
  CU(e);
  return SWGN_OK;
}

swgn_status swgn_batch_get_head_marginal(swgn_batch* b, int32_t w, int32_t n_tail, double* A, double* bvec) {
  if (!b || w < 0 || w >= b->n || !A || !bvec || n_tail <= 0 || n_tail > b->desc[w].n_f) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  TRState t;
  CU(cudaMemcpy(&t, b->d_state + w, sizeof(t), cudaMemcpyDeviceToHost));
  if (!t.have_reduced) return fail(SWGN_ERR_INVALID, "no reduced system available: run an export-mode solve (is_optimize = 0) first");
  const int nf = b->desc[w].n_f, m = nf - n_tail;
  double* dbuf = nullptr;
  const size_t na = (size_t)n_tail * n_tail, ns = head_marginal_scratch_doubles(m, n_tail);
  CU(cudaMalloc(&dbuf, sizeof(double) * (na + n_tail + ns)));
  cudaError_t e = launch_head_marginal(b->db, w, nf, n_tail, dbuf, dbuf + na, dbuf + na + n_tail, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(A, dbuf, sizeof(double) * na, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(bvec, dbuf + na, sizeof(double) * n_tail, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  cudaFree(dbuf);
  CU(e);
  return SWGN_OK;
}

swgn_status swgn_batch_get_marginal_prior(swgn_batch* b, int32_t w, int32_t n_tail, double* J0, double* r0, double* A, double* bvec) {
  if (!b || w < 0 || w >= b->n || !J0 || !r0 || n_tail <= 0 || n_tail > b->desc[w].n_f) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  TRState t;
  CU(cudaMemcpy(&t, b->d_state + w, sizeof(t), cudaMemcpyDeviceToHost));
  if (!t.have_reduced) return fail(SWGN_ERR_INVALID, "no reduced system available: run an export-mode solve (is_optimize = 0) first");
  const int nf = b->desc[w].n_f, m = nf - n_tail;
  const size_t na = (size_t)n_tail * n_tail;
  const size_t ns = std::max(head_marginal_scratch_doubles(m, n_tail), prior_sqrt_scratch_doubles(n_tail));
  double* dbuf = nullptr;  // A | b | J0 | r0 | scratch
  CU(cudaMalloc(&dbuf, sizeof(double) * (2 * na + 2 * (size_t)n_tail + ns)));
  double *dA = dbuf, *db = dbuf + na, *dJ = db + n_tail, *dr = dJ + na, *dscr = dr + n_tail;
  cudaError_t e = launch_head_marginal(b->db, w, nf, n_tail, dA, db, dscr, b->stream);
  if (e == cudaSuccess) e = launch_prior_sqrt(dA, db, n_tail, dJ, dr, dscr, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(J0, dJ, sizeof(double) * na, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r0, dr, sizeof(double) * n_tail, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess && A) e = cudaMemcpyAsync(A, dA, sizeof(double) * na, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess && bvec) e = cudaMemcpyAsync(bvec, db, sizeof(double) * n_tail, cudaMemcpyDeviceToHost, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  cudaFree(dbuf);
  CU(e);
  return SWGN_OK;
}

}  // extern "C"
namespace swgn {
// UpdateSchur + setmarginalizeinfo of every window in two launches; (J0, r0) of window w land in J0_ptr[w] / r0_ptr[w]
// (windows with n_tail[w] = 0 are skipped).  The results come back through a pinned staging slab (parked in the slab cache
// between calls) in one copy and are scattered from there: the callers' buffers are pageable and separate per window.
swgn_status batch_marginal_priors_to(swgn_batch* b, const int32_t* n_tail, double* const* J0_ptr, double* const* r0_ptr) {
  if (!b || !n_tail || !J0_ptr || !r0_ptr) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const bool dbg_t = std::getenv("SWGN_DEBUG_TIMING") != nullptr;
  auto t_mark = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!dbg_t) return;
    const auto t = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[swgn_batch_get_marginal_priors] %-18s %6.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_mark).count());
    t_mark = t;
  };
  std::vector<TRState> t(b->n);
  CU(cudaMemcpy(t.data(), b->d_state, sizeof(TRState) * b->n, cudaMemcpyDeviceToHost));
  std::vector<int64_t> off((size_t)4 * b->n, 0);
  int64_t n_out = 0, total = 0;  // results (J0 | r0 of every window) first, so that one copy of n_out doubles brings them back
  int max_m = 0, max_n = 0;
  for (int w = 0; w < b->n; ++w) {
    const int n = n_tail[w];
    if (n == 0) continue;
    if (n < 0 || n > b->desc[w].n_f) return fail(SWGN_ERR_INVALID, "bad n_tail");
    if (!t[w].have_reduced) return fail(SWGN_ERR_INVALID, "no reduced system available: run an export-mode solve (is_optimize = 0) first");
    off[4 * w + 2] = n_out;
    n_out += (int64_t)n * n + n;
  }
  total = n_out;
  for (int w = 0; w < b->n; ++w) {
    const int n = n_tail[w];
    if (n == 0) continue;
    const int m = b->desc[w].n_f - n;
    max_m = std::max(max_m, m);
    max_n = std::max(max_n, n);
    off[4 * w] = total;
    total += (int64_t)n * n;
    off[4 * w + 1] = total;
    total += n;
    off[4 * w + 3] = total;
    total += (int64_t)std::max(head_marginal_scratch_doubles(m, n), prior_sqrt_scratch_doubles(n));
  }
  if (total == 0) return SWGN_OK;
  mark("states + offsets");
  double* dbuf = nullptr;
  int64_t* doff = nullptr;
  int32_t* dn = nullptr;
  void* hslab = nullptr;
  size_t hcap = 0;
  // stream-ordered allocations: no device-wide synchronisation per call
  swgn::keep_pool_memory(b->device);
  cudaError_t e = cudaMallocAsync((void**)&dbuf, sizeof(double) * total, b->stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&doff, sizeof(int64_t) * off.size(), b->stream);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&dn, sizeof(int32_t) * b->n, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(doff, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dn, n_tail, sizeof(int32_t) * b->n, cudaMemcpyHostToDevice, b->stream);
  if (e == cudaSuccess) e = launch_marginal_priors(b->db, max_m, max_n, dn, doff, dbuf, b->stream);
  if (dbg_t) {
    mark("alloc + launch");
    cudaStreamSynchronize(b->stream);
    mark("kernels");
  }
  if (e == cudaSuccess) e = slab_alloc(1, b->device, sizeof(double) * (size_t)n_out, &hslab, &hcap);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hslab, dbuf, sizeof(double) * n_out, cudaMemcpyDeviceToHost, b->stream);
  if (dbuf) cudaFreeAsync(dbuf, b->stream);
  if (doff) cudaFreeAsync(doff, b->stream);
  if (dn) cudaFreeAsync(dn, b->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
  if (e != cudaSuccess) {
    slab_free(1, b->device, hslab, hcap);
    CU(e);
  }
  mark("read-back");
  for (int w = 0; w < b->n; ++w) {
    const int n = n_tail[w];
    if (n == 0) continue;
    const double* src = static_cast<const double*>(hslab) + off[4 * w + 2];
    std::memcpy(J0_ptr[w], src, sizeof(double) * (size_t)n * n);
    std::memcpy(r0_ptr[w], src + (size_t)n * n, sizeof(double) * n);
  }
  slab_free(1, b->device, hslab, hcap);
  mark("scatter");
  return SWGN_OK;
}
}  // namespace swgn
extern "C" {

swgn_status swgn_batch_get_marginal_priors(swgn_batch* b, const int32_t* n_tail, const int64_t* j_off, const int64_t* r_off, double* J0_all,
                                           double* r0_all) {
  if (!b || !n_tail || !j_off || !r_off || !J0_all || !r0_all) return fail(SWGN_ERR_INVALID, "bad arguments");
  std::vector<double*> Jp(b->n), rp(b->n);
  for (int w = 0; w < b->n; ++w) {
    Jp[w] = J0_all + j_off[w];
    rp[w] = r0_all + r_off[w];
  }
  return swgn::batch_marginal_priors_to(b, n_tail, Jp.data(), rp.data());
}

swgn_status swgn_batch_get_chain_frames(swgn_batch* b, int32_t w, int32_t* n_frames, double* frames) {
  if (!b || w < 0 || w >= b->n || !n_frames) return fail(SWGN_ERR_INVALID, "bad arguments");
  const WinDesc& d = b->desc[w];
  *n_frames = d.n_chain_frames;
  if (!frames || d.n_chain == 0) return SWGN_OK;
  CU(cudaSetDevice(b->device));
  std::vector<int32_t> rec((size_t)d.n_chain * 8);
  CU(cudaMemcpy(rec.data(), b->d_ipool + d.ioff[I_CHAIN], sizeof(int32_t) * rec.size(), cudaMemcpyDeviceToHost));
  for (int c = 0; c < d.n_chain; ++c) {
    const int m = rec[8 * c], k = rec[8 * c + 1], frame0 = rec[8 * c + 6];
    const ChainLayout L(m, k);
    std::vector<double> fl(4);
    const double* src = b->d_wpool + d.woff[W_CHAIN] + rec[8 * c + 5];
    CU(cudaMemcpy(fl.data(), src + L.w_flags, sizeof(double) * 4, cudaMemcpyDeviceToHost));
    if (fl[1] == (double)b->db.chain_epoch) {  // evaluated at least once since the inputs were loaded
      CU(cudaMemcpy(frames + (size_t)16 * frame0, src + L.w_frames, sizeof(double) * 16 * m, cudaMemcpyDeviceToHost));
    } else {  // untouched: the hidden states are still the uploaded ones
      std::vector<double> fr((size_t)m * CHAIN_FRAME_STRIDE);
      CU(cudaMemcpy(fr.data(), b->d_cpool + d.coff[C_CHAIN] + rec[8 * c + 4] + L.c_frame, sizeof(double) * fr.size(), cudaMemcpyDeviceToHost));
      for (int i = 0; i < m; ++i) std::memcpy(frames + (size_t)16 * (frame0 + i), fr.data() + (size_t)i * CHAIN_FRAME_STRIDE, sizeof(double) * 16);
    }
  }
  return SWGN_OK;
}

// ---- staged entry points ---------------------------------------------------------------------
swgn_status swgn_batch_evaluate(swgn_batch* b, int32_t w, double* cost, double* residuals, double* gradient) {
  if (!b || w < 0 || w >= b->n) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const WinDesc& d = b->desc[w];
  {
    const swgn_status hs = host_evaluate(b, W_X, w);
    if (hs != SWGN_OK) return hs;
  }
  launch_eval(b->db, EVAL_FORCE, w, b->stream);
  CU(cudaGetLastError());
  TRState t;
  CU(cudaMemcpyAsync(&t, b->d_state + w, sizeof(t), cudaMemcpyDeviceToHost, b->stream));
  if (residuals) CU(cudaMemcpyAsync(residuals, b->d_wpool + d.woff[W_RES], sizeof(double) * d.n_res, cudaMemcpyDeviceToHost, b->stream));
  if (gradient) CU(cudaMemcpyAsync(gradient, b->d_wpool + d.woff[W_G], sizeof(double) * d.n_t, cudaMemcpyDeviceToHost, b->stream));
  CU(cudaStreamSynchronize(b->stream));
  if (cost) *cost = t.x_cost;
  return SWGN_OK;
}

static swgn_status fetch_i(swgn_batch* b, int w, int arr, size_t n, std::vector<int32_t>* out) {
  out->resize(n);
  if (n == 0) return SWGN_OK;
  CU(cudaMemcpy(out->data(), b->d_ipool + b->desc[w].ioff[arr], sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  return SWGN_OK;
}

swgn_status swgn_batch_get_columns(swgn_batch* b, int32_t w, int32_t* n_cols, int32_t* block, int32_t* offset, int32_t* size) {
  if (!b || w < 0 || w >= b->n || !n_cols) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const WinDesc& d = b->desc[w];
  *n_cols = d.n_cols;
  std::vector<int32_t> t;
  swgn_status st;
  if (block) { if ((st = fetch_i(b, w, I_COL_BLOCK, d.n_cols, &t))) return st; std::copy(t.begin(), t.end(), block); }
  if (offset) { if ((st = fetch_i(b, w, I_COL_POS, d.n_cols, &t))) return st; std::copy(t.begin(), t.end(), offset); }
  if (size) { if ((st = fetch_i(b, w, I_COL_SIZE, d.n_cols, &t))) return st; std::copy(t.begin(), t.end(), size); }
  return SWGN_OK;
}

swgn_status swgn_batch_get_rows(swgn_batch* b, int32_t w, int32_t* n_rows, int32_t* factor, int32_t* offset) {
  if (!b || w < 0 || w >= b->n || !n_rows) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const WinDesc& d = b->desc[w];
  *n_rows = d.n_rows;
  std::vector<int32_t> t;
  swgn_status st;
  if (factor) { if ((st = fetch_i(b, w, I_ROW_FACTOR, d.n_rows, &t))) return st; std::copy(t.begin(), t.end(), factor); }
  if (offset) { if ((st = fetch_i(b, w, I_ROW_RES, d.n_rows, &t))) return st; std::copy(t.begin(), t.end(), offset); }
  return SWGN_OK;
}

swgn_status swgn_batch_get_dense_jacobian(swgn_batch* b, int32_t w, double* J) {
  if (!b || w < 0 || w >= b->n || !J) return fail(SWGN_ERR_INVALID, "bad arguments");
  CU(cudaSetDevice(b->device));
  const WinDesc& d = b->desc[w];
  std::vector<int32_t> row_res, row_nres, row_cell, cell_col, cell_val, col_pos, col_size;
  swgn_status st;
  if ((st = fetch_i(b, w, I_ROW_RES, d.n_rows, &row_res))) return st;
  if ((st = fetch_i(b, w, I_ROW_NRES, d.n_rows, &row_nres))) return st;
  if ((st = fetch_i(b, w, I_ROW_CELL, d.n_rows + 1, &row_cell))) return st;
  if ((st = fetch_i(b, w, I_CELL_COL, d.n_cells, &cell_col))) return st;
  if ((st = fetch_i(b, w, I_CELL_VAL, d.n_cells, &cell_val))) return st;
  if ((st = fetch_i(b, w, I_COL_POS, d.n_cols, &col_pos))) return st;
  if ((st = fetch_i(b, w, I_COL_SIZE, d.n_cols, &col_size))) return st;
  std::vector<double> v(std::max(d.n_jac, 1));
  CU(cudaMemcpy(v.data(), b->d_wpool + d.woff[W_JAC], sizeof(double) * d.n_jac, cudaMemcpyDeviceToHost));
  std::fill(J, J + (size_t)d.n_res * d.n_t, 0.0);
  for (int r = 0; r < d.n_rows; ++r)
    for (int c = row_cell[r]; c < row_cell[r + 1]; ++c) {
      const int col = cell_col[c], cs = col_size[col];
      for (int rr = 0; rr < row_nres[r]; ++rr)
        for (int k = 0; k < cs; ++k) J[(size_t)(row_res[r] + rr) * d.n_t + col_pos[col] + k] = v[cell_val[c] + rr * cs + k];
    }
  return SWGN_OK;
}

swgn_status swgn_batch_linear_solve(swgn_batch* b, int32_t w, const double* D, double* x) {
  if (!b || w < 0 || w >= b->n || !x) return fail(SWGN_ERR_INVALID, "bad arguments");
  if (!b->has_scopy) return fail(SWGN_ERR_UNSUPPORTED, "staged linear solve is only available for batches of <= 256 windows");
  CU(cudaSetDevice(b->device));
  const WinDesc& d = b->desc[w];
  cudaStream_t s = b->stream;
  if (D) CU(cudaMemcpyAsync(b->d_wpool + d.woff[W_LMD], D, sizeof(double) * d.n_t, cudaMemcpyHostToDevice, s));
  else CU(cudaMemsetAsync(b->d_wpool + d.woff[W_LMD], 0, sizeof(double) * d.n_t, s));
  DeviceBatch db = b->db;
  db.keep_copy = 1;
  db.params.export_mode = 0;
  b->db.keep_copy = 1;  // get_reduced reads the copy from now on
  {
    const swgn_status hs = host_evaluate(b, W_X, w);
    if (hs != SWGN_OK) return hs;
  }
  launch_eval(db, EVAL_FORCE, w, s);
  launch_schur(db, w, s);
  launch_chol(db, w, s);
  launch_backsub(db, w, s);
  CU(cudaGetLastError());
  TRState t;
  CU(cudaMemcpyAsync(&t, b->d_state + w, sizeof(t), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(x, b->d_wpool + d.woff[W_Y], sizeof(double) * d.n_t, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (!t.chol_ok) return fail(SWGN_ERR_INVALID, "reduced system is not positive definite");
  return SWGN_OK;
}

}  // extern "C"
