// Micro-benchmark behind DESIGN.md 6 "next step": how fast can one SM pull small operand slabs
// (64 B .. 4 KB, 16-byte aligned, scattered inside a 2.4 MB per-CTA region like one window's
// Jacobian / E-buffer pool) into the SM
//   (a) with 1-D bulk async copies (cp.async.bulk -> UBLKCP, one lane per warp issues, a ring of D
//       slots per warp, completion on mbarriers, lanes then read the slab from shared memory), or
//   (b) with the per-lane LDG.64 gather k_schur's phase 2 uses today (16 independent loads in flight
//       per lane, one 256-byte row per warp instruction)?
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/ubench_slab_gather.cu -o build_tools/ubench_slab_gather
// Run on the box:  build_tools/ubench_slab_gather > gpurun_out/ubench_slab_gather.txt
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CU(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

constexpr int kWarps = 8;
constexpr int kThreads = 32 * kWarps;
constexpr size_t kRegion = 2400000 / 16 * 16;  // bytes per CTA region

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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// scattered: anywhere in the CTA's region; sequential (seq != 0): the warp walks its eighth of the region
__device__ __forceinline__ uint32_t slab_offset(uint32_t cta, uint32_t warp, uint32_t t, uint32_t bytes, int seq) {
  if (seq) {
    // the warp walks 256 KB of its eighth again and again (a power of two: no division on the address path)
    return warp * (uint32_t)(kRegion / kWarps / 16 * 16) + ((t * bytes) & 0x3ffffu);
  }
  const uint32_t h = mix(cta * 0x9e3779b9u + warp * 0x85ebca6bu + t * 0xc2b2ae35u + 17u);
  return __umulhi(h, (uint32_t)(kRegion - bytes)) & ~15u;
}

// (a) bulk copies: ring of `depth` slots per warp
__global__ void __launch_bounds__(kThreads) k_bulk(const char* __restrict__ pool, double* __restrict__ sink, int bytes, int depth, int terms, int seq) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // kWarps * depth
  unsigned char* slots = smem + ((sizeof(uint64_t) * kWarps * depth + 127) & ~127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const char* region = pool + (size_t)blockIdx.x * kRegion;
  uint64_t* bar = bars + warp * depth;
  unsigned char* slot = slots + (size_t)warp * depth * bytes;
  if (lane == 0)
    for (int d = 0; d < depth; ++d) mbar_init(bar + d, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (lane == 0)
    for (int d = 0; d < depth && d < terms; ++d) {
      mbar_expect_tx(bar + d, bytes);
      bulk_g2s(slot + (size_t)d * bytes, region + slab_offset(blockIdx.x, warp, d, bytes, seq), bytes, bar + d);
    }
  double acc = 0.0;
  const int rows = bytes >= 256 ? bytes / 256 : 1;
  for (int t = 0; t < terms; ++t) {
    const int d = t % depth;
    mbar_wait(bar + d, (t / depth) & 1);
    const double* s = reinterpret_cast<const double*>(slot + (size_t)d * bytes);
    if (bytes >= 256) {
      for (int r = 0; r < rows; ++r) acc += s[r * 32 + lane];
    } else if (lane * 8 < bytes) {
      acc += s[lane];
    }
    __syncwarp();
    if (lane == 0 && t + depth < terms) {
      mbar_expect_tx(bar + d, bytes);
      bulk_g2s(slot + (size_t)d * bytes, region + slab_offset(blockIdx.x, warp, t + depth, bytes, seq), bytes, bar + d);
    }
  }
  if (acc == 123.456) sink[blockIdx.x * kThreads + threadIdx.x] = acc;
}

// (b) direct gather: 16 independent LDG.64 per lane in flight
__global__ void __launch_bounds__(kThreads) k_ldg(const char* __restrict__ pool, double* __restrict__ sink, int bytes, int terms, int seq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const char* region = pool + (size_t)blockIdx.x * kRegion;
  double acc = 0.0;
  const int lrows = bytes >= 256 ? 31 - __clz(bytes / 256) : 0;  // slab sizes are powers of two
  const bool on = bytes >= 256 || lane * 8 < bytes;
  // one "row" = one warp-wide load instruction; walk (term, row) pairs 16 at a time
  const uint32_t total = (uint32_t)terms << lrows;
  for (uint32_t i0 = 0; i0 < total; i0 += 16) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const uint32_t i = i0 + u;
      const uint32_t t = i >> lrows, r = i & ((1u << lrows) - 1);
      const double* s = reinterpret_cast<const double*>(region + slab_offset(blockIdx.x, warp, t, bytes, seq));
      v[u] = (on && i < total) ? __ldg(s + r * 32 + lane) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) acc += v[u];
  }
  if (acc == 123.456) sink[blockIdx.x * kThreads + threadIdx.x] = acc;
}

int main() {
  int dev = 0, sms = 0;
  CU(cudaSetDevice(dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ctas_per_sm = 3;
  const int grid = sms * ctas_per_sm;
  char* pool;
  double* sink;
  CU(cudaMalloc(&pool, (size_t)grid * kRegion));
  CU(cudaMemset(pool, 0, (size_t)grid * kRegion));
  CU(cudaMalloc(&sink, sizeof(double) * grid * kThreads));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  std::printf("# %d SMs, grid %d CTAs (%d per SM) x %d threads, region %.1f MB per CTA (total %.2f GB)\n", sms, grid, ctas_per_sm, kThreads,
              kRegion / 1e6, grid * (double)kRegion / 1e9);
  std::printf("| slab bytes | placement | mode | depth | ms | GB/s | slabs/us | B/cycle/SM @1.965GHz |\n|---|---|---|---|---|---|---|---|\n");
  const int sizes[] = {64, 128, 256, 512, 1024, 2048, 4096};
  for (int seq = 0; seq < 2; ++seq)
  for (int bytes : sizes) {
    const char* pl = seq ? "sequential" : "scattered";
    const int terms = (int)std::min<long long>(65536, (16ll << 20) / bytes);  // per warp
    const double total_bytes = (double)grid * kWarps * terms * bytes;
    const double slabs = (double)grid * kWarps * terms;
    for (int depth : {4, 16}) {
      const size_t sm = ((sizeof(uint64_t) * kWarps * depth + 127) & ~127) + (size_t)kWarps * depth * bytes;
      if (sm * ctas_per_sm > 200 * 1024) continue;
      CU(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      for (int rep = 0; rep < 2; ++rep) {
        CU(cudaEventRecord(e0));
        k_bulk<<<grid, kThreads, sm>>>(pool, sink, bytes, depth, terms, seq);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        CU(cudaGetLastError());
        float ms;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 1)
          std::printf("| %d | %s | bulk | %d | %.3f | %.0f | %.0f | %.1f |\n", bytes, pl, depth, ms, total_bytes / ms / 1e6, slabs / ms / 1e3,
                      total_bytes / (ms * 1e-3) / 1.965e9 / sms);
      }
    }
    for (int rep = 0; rep < 2; ++rep) {
      CU(cudaEventRecord(e0));
      k_ldg<<<grid, kThreads>>>(pool, sink, bytes, terms, seq);
      CU(cudaEventRecord(e1));
      CU(cudaEventSynchronize(e1));
      CU(cudaGetLastError());
      float ms;
      CU(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 1)
        std::printf("| %d | %s | ldg x16 | - | %.3f | %.0f | %.0f | %.1f |\n", bytes, pl, ms, total_bytes / ms / 1e6, slabs / ms / 1e3,
                    total_bytes / (ms * 1e-3) / 1.965e9 / sms);
    }
  }
  return 0;
}
