// Development aid: issue rate and latency of the FP64 tensor-core MMA shapes and of plain DFMA on this GPU.
// One CTA per SM, W warps per CTA, every warp issues N MMAs in `chains` independent accumulator chains.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_dmma.cu -o build_tools/ubench_dmma
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  if (SHAPE == 0)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
  else if (SHAPE == 1)
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  else if (SHAPE == 2)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int CHAINS>
__global__ void k(double* out, long long* cyc, int n) {
  double a[8], b[4], c[CHAINS][4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 1e-4 + i;
  for (int q = 0; q < CHAINS; ++q)
    for (int i = 0; i < 4; ++i) c[q][i] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) mma<SHAPE>(c[q], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < CHAINS; ++q)
    for (int i = 0; i < 4; ++i) s += c[q][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
__global__ void kfma(double* out, long long* cyc, int n) {
  double a = threadIdx.x * 1e-3, b = 1.0000001, c[CHAINS];
  for (int q = 0; q < CHAINS; ++q) c[q] = q;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) c[q] = fma(c[q], b, a);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < CHAINS; ++q) s += c[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int warps, int chains, double flop_per_op) {
  static double* out = nullptr;
  static long long* cyc = nullptr;
  if (!out) { cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8); }
  const int n = 2000;
  launch(out, cyc, n, warps);
  launch(out, cyc, n, warps);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0];
  const double ops = (double)n * chains * warps;
  printf("%-10s warps %2d chains %d: %7.2f cycles per op per SM, %7.1f cycles per op per warp, %6.1f FLOP/clk/SM  (%s)\n", name, warps, chains,
         c / ops, c / ((double)n * chains), flop_per_op * ops / c, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  for (int warps : {1, 4, 8, 16, 32}) {
    run("m8n8k4", [](double* o, long long* c, int n, int w) { k<0, 1><<<148, 32 * w>>>(o, c, n); }, warps, 1, 512);
    run("m8n8k4", [](double* o, long long* c, int n, int w) { k<0, 4><<<148, 32 * w>>>(o, c, n); }, warps, 4, 512);
    run("m16n8k4", [](double* o, long long* c, int n, int w) { k<1, 4><<<148, 32 * w>>>(o, c, n); }, warps, 4, 1024);
    run("m16n8k8", [](double* o, long long* c, int n, int w) { k<2, 4><<<148, 32 * w>>>(o, c, n); }, warps, 4, 2048);
    run("m16n8k16", [](double* o, long long* c, int n, int w) { k<3, 4><<<148, 32 * w>>>(o, c, n); }, warps, 4, 4096);
    run("dfma", [](double* o, long long* c, int n, int w) { kfma<1><<<148, 32 * w>>>(o, c, n); }, warps, 1, 64);
    run("dfma", [](double* o, long long* c, int n, int w) { kfma<8><<<148, 32 * w>>>(o, c, n); }, warps, 8, 64);
  }
  return 0;
}
