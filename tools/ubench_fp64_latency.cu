// Development aid: dependent-chain latency (cycles per link) of the scalar FP64 operations that make up the serial
// chains of k_chol (pivot chain of the Cholesky, triangular back-solve) on this GPU: one warp, one chain.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_fp64_latency.cu -o build_tools/ubench_fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 2048;

template <int OP>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = seed + lane;
  sm[lane + 32] = seed * 0.5;
  __syncthreads();
  double x = seed + 1e-3 * lane, y = 1.0 + 1e-9 * lane;
  const long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);                                   // DFMA
    if (OP == 1) x = x * y;                                             // DMUL
    if (OP == 2) x = x + y;                                             // DADD
    if (OP == 3) x = rsqrt(x) + 1.5;                                    // rsqrt (software sequence) + DADD
    if (OP == 4) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);      // 64-bit shuffle (two SHFL)
    if (OP == 5) x = __shfl_sync(0xffffffffu, x * y, i & 31);           // DMUL + shuffle: one step of the back-solve
    if (OP == 6) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31]; __syncwarp(); }  // STS + LDS round trip
    if (OP == 7) x = 1.0 / x + 0.75;                                    // division + DADD
    if (OP == 8) x = sqrt(x) + 0.75;                                    // sqrt + DADD
    if (OP == 9) {                                                      // one pivot step of the 8x8 tile factorisation
      const double piv = __shfl_sync(0xffffffffu, x, i & 7);
      const double ri = rsqrt(piv);
      const double tp = (lane == (i & 7)) ? piv * ri : x * ri;
      const double c = __shfl_sync(0xffffffffu, tp, (i + 1) & 7);
      x = fma(-c, tp, x + 4.0);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * 32 + lane] = x;
  if (lane == 0) cyc[OP] = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double));
  cudaMallocManaged(&cyc, 16 * sizeof(long long));
  const char* names[] = {"DFMA", "DMUL", "DADD", "rsqrt+DADD", "SHFL.64", "DMUL+SHFL.64", "STS+LDS", "1/x+DADD", "sqrt+DADD", "pivot step"};
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<1, 32>>>(out, cyc, 1.25);
    k<1><<<1, 32>>>(out, cyc, 1.25);
    k<2><<<1, 32>>>(out, cyc, 1.25);
    k<3><<<1, 32>>>(out, cyc, 1.25);
    k<4><<<1, 32>>>(out, cyc, 1.25);
    k<5><<<1, 32>>>(out, cyc, 1.25);
    k<6><<<1, 32>>>(out, cyc, 1.25);
    k<7><<<1, 32>>>(out, cyc, 1.25);
    k<8><<<1, 32>>>(out, cyc, 1.25);
    k<9><<<1, 32>>>(out, cyc, 1.25);
    cudaDeviceSynchronize();
  }
  for (int i = 0; i < 10; ++i) printf("%-14s %7.1f cycles per link\n", names[i], (double)cyc[i] / N);
  return 0;
}
