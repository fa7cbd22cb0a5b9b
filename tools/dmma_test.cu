// development aid: prints the fragment layout of mma.sync.m8n8k4.f64 as executed on this GPU
#include <cstdio>
__global__ void k(double* out) {
  int lane = threadIdx.x;
  // A[i][k] = 10*i + k ; B[k][n] = (k==0) ? n : 0  -> D[i][n] = (10 i) * n ... use identity-ish probes
  double a = 10.0 * (lane >> 2) + (lane & 3);          // assume A[row=lane>>2][col=lane&3]
  double b = ((lane & 3) == 1) ? (double)(lane >> 2) + 1 : 0.0;  // assume B[row=lane&3][col=lane>>2]; only k=1 row nonzero: B[1][n] = n+1
  double d0 = 0, d1 = 0;
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
  out[lane * 2] = d0;
  out[lane * 2 + 1] = d1;
}
int main() {
  double* d; cudaMalloc(&d, 64 * 8);
  k<<<1, 32>>>(d);
  double h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  // expected D[i][n] = A[i][1] * B[1][n] = (10 i + 1)(n + 1) at lane (i*4 + n/2), slot n%2
  int bad = 0;
  for (int i = 0; i < 8; ++i) for (int n = 0; n < 8; ++n) {
    double e = (10.0 * i + 1) * (n + 1), g = h[(i * 4 + n / 2) * 2 + (n & 1)];
    if (e != g) { if (bad < 8) printf("mismatch i=%d n=%d expect %g got %g\n", i, n, e, g); ++bad; }
  }
  printf("dmma layout %s (%d mismatches) err=%s\n", bad ? "WRONG" : "ok", bad, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
