// UpdateSchur on the device (RVI/swf/swf_gnss.cpp:25-61): after an export-mode solve the reduced system
// (S, r) = ceres::internal::{lhs_out, rhs_out} sits in W_S; the leading m = n_f - n rows are Schur-reduced
// onto the trailing n rows (the parameter_head blocks) with an EIGEN PSEUDO-INVERSE of A_mm,
//   A = A_nn - A_nm V diag(1/lambda_k if lambda_k > 1e-8 else 0) V' A_mn,   b = b_n - A_nm (...) b_m,
// exactly as the reference does for its marginalisation prior.  One CTA per call: A_mm and V in shared
// memory when they fit (m <= 100), otherwise in a global scratch (same code, L2-resident); the
// eigen-decomposition is the CTA-wide Jacobi of dev_eig.cuh (standing in for Eigen::SelfAdjointEigenSolver).
#include "dev_common.cuh"
#include "dev_eig.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {
constexpr int NT = 256;
constexpr double kEigEps = 1e-8;  // swf_gnss.cpp:45
constexpr int kMaxSharedM = 100;
}  // namespace

// scratch: W (m x (n+1)) then, when m > kMaxSharedM, A_mm and V (m x ld each)
__device__ void head_marginal_body(const DeviceBatch& b, int window, int n, double* A_out, double* b_out, double* scratch) {
  extern __shared__ __align__(16) double sm[];
  const WinDesc& d = b.desc[window];
  const double* S = b.wpool + d.woff[W_S];
  const int nf = d.n_f, ldS = d.ld, m = nf - n;
  const int tid = threadIdx.x;
  auto Sfull = [&](int i, int j) { return j >= i ? S[(size_t)i * ldS + j] : S[(size_t)j * ldS + i]; };  // selfadjointView<Upper>
  if (m == 0) {
    for (int o = tid; o < n * n; o += NT) A_out[o] = Sfull(o / n, o % n);
    for (int i = tid; i < n; i += NT) b_out[i] = S[(size_t)i * ldS + nf];
    return;
  }
  const int ld = m | 1;
  double* W = scratch;  // m x (n + 1)
  const bool in_smem = m <= kMaxSharedM;
  double* cs = sm;                       // 4 * ((m + 1) / 2) + 4
  double* red = cs + 4 * ((m + 1) / 2) + 4;
  double* A = in_smem ? red + 34 : scratch + (size_t)m * (n + 1);
  double* V = A + (size_t)m * ld;
  for (int o = tid; o < m * m; o += NT) {
    const int i = o / m, j = o - i * m;
    A[i * ld + j] = Sfull(i, j);
    V[i * ld + j] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  jacobi_eig<NT>(A, V, m, ld, cs, red);
  __syncthreads();
  // W = V' [A_mn | b_m]
  for (int o = tid; o < m * (n + 1); o += NT) {
    const int k = o / (n + 1), j = o - k * (n + 1);
    double acc = 0.0;
    for (int i = 0; i < m; ++i) acc += V[i * ld + k] * (j < n ? S[(size_t)i * ldS + m + j] : S[(size_t)i * ldS + nf]);
    W[o] = acc;
  }
  __syncthreads();
  __threadfence_block();
  for (int o = tid; o < n * (n + 1); o += NT) {
    const int i = o / (n + 1), j = o - i * (n + 1);
    double acc = 0.0;
    for (int k = 0; k < m; ++k) {
      const double lam = A[k * ld + k];
      if (lam > kEigEps) acc += W[(size_t)k * (n + 1) + i] * (1.0 / lam) * W[(size_t)k * (n + 1) + j];
    }
    if (j < n) A_out[(size_t)i * n + j] = Sfull(m + i, m + j) - acc;
    else b_out[i] = S[(size_t)(m + i) * ldS + nf] - acc;
  }
}
__global__ void __launch_bounds__(NT) k_head_marginal(DeviceBatch b, int window, int n, double* A_out, double* b_out, double* scratch) {
  head_marginal_body(b, window, n, A_out, b_out, scratch);
}
// every window of the batch: window w reduces onto its trailing n_tail[w] rows (0 = skip); off[4 * w + {0,1,2,3}] are
// the offsets of its A, b, J0 | r0 (J0 first, r0 behind it) and scratch inside buf
__global__ void __launch_bounds__(NT) k_head_marginal_batch(DeviceBatch b, const int32_t* n_tail, const int64_t* off, double* buf) {
  const int w = blockIdx.x, n = n_tail[w];
  if (n <= 0) return;
  head_marginal_body(b, w, n, buf + off[4 * w], buf + off[4 * w + 1], buf + off[4 * w + 3]);
}

// MarginalizationInfo::setmarginalizeinfo(..., Sqrt = true) (RVI/factor/marginalization_factor.cpp:449-475): the
// information form (A, b) of the head blocks becomes the next window's prior factor r = r0 + J0 (x [-] x0) with
//   J0 = sqrt(S) V',  r0 = S^-1/2 V' b,   A = V S V', eigenvalues <= 1e-8 dropped.
// One CTA; A and V in shared memory when n <= 100, otherwise in the global scratch (2 * n * (n|1) doubles).
__device__ void prior_sqrt_body(const double* A_in, const double* b_in, int n, double* J0, double* r0, double* scratch) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  const int ld = n | 1;
  const bool in_smem = n <= kMaxSharedM;
  double* cs = sm;
  double* red = cs + 4 * ((n + 1) / 2) + 4;
  double* A = in_smem ? red + 34 : scratch;
  double* V = A + (size_t)n * ld;
  for (int o = tid; o < n * n; o += NT) {
    const int i = o / n, j = o - i * n;
    A[i * ld + j] = j >= i ? A_in[(size_t)i * n + j] : A_in[(size_t)j * n + i];  // SelfAdjointEigenSolver reads one triangle
    V[i * ld + j] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();
  jacobi_eig<NT>(A, V, n, ld, cs, red);
  __syncthreads();
  for (int o = tid; o < n * n; o += NT) {
    const int i = o / n, c = o - i * n;
    const double lam = A[i * ld + i];
    J0[o] = (lam > kEigEps ? sqrt(lam) : 0.0) * V[c * ld + i];
  }
  for (int i = tid; i < n; i += NT) {
    const double lam = A[i * ld + i];
    double acc = 0.0;
    for (int c = 0; c < n; ++c) acc += V[c * ld + i] * b_in[c];
    r0[i] = (lam > kEigEps ? sqrt(1.0 / lam) : 0.0) * acc;
  }
}
__global__ void __launch_bounds__(NT) k_prior_sqrt(const double* A_in, const double* b_in, int n, double* J0, double* r0, double* scratch) {
  prior_sqrt_body(A_in, b_in, n, J0, r0, scratch);
}
__global__ void __launch_bounds__(NT) k_prior_sqrt_batch(const int32_t* n_tail, const int64_t* off, double* buf) {
  const int w = blockIdx.x, n = n_tail[w];
  if (n <= 0) return;
  double* J0 = buf + off[4 * w + 2];
  prior_sqrt_body(buf + off[4 * w], buf + off[4 * w + 1], n, J0, J0 + (size_t)n * n, buf + off[4 * w + 3]);
}

size_t prior_sqrt_scratch_doubles(int n) { return n > kMaxSharedM ? 2 * (size_t)n * (n | 1) + 2 : 2; }

cudaError_t launch_prior_sqrt(const double* A_dev, const double* b_dev, int n, double* J0_dev, double* r0_dev, double* scratch, cudaStream_t s) {
  size_t dyn = sizeof(double) * (size_t)(4 * ((n + 1) / 2) + 4 + 34);
  if (n <= kMaxSharedM) dyn += sizeof(double) * 2 * (size_t)n * (n | 1);
  if (dyn > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_prior_sqrt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  k_prior_sqrt<<<1, NT, dyn, s>>>(A_dev, b_dev, n, J0_dev, r0_dev, scratch);
  return cudaGetLastError();
}

size_t head_marginal_scratch_doubles(int m, int n) {
  size_t w = (size_t)m * (n + 1);
  if (m > kMaxSharedM) w += 2 * (size_t)m * (m | 1);
  return w + 2;
}

cudaError_t launch_head_marginal(const DeviceBatch& b, int window, int n_f, int n, double* A_dev, double* b_dev, double* scratch, cudaStream_t s) {
  const int m = n_f - n;
  size_t dyn = sizeof(double) * (size_t)(4 * ((m + 1) / 2) + 4 + 34);
  if (m <= kMaxSharedM) dyn += sizeof(double) * 2 * (size_t)m * (m | 1);
  if (dyn > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_head_marginal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
  }
  k_head_marginal<<<1, NT, dyn, s>>>(b, window, n, A_dev, b_dev, scratch);
  return cudaGetLastError();
}

// both steps for every window of the batch: two launches, one CTA per window
cudaError_t launch_marginal_priors(const DeviceBatch& b, int max_m, int max_n, const int32_t* n_tail_dev, const int64_t* off_dev, double* buf,
                                   cudaStream_t s) {
  size_t dyn1 = sizeof(double) * (size_t)(4 * ((max_m + 1) / 2) + 4 + 34);
  if (max_m <= kMaxSharedM) dyn1 += sizeof(double) * 2 * (size_t)max_m * (max_m | 1);
  size_t dyn2 = sizeof(double) * (size_t)(4 * ((max_n + 1) / 2) + 4 + 34);
  if (max_n <= kMaxSharedM) dyn2 += sizeof(double) * 2 * (size_t)max_n * (max_n | 1);
  cudaError_t e = cudaSuccess;
  if (dyn1 > 48 * 1024) e = cudaFuncSetAttribute(k_head_marginal_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn1);
  if (e == cudaSuccess && dyn2 > 48 * 1024) e = cudaFuncSetAttribute(k_prior_sqrt_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2);
  if (e != cudaSuccess) return e;
  k_head_marginal_batch<<<b.n_windows, NT, dyn1, s>>>(b, n_tail_dev, off_dev, buf);
  k_prior_sqrt_batch<<<b.n_windows, NT, dyn2, s>>>(n_tail_dev, off_dev, buf);
  return cudaGetLastError();
}

}  // namespace swgn
