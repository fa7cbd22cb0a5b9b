// K4: dense Cholesky of the reduced pose / speed-bias / GNSS-state system and the reduced solve,
// one CTA per window.  Replaces DenseSchurComplementSolver::SolveReducedLinearSystem
// (CERES/internal/ceres/schur_complement_solver.cc:223-268: Eigen LLT<Upper> + solve).
//
// S (n_f x n_f, row-major upper triangle, leading dimension ld) carries the rhs as column n_f, so
// the forward substitution U^T w = rhs is the same right-looking update as the factorisation.
// Blocked right-looking algorithm with an NB-row panel staged in shared memory:
//   panel  <- S[k0:k0+nb, k0:]           (HBM/L2 -> shared)
//   factor the panel rows in place (rank-1 steps, all threads)
//   S[k0:k0+nb, k0:] <- panel            (U rows are final)
//   trailing S[i, j] -= sum_p U[p,i] U[p,j]   (4x4 register tiles, panel operands from shared)
// then the backward solve U z = w in 32-row blocks.  The factor stays in W_S: it is the
// `lhs_out2 = llt.matrixL()` export (schur_complement_solver.cc:253-258) transposed.
#include "dev_common.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_chol(DeviceBatch b, int only_window, int NB) {
  __shared__ WinDesc sd;
  __shared__ int s_fail;
  __shared__ double s_diag[32][33];
  extern __shared__ double dyn[];  // panel NB x pw, then wv[n_f], zv[n_f]
  const int w = only_window >= 0 ? only_window : blockIdx.x;
  TRState* st = b.state + w;
  if (only_window < 0 && !(st->active && st->need_solve)) return;
  if (b.params.export_mode) return;  // the reduced system is exported, not solved
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nf = d.n_f, ld = d.ld, ncol = nf + 1;
  double* S = v.W(W_S);
  const int pw = (b.max_nf + 2) | 1;  // odd panel pitch: conflict-free column walks
  double* P = dyn;
  double* wv = dyn + (size_t)NB * pw;
  double* zv = wv + b.max_nf;
  if (tid == 0) s_fail = 0;
  __syncthreads();

  for (int k0 = 0; k0 < nf; k0 += NB) {
    const int nb = min(NB, nf - k0);
    const int width = ncol - k0;
    // ---- stage the panel
    for (int r = wid; r < nb; r += kThreads / 32) {
      const double* src = S + (size_t)(k0 + r) * ld + k0;
      double* dst = P + (size_t)r * pw;
      for (int j = r + lane; j < width; j += 32) dst[j] = src[j];
    }
    __syncthreads();
    // ---- factor the panel: row p is scaled by 1/sqrt(pivot), rows below get the rank-1 update
    for (int p = 0; p < nb; ++p) {
      const double piv = P[(size_t)p * pw + p];
      if (!(piv > 0.0)) {  // Eigen LLT: info() == NumericalIssue  -> LINEAR_SOLVER_FAILURE
        if (tid == 0) s_fail = 1;
      }
      const double x = sqrt(piv);
      __syncthreads();
      double* rowp = P + (size_t)p * pw;
      for (int j = p + tid; j < width; j += kThreads) rowp[j] = (j == p) ? x : rowp[j] / x;
      __syncthreads();
      const int nrow = nb - p - 1;
      if (nrow > 0) {
        // rows q = p+1 .. nb-1, columns j >= q
        const int wcols = width - (p + 1);
        for (int e = tid; e < nrow * wcols; e += kThreads) {
          const int q = p + 1 + e / wcols;
          const int j = p + 1 + e % wcols;
          if (j >= q) P[(size_t)q * pw + j] -= rowp[q] * rowp[j];
        }
      }
      __syncthreads();
    }
    if (s_fail) break;
    // ---- write the finished U rows back
    for (int r = wid; r < nb; r += kThreads / 32) {
      double* dst = S + (size_t)(k0 + r) * ld + k0;
      const double* src = P + (size_t)r * pw;
      for (int j = r + lane; j < width; j += 32) dst[j] = src[j];
    }
    // ---- trailing update, 4x4 tiles of the block upper triangle (columns include the rhs)
    const int t0 = k0 + nb;            // first trailing row/col
    const int tw = ncol - t0;          // trailing columns (incl. rhs)
    const int th = nf - t0;            // trailing rows
    if (th > 0) {
      const int TJ = (tw + 3) >> 2, TI = (th + 3) >> 2;
      for (int t = tid; t < TI * TJ; t += kThreads) {
        const int ti = t / TJ, tj = t - ti * TJ;
        if (tj < ti) continue;
        const int i0 = ti * 4, j0 = tj * 4;  // relative to t0
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
        const double* pa = P + nb + i0;  // panel column of trailing row i: (t0 + i0) - k0 = nb + i0
        const double* pb = P + nb + j0;
        const bool full_tile = (i0 + 4 <= th) && (j0 + 4 <= tw);
        if (full_tile) {
          for (int p = 0; p < nb; ++p) {
            const double* ra = pa + (size_t)p * pw;
            const double* rb = pb + (size_t)p * pw;
            const double a0 = ra[0], a1 = ra[1], a2 = ra[2], a3 = ra[3];
            const double b0 = rb[0], b1 = rb[1], b2 = rb[2], b3 = rb[3];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[0][2] += a0 * b2; acc[0][3] += a0 * b3;
            acc[1][0] += a1 * b0; acc[1][1] += a1 * b1; acc[1][2] += a1 * b2; acc[1][3] += a1 * b3;
            acc[2][0] += a2 * b0; acc[2][1] += a2 * b1; acc[2][2] += a2 * b2; acc[2][3] += a2 * b3;
            acc[3][0] += a3 * b0; acc[3][1] += a3 * b1; acc[3][2] += a3 * b2; acc[3][3] += a3 * b3;
          }
        } else {
          for (int p = 0; p < nb; ++p) {
            const double* ra = pa + (size_t)p * pw;
            const double* rb = pb + (size_t)p * pw;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const double av = (i0 + a < th) ? ra[a] : 0.0;
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[a][c] += av * ((j0 + c < tw) ? rb[c] : 0.0);
            }
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = i0 + a;
          if (i >= th) continue;
          double* row = S + (size_t)(t0 + i) * ld + t0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = j0 + c;
            if (j < tw && j >= i) row[j] -= acc[a][c];
          }
        }
      }
    }
    __syncthreads();
  }

  const int fail = s_fail;
  if (!fail) {
    // ---- backward solve U z = w  (w = column n_f), 32-row blocks from the bottom
    for (int i = tid; i < nf; i += kThreads) wv[i] = S[(size_t)i * ld + nf];
    __syncthreads();
    for (int k0 = ((nf - 1) / 32) * 32; k0 >= 0; k0 -= 32) {
      const int nb = min(32, nf - k0);
      for (int e = tid; e < nb * nb; e += kThreads) {
        const int r = e / nb, c = e - r * nb;
        s_diag[r][c] = (c >= r) ? S[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
      }
      __syncthreads();
      if (wid == 0) {
        double wl = lane < nb ? wv[k0 + lane] : 0.0;
        for (int i = nb - 1; i >= 0; --i) {
          double zi = 0.0;
          if (lane == i) zi = wl / s_diag[i][i];
          zi = __shfl_sync(0xffffffffu, zi, i);
          if (lane == i) wl = zi;
          if (lane < i) wl -= s_diag[lane][i] * zi;
        }
        if (lane < nb) zv[k0 + lane] = wl;
      }
      __syncthreads();
      for (int i = tid; i < k0; i += kThreads) {
        const double* row = S + (size_t)i * ld + k0;
        double s = 0.0;
        for (int j = 0; j < nb; ++j) s += row[j] * zv[k0 + j];
        wv[i] -= s;
      }
      __syncthreads();
    }
    double* Y = v.W(W_Y) + d.n_e;
    for (int i = tid; i < nf; i += kThreads) Y[i] = zv[i];
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

int chol_block(const DeviceBatch& b) { return b.max_nf <= 760 ? 32 : 16; }
size_t chol_smem(const DeviceBatch& b) {
  const int pw = (b.max_nf + 2) | 1;
  return sizeof(double) * ((size_t)chol_block(b) * pw + 2 * (size_t)b.max_nf);
}

}  // namespace

void launch_chol(const DeviceBatch& b, int only_window, cudaStream_t s) {
  const int grid = only_window >= 0 ? 1 : b.n_windows;
  k_chol<<<grid, kThreads, chol_smem(b), s>>>(b, only_window, chol_block(b));
}

void launch_tail_information(const DeviceBatch& b, int window, int n_tail, double* A_dev, cudaStream_t s) {
  k_tail_information<<<8, 256, 0, s>>>(b, window, n_tail, A_dev);
}

cudaError_t configure_chol(const DeviceBatch& b) {
  const size_t dyn = chol_smem(b);
  if (dyn > 227 * 1024) return cudaErrorInvalidValue;
  return cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
}

}  // namespace swgn
