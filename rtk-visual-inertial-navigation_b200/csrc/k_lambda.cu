// K7/K8: integer ambiguity resolution on the device.
//   lambda()/mlambda      RVI/gnss/src/lambda.cpp:58-235 (LtDL factorisation, integer Gauss
//                         decorrelation, depth-first search for the m best candidates)
//   matinv / solve        RVI/gnss/src/common_function.cpp:12-83,348-366 (LU with implicit scaling)
//   LambdaSearch decision RVI/swf/swf_lambda.cpp:8-53,101-245 (reference satellites, double
//                         differences, Qb = D Qy D', ratio tests)
// The search is inherently sequential per problem, so one thread owns one problem and thousands of
// problems run side by side.  The fix/no-fix decision must be bit-exact against
// the reference's scalar C code: this translation unit is compiled with -fmad=false so every
// multiply and add rounds separately, exactly like the reference built without FMA contraction,
// and every loop below keeps the reference's operation order.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/swgn.h"
#include "kernels.cuh"

namespace swgn {
namespace {

#define CMAT(p, n, r, c) (p)[(r) + (size_t)(c) * (n)]  // column-major like the RTKLIB routines

__device__ __forceinline__ double round_half_up(double x) { return floor(x + 0.5); }  // lambda.cpp:23
__device__ __forceinline__ double sgn_rtk(double x) { return x <= 0.0 ? -1.0 : 1.0; }  // lambda.cpp:22

// Q = L' diag(D) L, from the last row upwards (lambda.cpp:58-76); A is a scratch copy of Q
__device__ int factor_LtDL(int n, const double* Q, double* L, double* D, double* A) {
  for (int i = 0; i < n * n; ++i) A[i] = Q[i];
  for (int i = n - 1; i >= 0; --i) {
    D[i] = CMAT(A, n, i, i);
    if (D[i] <= 0.0) return -1;
    const double a = sqrt(D[i]);
    for (int j = 0; j <= i; ++j) CMAT(L, n, i, j) = CMAT(A, n, i, j) / a;
    for (int j = 0; j <= i - 1; ++j)
      for (int k = 0; k <= j; ++k) CMAT(A, n, j, k) -= CMAT(L, n, i, k) * CMAT(L, n, i, j);
    for (int j = 0; j <= i; ++j) CMAT(L, n, i, j) /= CMAT(L, n, i, i);
  }
  return 0;
}

__device__ void int_gauss(int n, double* L, double* Z, int i, int j) {  // lambda.cpp:78-85
  const int mu = (int)round_half_up(CMAT(L, n, i, j));
  if (mu == 0) return;
  for (int k = i; k < n; ++k) CMAT(L, n, k, j) -= (double)mu * CMAT(L, n, k, i);
  for (int k = 0; k < n; ++k) CMAT(Z, n, k, j) -= (double)mu * CMAT(Z, n, k, i);
}

__device__ void permute(int n, double* L, double* D, int j, double del, double* Z) {  // :87-104
  const double eta = D[j] / del;
  const double lam = D[j + 1] * CMAT(L, n, j + 1, j) / del;
  D[j] = eta * D[j + 1];
  D[j + 1] = del;
  for (int k = 0; k <= j - 1; ++k) {
    const double a0 = CMAT(L, n, j, k), a1 = CMAT(L, n, j + 1, k);
    CMAT(L, n, j, k) = -CMAT(L, n, j + 1, j) * a0 + a1;
    CMAT(L, n, j + 1, k) = eta * a0 + lam * a1;
  }
  CMAT(L, n, j + 1, j) = lam;
  for (int k = j + 2; k < n; ++k) {
    const double t = CMAT(L, n, k, j);
    CMAT(L, n, k, j) = CMAT(L, n, k, j + 1);
    CMAT(L, n, k, j + 1) = t;
  }
  for (int k = 0; k < n; ++k) {
    const double t = CMAT(Z, n, k, j);
    CMAT(Z, n, k, j) = CMAT(Z, n, k, j + 1);
    CMAT(Z, n, k, j + 1) = t;
  }
}

__device__ void reduce_lambda(int n, double* L, double* D, double* Z) {  // :106-121
  int j = n - 2, k = n - 2;
  while (j >= 0) {
    if (j <= k)
      for (int i = j + 1; i < n; ++i) int_gauss(n, L, Z, i, j);
    const double del = D[j] + CMAT(L, n, j + 1, j) * CMAT(L, n, j + 1, j) * D[j + 1];
    if (del + 1E-6 < D[j + 1]) {
      permute(n, L, D, j, del, Z);
      k = j;
      j = n - 2;
    } else {
      --j;
    }
  }
}

// mlambda search (lambda.cpp:123-191); S (n*n), dist, zb, z, step (n each) are scratch
__device__ int search_mlambda(int n, int m, const double* L, const double* D, const double* zs, double* zn, double* s,
                              double* S, double* dist, double* zb, double* z, double* step) {
  const int kLoopMax = 10000;
  for (int i = 0; i < n * n; ++i) S[i] = 0.0;
  int nn = 0, imax = 0, c;
  double maxdist = 1E99;
  int k = n - 1;
  dist[k] = 0.0;
  zb[k] = zs[k];
  z[k] = round_half_up(zb[k]);
  double y = zb[k] - z[k];
  step[k] = sgn_rtk(y);
  for (c = 0; c < kLoopMax; ++c) {
    const double newdist = dist[k] + y * y / D[k];
    if (newdist < maxdist) {
      if (k != 0) {
        dist[--k] = newdist;
        for (int i = 0; i <= k; ++i)
          CMAT(S, n, k, i) = CMAT(S, n, k + 1, i) + (z[k + 1] - zb[k + 1]) * CMAT(L, n, k + 1, i);
        zb[k] = zs[k] + CMAT(S, n, k, k);
        z[k] = round_half_up(zb[k]);
        y = zb[k] - z[k];
        step[k] = sgn_rtk(y);
      } else {
        if (nn < m) {
          if (nn == 0 || newdist > s[imax]) imax = nn;
          for (int i = 0; i < n; ++i) zn[i + (size_t)nn * n] = z[i];
          s[nn++] = newdist;
        } else {
          if (newdist < s[imax]) {
            for (int i = 0; i < n; ++i) zn[i + (size_t)imax * n] = z[i];
            s[imax] = newdist;
            imax = 0;
            for (int i = 0; i < m; ++i)
              if (s[imax] < s[i]) imax = i;
          }
          maxdist = s[imax];
        }
        z[0] += step[0];
        y = zb[0] - z[0];
        step[0] = -step[0] - sgn_rtk(step[0]);
      }
    } else {
      if (k == n - 1) break;
      ++k;
      z[k] += step[k];
      y = zb[k] - z[k];
      step[k] = -step[k] - sgn_rtk(step[k]);
    }
  }
  for (int i = 0; i < m - 1; ++i)
    for (int j = i + 1; j < m; ++j) {
      if (s[i] < s[j]) continue;
      const double t = s[i];
      s[i] = s[j];
      s[j] = t;
      for (int q = 0; q < n; ++q) {
        const double u = zn[q + (size_t)i * n];
        zn[q + (size_t)i * n] = zn[q + (size_t)j * n];
        zn[q + (size_t)j * n] = u;
      }
    }
  return c >= kLoopMax ? -1 : 0;
}

// LU decomposition with implicit-scaling partial pivoting (common_function.cpp:12-62);
// indx is kept in doubles' storage as int32
__device__ int lu_decompose(double* A, int n, int32_t* indx, double* vv) {
  int imax = 0;
  for (int i = 0; i < n; ++i) {
    double big = 0.0;
    for (int j = 0; j < n; ++j) {
      const double t = fabs(CMAT(A, n, i, j));
      if (t > big) big = t;
    }
    if (!(big > 0.0)) return -1;
    vv[i] = 1.0 / big;
  }
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < j; ++i) {
      double s = CMAT(A, n, i, j);
      for (int k = 0; k < i; ++k) s -= CMAT(A, n, i, k) * CMAT(A, n, k, j);
      CMAT(A, n, i, j) = s;
    }
    double big = 0.0;
    for (int i = j; i < n; ++i) {
      double s = CMAT(A, n, i, j);
      for (int k = 0; k < j; ++k) s -= CMAT(A, n, i, k) * CMAT(A, n, k, j);
      CMAT(A, n, i, j) = s;
      const double t = vv[i] * fabs(s);
      if (t >= big) {
        big = t;
        imax = i;
      }
    }
    if (j != imax) {
      for (int k = 0; k < n; ++k) {
        const double t = CMAT(A, n, imax, k);
        CMAT(A, n, imax, k) = CMAT(A, n, j, k);
        CMAT(A, n, j, k) = t;
      }
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (CMAT(A, n, j, j) == 0.0) return -1;
    if (j != n - 1) {
      const double t = 1.0 / CMAT(A, n, j, j);
      for (int i = j + 1; i < n; ++i) CMAT(A, n, i, j) *= t;
    }
  }
  return 0;
}

__device__ void lu_backsub(const double* A, int n, const int32_t* indx, double* b) {  // :65-83
  int ii = -1;
  for (int i = 0; i < n; ++i) {
    const int ip = indx[i];
    double s = b[ip];
    b[ip] = b[i];
    if (ii >= 0) {
      for (int j = ii; j < i; ++j) s -= CMAT(A, n, i, j) * b[j];
    } else if (s) {
      ii = i;
    }
    b[i] = s;
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= CMAT(A, n, i, j) * b[j];
    b[i] = s / CMAT(A, n, i, i);
  }
}

// matinv (common_function.cpp:348-366): A <- A^-1, B (n*n), vv (n), indx (n) scratch
__device__ int matinv_rtk(double* A, int n, double* B, double* vv, int32_t* indx) {
  for (int i = 0; i < n * n; ++i) B[i] = A[i];
  if (lu_decompose(B, n, indx, vv)) return -1;
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) A[i + (size_t)j * n] = 0.0;
    A[j + (size_t)j * n] = 1.0;
    lu_backsub(B, n, indx, A + (size_t)j * n);
  }
  return 0;
}

// work layout (doubles): L n^2 | Z n^2 | T n^2 (scratch: LtDL copy, search S, matinv B) | D n | z n |
//                        E n*m | dist n | zb n | zz n | step n | vv n | indx n (as int32)
__device__ int lambda_device(int n, int m, const double* a, const double* Q, double* F, double* s, double* work) {
  if (n <= 0 || m <= 0) return -1;
  double* L = work;
  double* Z = L + (size_t)n * n;
  double* T = Z + (size_t)n * n;
  double* D = T + (size_t)n * n;
  double* z = D + n;
  double* E = z + n;
  double* dist = E + (size_t)n * m;
  double* zb = dist + n;
  double* zz = zb + n;
  double* step = zz + n;
  double* vv = step + n;
  int32_t* indx = reinterpret_cast<int32_t*>(vv + n);
  for (int i = 0; i < n * n; ++i) {
    L[i] = 0.0;
    Z[i] = 0.0;
  }
  for (int i = 0; i < n; ++i) Z[i + (size_t)i * n] = 1.0;
  int info = factor_LtDL(n, Q, L, D, T);
  if (info) return info;
  reduce_lambda(n, L, D, Z);
  for (int i = 0; i < n; ++i) {  // z = Z' a
    double d = 0.0;
    for (int x = 0; x < n; ++x) d += Z[x + (size_t)i * n] * a[x];
    z[i] = 1.0 * d;
  }
  info = search_mlambda(n, m, L, D, z, E, s, T, dist, zb, zz, step);
  if (info) return info;
  // F = Z'^-1 E  (solve "T": matinv(Z) then matmul "TN", lambda.cpp:25-35); L is free now
  for (int i = 0; i < n * n; ++i) L[i] = Z[i];
  if (matinv_rtk(L, n, T, vv, indx)) return -1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double d = 0.0;
      for (int x = 0; x < n; ++x) d += L[x + (size_t)i * n] * E[x + (size_t)j * n];
      F[i + (size_t)j * n] = 1.0 * d;
    }
  return 0;
}

__global__ void k_lambda_batch(int n_problems, int m, const int32_t* n, const int64_t* aoff, const int64_t* qoff,
                               const double* a, const double* Q, double* F, double* s, int32_t* info, double* work,
                               const int64_t* woff) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_problems) return;
  info[p] = lambda_device(n[p], m, a + aoff[p], Q + qoff[p], F + aoff[p] * m, s + (size_t)p * m, work + woff[p]);
}

// ---- LambdaSearch decision for one window -----------------------------------------------------
// general inverse by LU with partial pivoting on row-major storage: the restatement of Eigen's
// MatrixXd::inverse() used at swf_lambda.cpp:101,228
__device__ bool inverse_lu(const double* A, int n, double* inv, double* lu, int32_t* piv, double* x) {
  for (int i = 0; i < n * n; ++i) lu[i] = A[i];
  for (int i = 0; i < n; ++i) piv[i] = i;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(lu[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(lu[(size_t)i * n + k]) > best) {
        best = fabs(lu[(size_t)i * n + k]);
        p = i;
      }
    if (best == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < n; ++j) {
        const double t = lu[(size_t)k * n + j];
        lu[(size_t)k * n + j] = lu[(size_t)p * n + j];
        lu[(size_t)p * n + j] = t;
      }
      const int t = piv[k];
      piv[k] = piv[p];
      piv[p] = t;
    }
    for (int i = k + 1; i < n; ++i) {
      lu[(size_t)i * n + k] /= lu[(size_t)k * n + k];
      const double f = lu[(size_t)i * n + k];
      for (int j = k + 1; j < n; ++j) lu[(size_t)i * n + j] -= f * lu[(size_t)k * n + j];
    }
  }
  for (int c = 0; c < n; ++c) {
    for (int i = 0; i < n; ++i) x[i] = (piv[i] == c) ? 1.0 : 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) x[i] -= lu[(size_t)i * n + j] * x[j];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) x[i] -= lu[(size_t)i * n + j] * x[j];
      x[i] /= lu[(size_t)i * n + i];
    }
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + c] = x[i];
  }
  return true;
}

struct FixArgs {
  int n, n_epochs, last_fix;
  const double* A;
  const double* y;
  const int32_t* epoch_begin;
  const int32_t* obs_amb;
  const int32_t* obs_sysfreq;
  int32_t* dd_pairs;
  double* F;
  swgn_fix_result* res;
  double* work;    // 8 n^2 + 8 n + lambda work
  int32_t* iwork;  // 6 n + n_obs
};

__device__ void ambiguity_fix_one(const FixArgs& g) {
  const int n = g.n;
  swgn_fix_result res;
  res.n_dd = 0; res.status = 0; res.search_ok = 0; res.n_different = 0;
  res.s[0] = res.s[1] = 0.0; res.s0_partial = res.s1_partial = 0.0;
  if (n < 6) {  // swf_lambda.cpp:96-99
    res.status = 1;
    *g.res = res;
    return;
  }
  double* Qy = g.work;
  double* lu = Qy + (size_t)n * n;
  double* T = lu + (size_t)n * n;       // D Qy   (nb x n)
  double* Qb = T + (size_t)n * n;       // nb x nb row-major
  double* Qcm = Qb + (size_t)n * n;     // column-major copy
  double* Qb2 = Qcm + (size_t)n * n;
  double* Qb2inv = Qb2 + (size_t)n * n;
  double* xv = Qb2inv + (size_t)n * n;
  double* bvec = xv + n;
  double* e1 = bvec + n;
  double* e2 = e1 + n;
  double* cost = e2 + n;
  double* lwork = cost + n;
  int32_t* piv = g.iwork;
  int32_t* used = piv + n;
  int32_t* rows_a = used + n;
  int32_t* rows_b = rows_a + n;
  int32_t* cand = rows_b + n;
  int32_t* different = cand + n;
  if (!inverse_lu(g.A, n, Qy, lu, piv, xv)) {  // :101
    res.status = 3;
    *g.res = res;
    return;
  }
  for (int i = 0; i < n; ++i) used[i] = 0;
  int nb = 0, last_count = 0, last_ref_count = 0;
  const double* y = g.y;
  for (int ir = g.n_epochs - 1; ir >= 0; --ir) {  // :126-177
    const int b0 = g.epoch_begin[ir], b1 = g.epoch_begin[ir + 1];
    int ref[6] = {-1, -1, -1, -1, -1, -1};
    for (int sf = 0; sf < 6; ++sf) {  // FindReferenceSatellites :8-53
      int nc = 0;
      for (int k = b0; k < b1; ++k)
        if (g.obs_sysfreq[k] == sf && g.obs_amb[k] >= 0 && !used[g.obs_amb[k]] && nc < n) cand[nc++] = k;
      if (nc == 0) continue;
      double mn = 0.0;
      for (int j = 0; j < nc; ++j) {
        const double s = y[g.obs_amb[cand[j]]];
        double cj = 0.0;
        for (int i = 0; i < nc; ++i) {
          double s2 = y[g.obs_amb[cand[i]]] - s;
          s2 -= round(s2);
          cj += fabs(s2);
        }
        cost[j] = cj;
        if (j == 0 || cj < mn) mn = cj;
      }
      for (int i = 0; i < nc; ++i)
        if (cost[i] == mn) ref[sf] = cand[i];  // the last minimiser wins (:41-46)
    }
    for (int j = 0; j < 6; ++j)
      if (ref[j] >= 0 && ir == g.n_epochs - 1) last_ref_count++;
    for (int k = b0; k < b1; ++k) {  // :136-175
      const int sf = g.obs_sysfreq[k];
      if (ref[sf] < 0) {
        ref[sf] = k;
        continue;
      }
      const int a = g.obs_amb[k];
      if (k == ref[sf]) continue;
      if (a >= 0 && used[a]) continue;
      if (a >= 0) used[a] = 1;
      const int bq = g.obs_amb[ref[sf]];
      if (a < 0 || bq < 0) continue;
      const double dlt = y[a] - y[bq];
      if (fabs(dlt - round(dlt)) < (g.last_fix ? 0.2 : 1.4)) {  // :163
        if (nb < n) {
          rows_a[nb] = a;
          rows_b[nb] = bq;
          ++nb;
        }
        if (ir == g.n_epochs - 1) last_count++;
      }
    }
  }
  res.n_dd = nb;
  for (int i = 0; i < nb; ++i) {
    g.dd_pairs[2 * i] = rows_a[i];
    g.dd_pairs[2 * i + 1] = rows_b[i];
  }
  if (last_count + last_ref_count < 6 || last_count < 4 || nb < 4) {  // :178,184
    res.status = 2;
    *g.res = res;
    return;
  }
  // Qb = (D Qy) D', b = D y with D = one +1 and one -1 per row; the dense products of the
  // restatement skip exact zeros and add the remaining terms in ascending column order
  for (int i = 0; i < nb; ++i) {
    const int ca = rows_a[i], cb = rows_b[i];
    const int k0 = ca < cb ? ca : cb, k1 = ca < cb ? cb : ca;
    const double v0 = ca < cb ? 1.0 : -1.0, v1 = -v0;
    for (int j = 0; j < n; ++j) {
      double c = 0.0;
      c += v0 * Qy[(size_t)k0 * n + j];
      c += v1 * Qy[(size_t)k1 * n + j];
      T[(size_t)i * n + j] = c;
    }
    double s = 0.0;
    for (int k = 0; k < n; ++k) {
      const double dk = (k == ca) ? 1.0 : ((k == cb) ? -1.0 : 0.0);
      s += dk * y[k];
    }
    bvec[i] = s;
  }
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < nb; ++j) Qb[(size_t)i * nb + j] = 0.0;
  for (int i = 0; i < nb; ++i)
    for (int k = 0; k < n; ++k) {
      const double a = T[(size_t)i * n + k];
      if (a == 0.0) continue;
      for (int j = 0; j < nb; ++j) {
        const double dt = (k == rows_a[j]) ? 1.0 : ((k == rows_b[j]) ? -1.0 : 0.0);
        Qb[(size_t)i * nb + j] += a * dt;
      }
    }
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < nb; ++j) Qcm[i + (size_t)j * nb] = Qb[(size_t)i * nb + j];
  double s[2] = {0.0, 0.0};
  if (lambda_device(nb, 2, bvec, Qcm, g.F, s, lwork)) {  // :201
    res.status = 3;
    *g.res = res;
    return;
  }
  res.s[0] = s[0];
  res.s[1] = s[1];
  // partial ratio test :204-233
  int ndiff = 0;
  for (int i = 0; i < nb; ++i) {
    e1[i] = g.F[i] - bvec[i];
    e2[i] = g.F[i + nb] - bvec[i];
    if (!(fabs(g.F[i] - g.F[i + nb]) < 1e-2)) different[ndiff++] = i;
  }
  for (int i = 0; i < nb * nb; ++i) Qb2[i] = Qb[i];
  for (int d = 0; d < ndiff; ++d) {
    const int i0 = different[d];
    e1[i0] = e2[i0] = 0;
    for (int j0 = 0; j0 < nb; ++j0) {
      if (i0 == j0) Qb2[(size_t)i0 * nb + j0] = 1;
      else Qb2[(size_t)i0 * nb + j0] = Qb2[(size_t)j0 * nb + i0] = 0;
    }
  }
  double same_cost = 0.0;
  if (inverse_lu(Qb2, nb, Qb2inv, lu, piv, xv)) {
    for (int i = 0; i < nb; ++i) {
      double t = 0.0;
      for (int j = 0; j < nb; ++j) t += Qb2inv[(size_t)i * nb + j] * e1[j];
      same_cost += e1[i] * t;
    }
  }
  const double s1 = s[1] - same_cost;
  double s0 = s[0] - same_cost;
  if (fabs(s0) < 1e-3) s0 = 1e-3;
  res.s0_partial = s0;
  res.s1_partial = s1;
  res.n_different = ndiff;
  res.search_ok = (s[0] <= 0.0 || s[1] / s[0] >= 2 || s1 / s0 >= 2) ? 1 : 0;  // :237
  res.status = 0;
  *g.res = res;
}

__global__ void k_ambiguity_fix(FixArgs g) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  ambiguity_fix_one(g);
}

}  // namespace

size_t fix_work_doubles(int n) { return (size_t)8 * n * n + (size_t)8 * n + lambda_work_doubles(n, 2); }
size_t fix_work_ints(int n) { return (size_t)6 * n + 8; }

// One thread per window: the same sequential decision code, thousands of windows side by side.
__global__ void k_ambiguity_fix_batch(int n_windows, int n, const double* A_all, const double* y_all, const int32_t* win_epoch,
                                      const int32_t* epoch_begin, const int32_t* obs_amb, const int32_t* obs_sysfreq,
                                      const int32_t* last_fix, const int32_t* have_A, int32_t* dd_pairs, double* F, swgn_fix_result* res,
                                      double* work, int32_t* iwork, size_t nwork, size_t niwork) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_windows) return;
  if (!have_A[w]) {  // no Cholesky factor (the last reduced solve failed): nothing to search
    swgn_fix_result r;
    r.n_dd = 0; r.status = 3; r.search_ok = 0; r.n_different = 0;
    r.s[0] = r.s[1] = 0.0; r.s0_partial = r.s1_partial = 0.0;
    res[w] = r;
    return;
  }
  FixArgs g;
  g.n = n;
  g.n_epochs = win_epoch[w + 1] - win_epoch[w];
  g.last_fix = last_fix ? last_fix[w] : 0;
  g.A = A_all + (size_t)w * n * n;
  g.y = y_all + (size_t)w * n;
  g.epoch_begin = epoch_begin + win_epoch[w];
  g.obs_amb = obs_amb;
  g.obs_sysfreq = obs_sysfreq;
  g.dd_pairs = dd_pairs + (size_t)2 * n * w;
  g.F = F + (size_t)2 * n * w;
  g.res = res + w;
  g.work = work + nwork * w;
  g.iwork = iwork + niwork * w;
  ambiguity_fix_one(g);
}

void launch_ambiguity_fix_batch(int n_windows, int n, const double* A_all, const double* y_all, const int32_t* win_epoch,
                                const int32_t* epoch_begin, const int32_t* obs_amb, const int32_t* obs_sysfreq, const int32_t* last_fix,
                                const int32_t* have_A, int32_t* dd_pairs, double* F, swgn_fix_result* res, double* work, int32_t* iwork,
                                cudaStream_t s) {
  const int threads = 32;
  k_ambiguity_fix_batch<<<(n_windows + threads - 1) / threads, threads, 0, s>>>(n_windows, n, A_all, y_all, win_epoch, epoch_begin, obs_amb,
                                                                               obs_sysfreq, last_fix, have_A, dd_pairs, F, res, work, iwork,
                                                                               fix_work_doubles(n), fix_work_ints(n));
}

size_t lambda_work_doubles(int n, int m) { return (size_t)3 * n * n + (size_t)n * m + (size_t)8 * n + 8; }

void launch_lambda_batch(int n_problems, int m, const int32_t* n_dev, const int64_t* aoff_dev, const int64_t* qoff_dev,
                         const double* a_dev, const double* Q_dev, double* F_dev, double* s_dev, int32_t* info_dev,
                         double* work_dev, const int64_t* woff_dev, cudaStream_t s) {
  const int threads = 32;
  k_lambda_batch<<<(n_problems + threads - 1) / threads, threads, 0, s>>>(n_problems, m, n_dev, aoff_dev, qoff_dev, a_dev, Q_dev,
                                                                         F_dev, s_dev, info_dev, work_dev, woff_dev);
}

}  // namespace swgn

// ---- C ABI ------------------------------------------------------------------------------------
namespace {
thread_local char g_lerr[256];
}
extern "C" const char* swgn_lambda_last_error(void) { return g_lerr; }

#define LCU(call)                                                                 \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) {                                                      \
      snprintf(g_lerr, sizeof(g_lerr), "%s: %s", #call, cudaGetErrorString(e_));  \
      st = SWGN_ERR_CUDA;                                                         \
      goto done;                                                                  \
    }                                                                             \
  } while (0)

#include <cstdio>
#include <vector>

extern "C" swgn_status swgn_lambda_batch(int32_t device, int32_t n_problems, const int32_t* n, int32_t m, const double* a,
                                         const double* Q, double* F, double* s, int32_t* info) {
  if (n_problems <= 0 || !n || m <= 0 || !a || !Q || !F || !s || !info) return SWGN_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SWGN_ERR_NO_DEVICE;
  swgn_status st = SWGN_OK;
  std::vector<int64_t> aoff(n_problems), qoff(n_problems), woff(n_problems);
  int64_t ao = 0, qo = 0, wo = 0;
  for (int p = 0; p < n_problems; ++p) {
    if (n[p] <= 0) return SWGN_ERR_INVALID;
    aoff[p] = ao;
    qoff[p] = qo;
    woff[p] = wo;
    ao += n[p];
    qo += (int64_t)n[p] * n[p];
    wo += (int64_t)swgn::lambda_work_doubles(n[p], m);
  }
  int32_t *d_n = nullptr, *d_info = nullptr;
  int64_t *d_aoff = nullptr, *d_qoff = nullptr, *d_woff = nullptr;
  double *d_a = nullptr, *d_Q = nullptr, *d_F = nullptr, *d_s = nullptr, *d_work = nullptr;
  LCU(cudaSetDevice(device));
  LCU(cudaMalloc(&d_n, sizeof(int32_t) * n_problems));
  LCU(cudaMalloc(&d_info, sizeof(int32_t) * n_problems));
  LCU(cudaMalloc(&d_aoff, sizeof(int64_t) * n_problems));
  LCU(cudaMalloc(&d_qoff, sizeof(int64_t) * n_problems));
  LCU(cudaMalloc(&d_woff, sizeof(int64_t) * n_problems));
  LCU(cudaMalloc(&d_a, sizeof(double) * ao));
  LCU(cudaMalloc(&d_Q, sizeof(double) * qo));
  LCU(cudaMalloc(&d_F, sizeof(double) * ao * m));
  LCU(cudaMalloc(&d_s, sizeof(double) * n_problems * m));
  LCU(cudaMalloc(&d_work, sizeof(double) * wo));
  LCU(cudaMemcpy(d_n, n, sizeof(int32_t) * n_problems, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_aoff, aoff.data(), sizeof(int64_t) * n_problems, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_qoff, qoff.data(), sizeof(int64_t) * n_problems, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_woff, woff.data(), sizeof(int64_t) * n_problems, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_a, a, sizeof(double) * ao, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_Q, Q, sizeof(double) * qo, cudaMemcpyHostToDevice));
  LCU(cudaMemset(d_F, 0, sizeof(double) * ao * m));
  LCU(cudaMemset(d_s, 0, sizeof(double) * n_problems * m));
  swgn::launch_lambda_batch(n_problems, m, d_n, d_aoff, d_qoff, d_a, d_Q, d_F, d_s, d_info, d_work, d_woff, 0);
  LCU(cudaGetLastError());
  LCU(cudaMemcpy(F, d_F, sizeof(double) * ao * m, cudaMemcpyDeviceToHost));
  LCU(cudaMemcpy(s, d_s, sizeof(double) * n_problems * m, cudaMemcpyDeviceToHost));
  LCU(cudaMemcpy(info, d_info, sizeof(int32_t) * n_problems, cudaMemcpyDeviceToHost));
done:
  cudaFree(d_n); cudaFree(d_info); cudaFree(d_aoff); cudaFree(d_qoff); cudaFree(d_woff);
  cudaFree(d_a); cudaFree(d_Q); cudaFree(d_F); cudaFree(d_s); cudaFree(d_work);
  return st;
}

extern "C" swgn_status swgn_ambiguity_fix(int32_t device, int32_t n, const double* A, const double* y, int32_t n_epochs,
                                          const int32_t* epoch_begin, const int32_t* obs_amb, const int32_t* obs_sysfreq,
                                          int32_t last_fix, int32_t* dd_pairs, double* F, swgn_fix_result* result) {
  if (n <= 0 || !A || !y || n_epochs <= 0 || !epoch_begin || !obs_amb || !obs_sysfreq || !dd_pairs || !F || !result)
    return SWGN_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SWGN_ERR_NO_DEVICE;
  swgn_status st = SWGN_OK;
  const int n_obs = epoch_begin[n_epochs];
  const size_t nwork = swgn::fix_work_doubles(n);
  const size_t niwork = swgn::fix_work_ints(n);
  double *d_A = nullptr, *d_y = nullptr, *d_F = nullptr, *d_work = nullptr;
  int32_t *d_eb = nullptr, *d_oa = nullptr, *d_sf = nullptr, *d_pairs = nullptr, *d_iwork = nullptr;
  swgn_fix_result* d_res = nullptr;
  swgn::FixArgs g;
  LCU(cudaSetDevice(device));
  LCU(cudaMalloc(&d_A, sizeof(double) * n * n));
  LCU(cudaMalloc(&d_y, sizeof(double) * n));
  LCU(cudaMalloc(&d_F, sizeof(double) * 2 * n));
  LCU(cudaMalloc(&d_work, sizeof(double) * nwork));
  LCU(cudaMalloc(&d_eb, sizeof(int32_t) * (n_epochs + 1)));
  LCU(cudaMalloc(&d_oa, sizeof(int32_t) * (n_obs + 1)));
  LCU(cudaMalloc(&d_sf, sizeof(int32_t) * (n_obs + 1)));
  LCU(cudaMalloc(&d_pairs, sizeof(int32_t) * 2 * n));
  LCU(cudaMalloc(&d_iwork, sizeof(int32_t) * niwork));
  LCU(cudaMalloc(&d_res, sizeof(swgn_fix_result)));
  LCU(cudaMemcpy(d_A, A, sizeof(double) * n * n, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_y, y, sizeof(double) * n, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_eb, epoch_begin, sizeof(int32_t) * (n_epochs + 1), cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_oa, obs_amb, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice));
  LCU(cudaMemcpy(d_sf, obs_sysfreq, sizeof(int32_t) * n_obs, cudaMemcpyHostToDevice));
  LCU(cudaMemset(d_F, 0, sizeof(double) * 2 * n));
  LCU(cudaMemset(d_pairs, 0, sizeof(int32_t) * 2 * n));
  g.n = n; g.n_epochs = n_epochs; g.last_fix = last_fix;
  g.A = d_A; g.y = d_y; g.epoch_begin = d_eb; g.obs_amb = d_oa; g.obs_sysfreq = d_sf;
  g.dd_pairs = d_pairs; g.F = d_F; g.res = d_res; g.work = d_work; g.iwork = d_iwork;
  swgn::k_ambiguity_fix<<<1, 32>>>(g);
  LCU(cudaGetLastError());
  LCU(cudaMemcpy(result, d_res, sizeof(swgn_fix_result), cudaMemcpyDeviceToHost));
  LCU(cudaMemcpy(dd_pairs, d_pairs, sizeof(int32_t) * 2 * (result->n_dd > n ? n : result->n_dd), cudaMemcpyDeviceToHost));
  LCU(cudaMemcpy(F, d_F, sizeof(double) * 2 * (result->n_dd > n ? n : result->n_dd), cudaMemcpyDeviceToHost));
done:
  cudaFree(d_A); cudaFree(d_y); cudaFree(d_F); cudaFree(d_work); cudaFree(d_eb); cudaFree(d_oa); cudaFree(d_sf);
  cudaFree(d_pairs); cudaFree(d_iwork); cudaFree(d_res);
  return st;
}
