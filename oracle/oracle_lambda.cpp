// TEST INFRASTRUCTURE -- see oracle_core.h.  Restatement of the integer-ambiguity path:
//   lambda()/mlambda          RVI/gnss/src/lambda.cpp:58-235
//   matinv (LU) / solve       RVI/gnss/src/common_function.cpp:12-83,348-366, lambda.cpp:25-35
//   LambdaSearch decision     RVI/swf/swf_lambda.cpp:8-53,101-245
// The floating-point operation order follows the reference statement by statement so that the
// integer decision (and the two squared norms) can be compared bit for bit with the reference's
// own lambda.cpp compiled into oracle/_ref (tests/test_oracle_ref.py).
#include <algorithm>

#include "oracle_core.h"

namespace oracle {
namespace {

struct CM {  // column-major view, the storage convention of the RTKLIB routines
  double* p;
  int n;
  double& operator()(int r, int c) const { return p[r + (size_t)c * n]; }
};

inline double round_half_up(double x) { return std::floor(x + 0.5); }   // lambda.cpp:23
inline double sgn_rtk(double x) { return x <= 0.0 ? -1.0 : 1.0; }      // lambda.cpp:22

// Q = L' diag(D) L, processed from the last row upwards (lambda.cpp:58-76)
int factor_LtDL(int n, const double* Q, double* Lp, double* D) {
  std::vector<double> work(Q, Q + (size_t)n * n);
  CM A{work.data(), n}, L{Lp, n};
  for (int i = n - 1; i >= 0; --i) {
    D[i] = A(i, i);
    if (D[i] <= 0.0) return -1;
    const double a = std::sqrt(D[i]);
    for (int j = 0; j <= i; ++j) L(i, j) = A(i, j) / a;
    for (int j = 0; j <= i - 1; ++j)
      for (int k = 0; k <= j; ++k) A(j, k) -= L(i, k) * L(i, j);
    for (int j = 0; j <= i; ++j) L(i, j) /= L(i, i);
  }
  return 0;
}

// integer Gauss transformation of column j by row i (lambda.cpp:78-85)
void int_gauss(int n, double* Lp, double* Zp, int i, int j) {
  CM L{Lp, n}, Z{Zp, n};
  const int mu = (int)round_half_up(L(i, j));
  if (mu == 0) return;
  for (int k = i; k < n; ++k) L(k, j) -= (double)mu * L(k, i);
  for (int k = 0; k < n; ++k) Z(k, j) -= (double)mu * Z(k, i);
}

// swap of adjacent ambiguities j, j+1 (lambda.cpp:87-104)
void permute(int n, double* Lp, double* D, int j, double del, double* Zp) {
  CM L{Lp, n}, Z{Zp, n};
  const double eta = D[j] / del;
  const double lam = D[j + 1] * L(j + 1, j) / del;
  D[j] = eta * D[j + 1];
  D[j + 1] = del;
  for (int k = 0; k <= j - 1; ++k) {
    const double a0 = L(j, k), a1 = L(j + 1, k);
    L(j, k) = -L(j + 1, j) * a0 + a1;
    L(j + 1, k) = eta * a0 + lam * a1;
  }
  L(j + 1, j) = lam;
  for (int k = j + 2; k < n; ++k) std::swap(L(k, j), L(k, j + 1));
  for (int k = 0; k < n; ++k) std::swap(Z(k, j), Z(k, j + 1));
}

// decorrelation (lambda.cpp:106-121)
void reduce_lambda(int n, double* Lp, double* D, double* Zp) {
  CM L{Lp, n};
  int j = n - 2, k = n - 2;
  while (j >= 0) {
    if (j <= k)
      for (int i = j + 1; i < n; ++i) int_gauss(n, Lp, Zp, i, j);
    const double del = D[j] + L(j + 1, j) * L(j + 1, j) * D[j + 1];
    if (del + 1E-6 < D[j + 1]) {
      permute(n, Lp, D, j, del, Zp);
      k = j;
      j = n - 2;
    } else {
      --j;
    }
  }
}

// mlambda depth-first search for the m best candidates (lambda.cpp:123-191)
int search_mlambda(int n, int m, const double* Lp, const double* D, const double* zs, double* zn,
                   double* s) {
  const int kLoopMax = 10000;
  std::vector<double> Sbuf((size_t)n * n, 0.0), dist(n), zb(n), z(n), step(n);
  CM S{Sbuf.data(), n};
  CM L{const_cast<double*>(Lp), n};
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
        for (int i = 0; i <= k; ++i) S(k, i) = S(k + 1, i) + (z[k + 1] - zb[k + 1]) * L(k + 1, i);
        zb[k] = zs[k] + S(k, k);
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
      std::swap(s[i], s[j]);
      for (int q = 0; q < n; ++q) std::swap(zn[q + (size_t)i * n], zn[q + (size_t)j * n]);
    }
  return c >= kLoopMax ? -1 : 0;
}

// LU decomposition with implicit-scaling partial pivoting (common_function.cpp:12-62)
int lu_decompose(double* Ap, int n, int* indx) {
  CM A{Ap, n};
  std::vector<double> vv(n);
  int imax = 0;
  for (int i = 0; i < n; ++i) {
    double big = 0.0;
    for (int j = 0; j < n; ++j) big = std::max(big, std::fabs(A(i, j)));
    if (!(big > 0.0)) return -1;
    vv[i] = 1.0 / big;
  }
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < j; ++i) {
      double s = A(i, j);
      for (int k = 0; k < i; ++k) s -= A(i, k) * A(k, j);
      A(i, j) = s;
    }
    double big = 0.0;
    for (int i = j; i < n; ++i) {
      double s = A(i, j);
      for (int k = 0; k < j; ++k) s -= A(i, k) * A(k, j);
      A(i, j) = s;
      const double t = vv[i] * std::fabs(s);
      if (t >= big) {
        big = t;
        imax = i;
      }
    }
    if (j != imax) {
      for (int k = 0; k < n; ++k) std::swap(A(imax, k), A(j, k));
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (A(j, j) == 0.0) return -1;
    if (j != n - 1) {
      const double t = 1.0 / A(j, j);
      for (int i = j + 1; i < n; ++i) A(i, j) *= t;
    }
  }
  return 0;
}

void lu_backsub(const double* Ap, int n, const int* indx, double* b) {  // :65-83
  CM A{const_cast<double*>(Ap), n};
  int ii = -1;
  for (int i = 0; i < n; ++i) {
    const int ip = indx[i];
    double s = b[ip];
    b[ip] = b[i];
    if (ii >= 0) {
      for (int j = ii; j < i; ++j) s -= A(i, j) * b[j];
    } else if (s) {
      ii = i;
    }
    b[i] = s;
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A(i, j) * b[j];
    b[i] = s / A(i, i);
  }
}

}  // namespace

int matinv_rtk(double* A, int n) {  // returns 0 on success (the reference returns 1)
  std::vector<double> B(A, A + (size_t)n * n);
  std::vector<int> indx(n);
  if (lu_decompose(B.data(), n, indx.data())) return -1;
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) A[i + (size_t)j * n] = 0.0;
    A[j + (size_t)j * n] = 1.0;
    lu_backsub(B.data(), n, indx.data(), A + (size_t)j * n);
  }
  return 0;
}

int lambda_rtk(int n, int m, const double* a, const double* Q, double* F, double* s) {
  if (n <= 0 || m <= 0) return -1;
  std::vector<double> L((size_t)n * n, 0.0), D(n), Z((size_t)n * n, 0.0), z(n), E((size_t)n * m);
  for (int i = 0; i < n; ++i) Z[i + (size_t)i * n] = 1.0;
  int info = factor_LtDL(n, Q, L.data(), D.data());
  if (info) return info;
  reduce_lambda(n, L.data(), D.data(), Z.data());
  for (int i = 0; i < n; ++i) {  // z = Z' a   (matmul "TN", lambda.cpp:222)
    double d = 0.0;
    for (int x = 0; x < n; ++x) d += Z[x + (size_t)i * n] * a[x];
    z[i] = 1.0 * d;
  }
  info = search_mlambda(n, m, L.data(), D.data(), z.data(), E.data(), s);
  if (info) return info;
  // F = Z'^-1 E : solve("T", Z, E, n, m, F) = matinv(Z) then matmul("TN")  (lambda.cpp:25-35)
  std::vector<double> B = Z;
  if (matinv_rtk(B.data(), n)) return -1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      double d = 0.0;
      for (int x = 0; x < n; ++x) d += B[x + (size_t)i * n] * E[x + (size_t)j * n];
      F[i + (size_t)j * n] = 1.0 * d;
    }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// LambdaSearch decision.  The reference walks its rover epochs newest -> oldest, picks one
// reference satellite per (system, frequency) among the ambiguities not yet used, and adds one
// double-difference row per remaining ambiguity that passes the fractional gate.
// ---------------------------------------------------------------------------------------------
int ambiguity_fix(int n, const double* A, const double* y, int n_epochs, const int* epoch_begin,
                  const int* obs_amb, const int* obs_sysfreq, int last_fix, int* dd_pairs,
                  double* F, swgn_fix_result* res) {
  std::memset(res, 0, sizeof(*res));
  if (n < 6) {                                                   // swf_lambda.cpp:96-99
    res->status = 1;
    return 0;
  }
  Mat Am(n, n);
  for (int i = 0; i < n * n; ++i) Am.a[i] = A[i];
  Mat Qy;
  if (!inverse_lu(Am, &Qy)) {                                    // :101 A.inverse()
    res->status = 3;
    return 0;
  }
  std::vector<char> used(n, 0);
  std::vector<int> rows_a, rows_b;
  int last_count = 0, last_ref_count = 0;
  for (int ir = n_epochs - 1; ir >= 0; --ir) {                   // :126-177
    const int b0 = epoch_begin[ir], b1 = epoch_begin[ir + 1];
    int ref[6] = {-1, -1, -1, -1, -1, -1};  // index into the epoch's observation list
    // FindReferenceSatellites :8-53
    for (int sf = 0; sf < 6; ++sf) {
      std::vector<int> cand;
      for (int k = b0; k < b1; ++k)
        if (obs_sysfreq[k] == sf && obs_amb[k] >= 0 && !used[obs_amb[k]]) cand.push_back(k);
      if (cand.empty()) continue;
      std::vector<double> cost(cand.size(), 0.0);
      for (size_t j = 0; j < cand.size(); ++j) {
        const double s = y[obs_amb[cand[j]]];
        for (size_t i = 0; i < cand.size(); ++i) {
          double s2 = y[obs_amb[cand[i]]] - s;
          s2 -= std::round(s2);
          cost[j] += std::fabs(s2);
        }
      }
      const double mn = *std::min_element(cost.begin(), cost.end());
      for (size_t i = 0; i < cand.size(); ++i)
        if (cost[i] == mn) ref[sf] = cand[i];  // the LAST minimiser wins (:41-46)
    }
    for (int j = 0; j < 6; ++j)
      if (ref[j] >= 0 && ir == n_epochs - 1) last_ref_count++;
    for (int k = b0; k < b1; ++k) {                              // :136-175
      const int sf = obs_sysfreq[k];
      if (ref[sf] < 0) {
        ref[sf] = k;
        continue;
      }
      // the reference tests RTK_Npoint->use on the ambiguity object; an observation whose
      // ambiguity is outside A has no usable index and is skipped at :154
      const int a = obs_amb[k];
      if (k == ref[sf]) continue;
      if (a >= 0 && used[a]) continue;
      if (a >= 0) used[a] = 1;
      const int b = obs_amb[ref[sf]];
      if (a < 0 || b < 0) continue;
      const double d = y[a] - y[b];
      if (std::fabs(d - std::round(d)) < (last_fix ? 0.2 : 1.4)) {  // :163
        rows_a.push_back(a);
        rows_b.push_back(b);
        if (ir == n_epochs - 1) last_count++;
      }
    }
  }
  const int nb = (int)rows_a.size();
  res->n_dd = nb;
  for (int i = 0; i < nb; ++i) {
    dd_pairs[2 * i] = rows_a[i];
    dd_pairs[2 * i + 1] = rows_b[i];
  }
  if (last_count + last_ref_count < 6 || last_count < 4 || nb < 4) {  // :178,184
    res->status = 2;
    return 0;
  }
  // Qb = D Qy D', b = D y  (:189-190), D has one +1 and one -1 per row
  Mat Dm(nb, n);
  for (int i = 0; i < nb; ++i) {
    Dm(i, rows_a[i]) = 1.0;
    Dm(i, rows_b[i]) = -1.0;
  }
  Mat Qb = matmul(matmul(Dm, Qy), transpose(Dm));
  std::vector<double> bvec(nb);
  for (int i = 0; i < nb; ++i) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += Dm(i, k) * y[k];
    bvec[i] = s;
  }
  // Eigen MatrixXd is column-major: Qb.data() is column-major; Qb is symmetric up to rounding,
  // transpose explicitly to keep the exact element placement.
  std::vector<double> Qcm((size_t)nb * nb);
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < nb; ++j) Qcm[i + (size_t)j * nb] = Qb(i, j);
  double s[2] = {0, 0};
  if (lambda_rtk(nb, 2, bvec.data(), Qcm.data(), F, s)) {        // :201
    res->status = 3;
    return 0;
  }
  res->s[0] = s[0];
  res->s[1] = s[1];
  // partial ratio test :204-233
  std::vector<double> e1(nb), e2(nb);
  std::vector<int> different;
  for (int i = 0; i < nb; ++i) {
    e1[i] = F[i] - bvec[i];
    e2[i] = F[i + nb] - bvec[i];
    if (!(std::fabs(F[i] - F[i + nb]) < 1e-2)) different.push_back(i);
  }
  Mat Qb2 = Qb;
  for (int i0 : different) {
    e1[i0] = e2[i0] = 0;
    for (int j0 = 0; j0 < nb; ++j0) {
      if (i0 == j0) Qb2(i0, j0) = 1;
      else Qb2(i0, j0) = Qb2(j0, i0) = 0;
    }
  }
  Mat Qb2inv;
  double same_cost = 0.0;
  if (inverse_lu(Qb2, &Qb2inv)) {
    for (int i = 0; i < nb; ++i) {
      double t = 0.0;
      for (int j = 0; j < nb; ++j) t += Qb2inv(i, j) * e1[j];
      same_cost += e1[i] * t;
    }
  }
  double s1 = s[1] - same_cost;
  double s0 = s[0] - same_cost;
  if (std::fabs(s0) < 1e-3) s0 = 1e-3;
  res->s0_partial = s0;
  res->s1_partial = s1;
  res->n_different = (int)different.size();
  res->search_ok = (s[0] <= 0.0 || s[1] / s[0] >= 2 || s1 / s0 >= 2) ? 1 : 0;  // :237
  res->status = 0;
  return 0;
}

}  // namespace oracle
