// TEST INFRASTRUCTURE -- see oracle_core.h.  Dense helpers standing in for the Eigen calls of the
// reference (LLT<Upper>, inverse(), SelfAdjointEigenSolver, Quaternion::toRotationMatrix).
#include <algorithm>

#include "oracle_core.h"

namespace oracle {

Mat matmul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int k = 0; k < A.c; ++k) {
      double a = A(i, k);
      if (a == 0.0) continue;
      for (int j = 0; j < B.c; ++j) C(i, j) += a * B(k, j);
    }
  return C;
}

Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j);
  return T;
}

// Eigen's unblocked LLT kernel (Eigen/src/Cholesky/LLT.h, llt_inplace::unblocked) visits the
// matrix column by column of the lower factor; on the row-major upper triangle that is row k:
// x = a_kk - |u_{0..k-1,k}|^2 ; fail if x <= 0 ; u_kj = (a_kj - sum_p u_pk u_pj) / sqrt(x).
bool llt_upper_inplace(double* a, int n) {
  for (int k = 0; k < n; ++k) {
    double x = a[(size_t)k * n + k];
    for (int p = 0; p < k; ++p) x -= a[(size_t)p * n + k] * a[(size_t)p * n + k];
    if (!(x > 0.0)) return false;
    x = std::sqrt(x);
    a[(size_t)k * n + k] = x;
    for (int j = k + 1; j < n; ++j) {
      double s = a[(size_t)k * n + j];
      for (int p = 0; p < k; ++p) s -= a[(size_t)p * n + k] * a[(size_t)p * n + j];
      a[(size_t)k * n + j] = s / x;
    }
  }
  return true;
}

void llt_upper_solve(const double* u, int n, double* x) {
  // U^T w = b (forward), then U x = w (backward)
  for (int i = 0; i < n; ++i) {
    double s = x[i];
    for (int p = 0; p < i; ++p) s -= u[(size_t)p * n + i] * x[p];
    x[i] = s / u[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int p = i + 1; p < n; ++p) s -= u[(size_t)i * n + p] * x[p];
    x[i] = s / u[(size_t)i * n + i];
  }
}

bool invert_psd(const double* m, int n, double* inv) {
  std::vector<double> u(m, m + (size_t)n * n);
  if (!llt_upper_inplace(u.data(), n)) {
    // Eigen's llt().solve() does not check info(); mirror "garbage in" with NaNs so that the
    // caller's IsArrayValid test (dogleg_strategy.cc:589) rejects the solve.
    for (int i = 0; i < n * n; ++i) inv[i] = std::numeric_limits<double>::quiet_NaN();
    return false;
  }
  std::vector<double> col(n);
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) col[i] = (i == j) ? 1.0 : 0.0;
    llt_upper_solve(u.data(), n, col.data());
    for (int i = 0; i < n; ++i) inv[(size_t)i * n + j] = col[i];
  }
  return true;
}

bool inverse_lu(const Mat& A, Mat* inv) {
  int n = A.r;
  Mat lu = A;
  std::vector<int> piv(n);
  for (int i = 0; i < n; ++i) piv[i] = i;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = std::fabs(lu(k, k));
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(lu(i, k)) > best) {
        best = std::fabs(lu(i, k));
        p = i;
      }
    if (best == 0.0) return false;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(lu(k, j), lu(p, j));
      std::swap(piv[k], piv[p]);
    }
    for (int i = k + 1; i < n; ++i) {
      lu(i, k) /= lu(k, k);
      double f = lu(i, k);
      for (int j = k + 1; j < n; ++j) lu(i, j) -= f * lu(k, j);
    }
  }
  *inv = Mat(n, n);
  std::vector<double> x(n);
  for (int c = 0; c < n; ++c) {
    for (int i = 0; i < n; ++i) x[i] = (piv[i] == c) ? 1.0 : 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) x[i] -= lu(i, j) * x[j];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) x[i] -= lu(i, j) * x[j];
      x[i] /= lu(i, i);
    }
    for (int i = 0; i < n; ++i) (*inv)(i, c) = x[i];
  }
  return true;
}

void eig_sym(const Mat& A, std::vector<double>* w, Mat* V) {
  int n = A.r;
  Mat a = A;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) a(i, j) = a(j, i);  // selfadjoint view of the upper part
  Mat v = Mat::Identity(n);
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += a(i, i) * a(i, i);
      for (int j = i + 1; j < n; ++j) off += a(i, j) * a(i, j);
    }
    if (off <= 1e-32 * (diag + 1e-300)) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = a(p, q);
        if (apq == 0.0) continue;
        double theta = (a(q, q) - a(p, p)) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          double akp = a(k, p), akq = a(k, q);
          a(k, p) = c * akp - s * akq;
          a(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          double apk = a(p, k), aqk = a(q, k);
          a(p, k) = c * apk - s * aqk;
          a(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          double vkp = v(k, p), vkq = v(k, q);
          v(k, p) = c * vkp - s * vkq;
          v(k, q) = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int x, int y) { return a(x, x) < a(y, y); });
  w->resize(n);
  *V = Mat(n, n);
  for (int j = 0; j < n; ++j) {
    (*w)[j] = a(idx[j], idx[j]);
    for (int i = 0; i < n; ++i) (*V)(i, j) = v(i, idx[j]);
  }
}

void qtoR(const Quat& q, double R[9]) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

void skew(const double v[3], double S[9]) {
  S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}

void Qleft_br(const Quat& q, double M[9]) {
  double v[3] = {q.x, q.y, q.z}, S[9];
  skew(v, S);
  for (int i = 0; i < 9; ++i) M[i] = S[i];
  M[0] += q.w; M[4] += q.w; M[8] += q.w;
}

void Qright_br(const Quat& q, double M[9]) {
  double v[3] = {q.x, q.y, q.z}, S[9];
  skew(v, S);
  for (int i = 0; i < 9; ++i) M[i] = -S[i];
  M[0] += q.w; M[4] += q.w; M[8] += q.w;
}

void pose_plus(const double* x, const double* delta, double* out) {
  for (int i = 0; i < 3; ++i) out[i] = x[i] + delta[i];
  Quat q = pose_q(x);
  Quat dq = deltaQ(delta + 3);
  Quat r = qnormalized(qmul(q, dq));
  out[3] = r.x; out[4] = r.y; out[5] = r.z; out[6] = r.w;
}

void cauchy_loss(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * std::log(sum);
  rho[1] = std::max(std::numeric_limits<double>::min(), inv);
  rho[2] = -c * (inv * inv);
}

}  // namespace oracle
