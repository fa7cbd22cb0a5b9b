// ceres/small_blas.h -- the subset of CERES/internal/ceres/small_blas.h the application calls
// directly from its GNSS-IMU factor (RVI/factor/gnss_imu_factor.cpp:358-429,529-616): plain
// row-major small matrix products with kOperation = +1 (add), -1 (subtract), 0 (assign).
#ifndef SWGN_CERES_SMALL_BLAS_H_
#define SWGN_CERES_SMALL_BLAS_H_
namespace ceres {
namespace internal {
// C(r0.., c0..) op= A * B
template <int kRowA, int kColA, int kRowB, int kColB, int kOperation>
inline void MatrixMatrixMultiply(const double* A, const int num_row_a, const int num_col_a, const double* B,
                                 const int /*num_row_b*/, const int num_col_b, double* C, const int start_row_c,
                                 const int start_col_c, const int /*row_stride_c*/, const int col_stride_c) {
  for (int i = 0; i < num_row_a; ++i)
    for (int j = 0; j < num_col_b; ++j) {
      double t = 0.0;
      for (int k = 0; k < num_col_a; ++k) t += A[i * num_col_a + k] * B[k * num_col_b + j];
      double& c = C[(i + start_row_c) * col_stride_c + start_col_c + j];
      if (kOperation > 0) c += t;
      else if (kOperation < 0) c -= t;
      else c = t;
    }
}
// C(r0.., c0..) op= A' * B
template <int kRowA, int kColA, int kRowB, int kColB, int kOperation>
inline void MatrixTransposeMatrixMultiply(const double* A, const int num_row_a, const int num_col_a, const double* B,
                                          const int /*num_row_b*/, const int num_col_b, double* C,
                                          const int start_row_c, const int start_col_c, const int /*row_stride_c*/,
                                          const int col_stride_c) {
  for (int i = 0; i < num_col_a; ++i)
    for (int j = 0; j < num_col_b; ++j) {
      double t = 0.0;
      for (int k = 0; k < num_row_a; ++k) t += A[k * num_col_a + i] * B[k * num_col_b + j];
      double& c = C[(i + start_row_c) * col_stride_c + start_col_c + j];
      if (kOperation > 0) c += t;
      else if (kOperation < 0) c -= t;
      else c = t;
    }
}
template <int kRowA, int kColA, int kOperation>
inline void MatrixVectorMultiply(const double* A, const int num_row_a, const int num_col_a, const double* b, double* c) {
  for (int i = 0; i < num_row_a; ++i) {
    double t = 0.0;
    for (int k = 0; k < num_col_a; ++k) t += A[i * num_col_a + k] * b[k];
    if (kOperation > 0) c[i] += t;
    else if (kOperation < 0) c[i] -= t;
    else c[i] = t;
  }
}
template <int kRowA, int kColA, int kOperation>
inline void MatrixTransposeVectorMultiply(const double* A, const int num_row_a, const int num_col_a, const double* b,
                                          double* c) {
  for (int i = 0; i < num_col_a; ++i) {
    double t = 0.0;
    for (int k = 0; k < num_row_a; ++k) t += A[k * num_col_a + i] * b[k];
    if (kOperation > 0) c[i] += t;
    else if (kOperation < 0) c[i] -= t;
    else c[i] = t;
  }
}
}  // namespace internal
}  // namespace ceres
#endif
