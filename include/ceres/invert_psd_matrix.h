// ceres/invert_psd_matrix.h -- ceres::internal::InvertPSDMatrix as the application calls it
// (RVI/factor/gnss_imu_factor.cpp:404, kSize = 15, assume_full_rank = true).  Semantics of
// CERES/internal/ceres/invert_psd_matrix.h:50-74: full rank -> Eigen's inverse() for fixed sizes below 5,
// otherwise LLT of the upper triangle solved against the identity; rank deficient -> thin-SVD solve.
// Needs Eigen (the application's build has it; this repository's own build does not and compiles this to nothing).
// Exercised by oracle/build_ref.sh, which compiles the reference's gnss_imu_factor.cpp against it and a stand-in Eigen.
// The device-side counterpart used by the solver is the LLT-solve-identity in csrc/k_chain.cu.
#ifndef SWGN_CERES_INVERT_PSD_MATRIX_H_
#define SWGN_CERES_INVERT_PSD_MATRIX_H_
#include "ceres/internal/eigen.h"
#ifdef SWGN_HAVE_EIGEN
#include <Eigen/Cholesky>
#include <Eigen/LU>
#include <Eigen/SVD>
namespace ceres {
namespace internal {
template <int kSize>
typename EigenTypes<kSize, kSize>::Matrix InvertPSDMatrix(const bool assume_full_rank, const typename EigenTypes<kSize, kSize>::Matrix& m) {
  using Square = typename EigenTypes<kSize, kSize>::Matrix;
  const int n = static_cast<int>(m.rows());
  const Square identity = Square::Identity(n, n);
  if (!assume_full_rank) {
    using Thin = typename EigenTypes<kSize, Eigen::Dynamic>::Matrix;  // thin SVD wants a dynamic column count
    return Eigen::JacobiSVD<Thin>(m, Eigen::ComputeThinU | Eigen::ComputeThinV).solve(identity);
  }
  if (kSize > 0 && kSize < 5) return m.inverse();
  return m.template selfadjointView<Eigen::Upper>().llt().solve(identity);
}
}  // namespace internal
}  // namespace ceres
#endif
#endif
