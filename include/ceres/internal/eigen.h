// ceres/internal/eigen.h -- the Eigen aliases the application uses by their Ceres names
// (ceres::Matrix / ceres::ConstMatrixRef in RVI/swf/swf_gnss.cpp:28,85; EigenTypes<> through
// InvertPSDMatrix in RVI/factor/gnss_imu_factor.cpp:404).  Only meaningful where Eigen is installed
// (the application's build); this repository's own build has no Eigen and never includes it.
#ifndef SWGN_CERES_INTERNAL_EIGEN_H_
#define SWGN_CERES_INTERNAL_EIGEN_H_
#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#define SWGN_HAVE_EIGEN 1
#include <Eigen/Core>
namespace ceres {
using Vector = Eigen::Matrix<double, Eigen::Dynamic, 1>;
using Matrix = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;  // row-major, like every J / S buffer here
using VectorRef = Eigen::Map<Vector>;
using MatrixRef = Eigen::Map<Matrix>;
using ConstVectorRef = Eigen::Map<const Vector>;
using ConstMatrixRef = Eigen::Map<const Matrix>;
using ColMajorMatrix = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::ColMajor>;
// statically sized variants; a single column cannot be row-major in Eigen
template <int kRows = Eigen::Dynamic, int kCols = Eigen::Dynamic>
struct EigenTypes {
  using Matrix = Eigen::Matrix<double, kRows, kCols, (kCols == 1 ? Eigen::ColMajor : Eigen::RowMajor)>;
  using MatrixRef = Eigen::Map<Matrix>;
  using ConstMatrixRef = Eigen::Map<const Matrix>;
  using Vector = Eigen::Matrix<double, kRows, 1>;
  using VectorRef = Eigen::Map<Vector>;
  using ConstVectorRef = Eigen::Map<const Vector>;
};
}  // namespace ceres
#endif
#endif
#endif
