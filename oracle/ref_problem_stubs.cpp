// TEST INFRASTRUCTURE.  libref_gnss.so holds the reference's factor classes without a ceres::Problem implementation; the
// two Problem methods gnss_imu_factor.cpp references (only behind USE_GLOBAL_OPTIMIZATION, which oracle/ref_globals.cpp
// leaves false) are defined here so that the library links.  libswgn_refdemo.so links the real shim instead.
#include <cstdio>
#include <cstdlib>

#include "ceres/problem.h"
namespace ceres {
ResidualBlockId Problem::AddResidualBlock(CostFunction*, LossFunction*, const std::vector<double*>&) {
  std::fprintf(stderr, "libref_gnss.so: ceres::Problem is not available in this library\n");
  std::abort();
}
void Problem::RemoveParameterBlock(const double*) {
  std::fprintf(stderr, "libref_gnss.so: ceres::Problem is not available in this library\n");
  std::abort();
}
}  // namespace ceres
