// ceres/types.h -- source-compatibility shim over the swgn C ABI (include/swgn.h).
// Mirrors the subset of CERES/include/ceres/types.h the reference application uses
// (RVI/swf/swf.cpp:25-30, swf_gnss.cpp:204-215).  Host-only; no Eigen, no glog.
#ifndef SWGN_CERES_TYPES_H_
#define SWGN_CERES_TYPES_H_
namespace ceres {
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum DoglegType { TRADITIONAL_DOGLEG, SUBSPACE_DOGLEG };
enum MinimizerType { LINE_SEARCH, TRUST_REGION };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
enum LoggingType { SILENT, PER_MINIMIZER_ITERATION };
}  // namespace ceres
#endif
