// ceres/schur_complement_solver.h -- the reference's non-standard side channel (M2:
// CERES/internal/ceres/schur_complement_solver.h:55-63, .cc:57-66,172-188,253-258).  The shim
// mirrors the globals per process after every Solve():
//   parameter_head non-empty and !is_optimize : lhs_out / rhs_out / hs_row = reduced system S, r
//                                               (row-major, upper triangle meaningful), state untouched
//   parameter_head non-empty and  is_optimize : lhs_out2 = lower Cholesky factor of the last reduced solve
#ifndef SWGN_CERES_SCHUR_COMPLEMENT_SOLVER_H_
#define SWGN_CERES_SCHUR_COMPLEMENT_SOLVER_H_
#include <vector>
namespace ceres {
namespace internal {
#define RHSROWLIMIT 1024
extern double lhs_out[RHSROWLIMIT * RHSROWLIMIT], rhs_out[RHSROWLIMIT];
extern double lhs_out2[RHSROWLIMIT * RHSROWLIMIT];
extern int hs_row;
extern bool is_optimize;
extern std::vector<double*> parameter_head;
extern std::vector<int> parameter_block_size;
}  // namespace internal
}  // namespace ceres
#endif
