// ceres/solver.h -- shim of CERES/include/ceres/solver.h:233-737 (Options), Summary and
// ceres::Solve (CERES/internal/ceres/solver.cc:604): the fields the reference sets and reads
// (RVI/swf/swf.cpp:25-30, swf_image.cpp:212-230, swf_core.cpp:409).  The device path implements the
// reference's two configurations: DENSE_SCHUR with a user ordering and either TRADITIONAL DOGLEG with
// jacobi_scaling = false (the sliding-window solve) or LEVENBERG_MARQUARDT with / without jacobi_scaling
// (Ceres' defaults: the per-epoch GNSS solves, swf_gnss.cpp:204-215,563-573); max_solver_time_in_seconds is
// accepted and ignored (the device solve is bounded by max_num_iterations).
#ifndef SWGN_CERES_SOLVER_H_
#define SWGN_CERES_SOLVER_H_
#include <memory>
#include <string>

#include "ceres/ordered_groups.h"
#include "ceres/problem.h"
#include "ceres/types.h"

namespace ceres {
class Solver {
 public:
  struct Options {
    MinimizerType minimizer_type = TRUST_REGION;
    TrustRegionStrategyType trust_region_strategy_type = LEVENBERG_MARQUARDT;  // solver.h:233
    DoglegType dogleg_type = TRADITIONAL_DOGLEG;
    bool use_nonmonotonic_steps = false;
    int max_num_iterations = 50;
    double max_solver_time_in_seconds = 1e9;
    int num_threads = 1;
    double initial_trust_region_radius = 1e4;
    double max_trust_region_radius = 1e16;
    double min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3;
    double min_lm_diagonal = 1e-6;
    double max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    double function_tolerance = 1e-6;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    LinearSolverType linear_solver_type = DENSE_QR;
    std::shared_ptr<ParameterBlockOrdering> linear_solver_ordering;
    bool jacobi_scaling = true;
    LoggingType logging_type = PER_MINIMIZER_ITERATION;
    bool minimizer_progress_to_stdout = false;
    bool update_state_every_iteration = false;
    int device = 0;  // shim extension: CUDA device ordinal
  };
  struct Summary {
    std::string BriefReport() const;
    std::string FullReport() const { return BriefReport(); }
    bool IsSolutionUsable() const { return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE || termination_type == USER_SUCCESS; }
    TerminationType termination_type = FAILURE;
    std::string message = "ceres::Solve was not called.";
    double initial_cost = -1, final_cost = -1, fixed_cost = -1;
    int num_successful_steps = -1, num_unsuccessful_steps = -1;
    int num_linear_solves = -1;
    double preprocessor_time_in_seconds = -1, minimizer_time_in_seconds = -1, total_time_in_seconds = -1;
    int num_parameter_blocks = -1, num_residual_blocks = -1, num_residuals = -1;
    int num_parameter_blocks_reduced = -1, num_residuals_reduced = -1;
  };
};
void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary);
}  // namespace ceres
#endif
