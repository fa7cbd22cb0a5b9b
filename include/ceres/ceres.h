// ceres/ceres.h -- umbrella header of the source-compatibility shim.
#ifndef SWGN_CERES_CERES_H_
#define SWGN_CERES_CERES_H_
#include "ceres/cost_function.h"
#include "ceres/internal/eigen.h"  // ceres::Matrix / ConstMatrixRef where Eigen is installed (RVI/swf/swf_gnss.cpp:28,85); empty otherwise
#include "ceres/local_parameterization.h"
#include "ceres/loss_function.h"
#include "ceres/ordered_groups.h"
#include "ceres/problem.h"
#include "ceres/sized_cost_function.h"
#include "ceres/solver.h"
#include "ceres/types.h"
#endif
