// ceres/swgn_adapter.h -- how application factor classes reach the device (shim extension).
//
// The CUDA solver cannot call virtual CostFunction::Evaluate; instead every recognised factor
// class is described to it as a device factor record read from the object's PUBLIC data members
// (factor code itself unchanged).  An adapter is registered per concrete type once per process:
//
//   ceres::swgn::RegisterAdapter(typeid(projection_factor), &adapt_projection);
//
// and fills a FactorRecord when ceres::Solve flattens the Problem.  Types without an adapter make
// Solve() return FAILURE with a message naming the type (the host-evaluated generic path for
// stateful factors such as IMUGNSSFactor is listed as "next" in SURVEY.md 8f).
#ifndef SWGN_CERES_SWGN_ADAPTER_H_
#define SWGN_CERES_SWGN_ADAPTER_H_
#include <typeindex>
#include <vector>

#include "ceres/cost_function.h"
#include "swgn.h"

namespace ceres {
namespace swgn {
enum FactorKind { kProjection = 0, kImu = 1, kGnss = 2, kPrior = 3, kUnit = 4 };
struct FactorRecord {
  int kind = -1;
  int gnss_kind = -1;            // SWGN_GNSS_* for kGnss
  std::vector<double> data;      // kProjection: uv[2]; kImu: SWGN_IMU_STRIDE; kGnss: SWGN_GNSS_STRIDE;
                                 // kUnit: istd; kPrior: see below
  // kPrior (dense linear factor r = r0 + J0 (x [-] x0)): rows n, per keep block its first tangent
  // column; x0 concatenated in parameter order (global sizes), J0 n x n row-major, r0[n]
  int prior_n = 0;
  std::vector<int> prior_blk_idx;
  std::vector<double> prior_x0, prior_J, prior_r0;
};
// application globals the device factors read (RVI/parameter/parameters.h:88,94,100; swf.cpp:47)
struct Globals {
  double Pbg[3] = {0, 0, 0};
  double gravity[3] = {0, 0, 0};            // Rwgw * G
  double proj_sqrt_info[4] = {1, 0, 0, 1};  // projection_factor::sqrt_info
};
typedef bool (*Adapter)(const CostFunction* cost_function, FactorRecord* out);
void RegisterAdapter(const std::type_index& type, Adapter adapter);
void SetGlobals(const Globals& g);
const Globals& GetGlobals();
}  // namespace swgn
}  // namespace ceres
#endif
