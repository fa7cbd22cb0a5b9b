// ceres/swgn_adapter.h -- how application factor classes reach the device (shim extension).
//
// The CUDA solver cannot call virtual CostFunction::Evaluate; instead every recognised factor
// class is described to it as a device factor record read from the object's PUBLIC data members
// (factor code itself unchanged).  An adapter is registered per concrete type once per process:
//
//   ceres::swgn::RegisterAdapter(typeid(projection_factor), &adapt_projection);
//
// and fills a FactorRecord when ceres::Solve flattens the Problem.  Types without an adapter are HOST-EVALUATED (kHost):
// their own Evaluate() is called on the host at every evaluation point and the residuals / Jacobians are uploaded --
// correct for any CostFunction (the reference's initialisation factors), at two host round trips per iteration.  The stateful IMUGNSSFactor is described
// as a chain record (kChain); its hidden GNSS-frame states are written back into the user memory
// the factor points at (gnss_poses[i], gnss_speed_bias[i]) when Solve returns.
#ifndef SWGN_CERES_SWGN_ADAPTER_H_
#define SWGN_CERES_SWGN_ADAPTER_H_
#include <typeindex>
#include <vector>

#include "ceres/cost_function.h"
#include "swgn.h"

namespace ceres {
namespace swgn {
enum FactorKind { kProjection = 0, kImu = 1, kGnss = 2, kPrior = 3, kUnit = 4, kChain = 5, kHost = 6, kNumKinds = 7 };
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
  // kChain (IMUGNSSFactor, RVI/factor/gnss_imu_factor.h): m hidden frames and k phase biases (the
  // residual block's parameters are pose_i, sb_i, pose_j, sb_j, N_0..N_{k-1}); arrays exactly as the
  // chain_* fields of swgn_graph; chain_pose_ptr / chain_sb_ptr are the user arrays of the hidden
  // frames (7 and 9 doubles), updated after the solve
  int chain_m = 0;
  std::vector<double> chain_frames, chain_frame_N, chain_N, chain_imu;
  std::vector<double*> chain_pose_ptr, chain_sb_ptr;
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
