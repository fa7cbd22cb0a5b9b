// ceres/sized_cost_function.h -- shim of CERES/include/ceres/sized_cost_function.h.
#ifndef SWGN_CERES_SIZED_COST_FUNCTION_H_
#define SWGN_CERES_SIZED_COST_FUNCTION_H_
#include "ceres/cost_function.h"
namespace ceres {
template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    set_num_residuals(kNumResiduals);
    *mutable_parameter_block_sizes() = std::vector<int32_t>{Ns...};
  }
  virtual ~SizedCostFunction() {}
};
}  // namespace ceres
#endif
