// ceres/residual_block.h -- the reference installs this INTERNAL header publicly (M1) and the
// application dereferences ResidualBlockId for the per-residual mask `is_use` (M3:
// CERES/internal/ceres/residual_block.h:135; RVI/swf/swf_gnss.cpp:653, swf_image.cpp:353-358).
#ifndef SWGN_CERES_RESIDUAL_BLOCK_H_
#define SWGN_CERES_RESIDUAL_BLOCK_H_
#include <vector>
namespace ceres {
class CostFunction;
class LossFunction;
namespace internal {
class ResidualBlock {
 public:
  ResidualBlock(const CostFunction* cost, const LossFunction* loss, const std::vector<double*>& params, int index)
      : is_use(true), cost_function_(cost), loss_function_(loss), parameter_blocks_(params), index_(index) {}
  const CostFunction* cost_function() const { return cost_function_; }
  const LossFunction* loss_function() const { return loss_function_; }
  const std::vector<double*>& parameter_blocks() const { return parameter_blocks_; }
  int NumParameterBlocks() const { return (int)parameter_blocks_.size(); }
  int index() const { return index_; }
  void set_index(int i) { index_ = i; }
  bool is_use;  // false: dropped from the reduced program, cost still counted in fixed_cost

 private:
  const CostFunction* cost_function_;
  const LossFunction* loss_function_;
  std::vector<double*> parameter_blocks_;
  int index_;
};
}  // namespace internal
typedef internal::ResidualBlock* ResidualBlockId;
}  // namespace ceres
#endif
