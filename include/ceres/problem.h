// ceres/problem.h -- shim of CERES/include/ceres/problem.h:122-551: the long-lived, mutable factor
// graph the reference keeps in `ceres::Problem my_problem` (RVI/swf/swf.h:246).  Parameter blocks
// are identified by raw double* identity and the solver reads / writes user memory in place.
// Host-only bookkeeping; ceres::Solve (solver.h) flattens it into a swgn_graph and runs the CUDA
// solver through the C ABI.  Semantics follow CERES/internal/ceres/problem_impl.cc:280-478,886.
#ifndef SWGN_CERES_PROBLEM_H_
#define SWGN_CERES_PROBLEM_H_
#include <map>
#include <memory>
#include <set>
#include <unordered_map>
#include <vector>

#include "ceres/cost_function.h"
#include "ceres/local_parameterization.h"
#include "ceres/loss_function.h"
#include "ceres/residual_block.h"
#include "ceres/types.h"

namespace ceres {
class Problem {
 public:
  struct Options {
    Ownership cost_function_ownership = TAKE_OWNERSHIP;
    Ownership loss_function_ownership = TAKE_OWNERSHIP;
    Ownership local_parameterization_ownership = TAKE_OWNERSHIP;
    bool enable_fast_removal = true;  // problem.h:150 (the shim always keeps the reverse index)
    bool disable_all_safety_checks = false;
  };
  struct ParameterBlockInfo {
    int size = 0;
    bool constant = false;
    LocalParameterization* parameterization = nullptr;
    int index = 0;  // insertion order (ties inside an ordering group are broken by it)
    std::set<internal::ResidualBlock*> residual_blocks;
  };

  Problem() {}
  explicit Problem(const Options& options) : options_(options) {}
  Problem(const Problem&) = delete;
  void operator=(const Problem&) = delete;
  ~Problem();

  ResidualBlockId AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function,
                                   const std::vector<double*>& parameter_blocks);
  template <typename... Ts>
  ResidualBlockId AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function, double* x0, Ts*... xs) {
    return AddResidualBlock(cost_function, loss_function, std::vector<double*>{x0, xs...});
  }
  void AddParameterBlock(double* values, int size);
  void AddParameterBlock(double* values, int size, LocalParameterization* local_parameterization);
  void RemoveParameterBlock(const double* values);  // cascades to the dependent residual blocks
  void RemoveResidualBlock(ResidualBlockId residual_block);
  void SetParameterBlockConstant(const double* values);
  void SetParameterBlockVariable(double* values);
  bool IsParameterBlockConstant(const double* values) const;
  void SetParameterization(double* values, LocalParameterization* local_parameterization);
  const LocalParameterization* GetParameterization(const double* values) const;
  bool HasParameterBlock(const double* values) const { return blocks_.count(const_cast<double*>(values)) > 0; }
  int ParameterBlockSize(const double* values) const;
  int ParameterBlockLocalSize(const double* values) const;
  int NumParameterBlocks() const { return (int)blocks_.size(); }
  int NumParameters() const;
  int NumResidualBlocks() const { return (int)residual_blocks_.size(); }
  int NumResiduals() const;
  void GetParameterBlocks(std::vector<double*>* parameter_blocks) const;
  void GetResidualBlocks(std::vector<ResidualBlockId>* residual_blocks) const;
  void GetParameterBlocksForResidualBlock(const ResidualBlockId residual_block, std::vector<double*>* parameter_blocks) const;
  const CostFunction* GetCostFunctionForResidualBlock(const ResidualBlockId residual_block) const { return residual_block->cost_function(); }
  const LossFunction* GetLossFunctionForResidualBlock(const ResidualBlockId residual_block) const { return residual_block->loss_function(); }
  void GetResidualBlocksForParameterBlock(const double* values, std::vector<ResidualBlockId>* residual_blocks) const;

  // --- used by ceres::Solve
  const std::unordered_map<double*, ParameterBlockInfo>& parameter_block_map() const { return blocks_; }
  const std::vector<internal::ResidualBlock*>& residual_block_list() const { return residual_blocks_; }

 private:
  void Fatal(const char* what) const;  // CHECK-failure of the original: message + abort
  void Release(const CostFunction* c);
  void Release(const LossFunction* l);
  Options options_;
  std::unordered_map<double*, ParameterBlockInfo> blocks_;
  std::vector<internal::ResidualBlock*> residual_blocks_;  // program order = AddResidualBlock order
  std::map<const void*, int> cost_refs_, loss_refs_, param_refs_;
  int next_block_index_ = 0;
};
}  // namespace ceres
#endif
