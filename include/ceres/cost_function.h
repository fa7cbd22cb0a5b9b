// ceres/cost_function.h -- shim of CERES/include/ceres/cost_function.h:116-140.
#ifndef SWGN_CERES_COST_FUNCTION_H_
#define SWGN_CERES_COST_FUNCTION_H_
#include <cstdint>
#include <vector>
namespace ceres {
class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  CostFunction(const CostFunction&) = delete;
  void operator=(const CostFunction&) = delete;
  virtual ~CostFunction() {}
  // jacobians[i] is row-major num_residuals x parameter_block_sizes()[i] (GLOBAL size); jacobians
  // and any jacobians[i] may be null; return false = evaluation failure.
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int num_residuals) { num_residuals_ = num_residuals; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_;
};
}  // namespace ceres
#endif
