// ceres/loss_function.h -- shim of CERES/include/ceres/loss_function.h:114,207-217.
// The device path applies CauchyLoss (the only loss the reference puts into the window problem,
// RVI/swf/swf_image.cpp:98-100); other losses make Solve() report an unsupported problem.
#ifndef SWGN_CERES_LOSS_FUNCTION_H_
#define SWGN_CERES_LOSS_FUNCTION_H_
#include <algorithm>
#include <cmath>
#include <limits>
namespace ceres {
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
class TrivialLoss : public LossFunction {
 public:
  void Evaluate(double s, double rho[3]) const override {
    rho[0] = s;
    rho[1] = 1.0;
    rho[2] = 0.0;
  }
};
class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : a_(a), b_(a * a), c_(1 / (a * a)) {}
  void Evaluate(double s, double rho[3]) const override {  // CERES/internal/ceres/loss_function.cc:73-80
    const double sum = 1.0 + s * c_;
    const double inv = 1.0 / sum;
    rho[0] = b_ * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c_ * (inv * inv);
  }
  double a() const { return a_; }

 private:
  const double a_, b_, c_;
};
}  // namespace ceres
#endif
