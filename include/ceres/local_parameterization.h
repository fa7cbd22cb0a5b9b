// ceres/local_parameterization.h -- shim of CERES/include/ceres/local_parameterization.h:123-149.
#ifndef SWGN_CERES_LOCAL_PARAMETERIZATION_H_
#define SWGN_CERES_LOCAL_PARAMETERIZATION_H_
namespace ceres {
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual bool MultiplyByJacobian(const double* /*x*/, const int /*num_rows*/, const double* /*global_matrix*/,
                                  double* /*local_matrix*/) const {
    return false;
  }
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
}  // namespace ceres
#endif
