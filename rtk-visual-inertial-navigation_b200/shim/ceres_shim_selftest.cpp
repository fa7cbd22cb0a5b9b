// Host-only checks of the ceres::Problem bookkeeping (no device): the semantics the reference
// relies on (CERES/internal/ceres/problem_impl.cc:280-478,886; problem_test.cc): pointer identity,
// cascading RemoveParameterBlock, residual-block index reuse, constness, ordering groups, is_use.
#include <cmath>
#include <cstdio>
#include <vector>

#include "ceres/ceres.h"
#include "ceres/schur_complement_solver.h"
#include "ceres/small_blas.h"

namespace {
struct Unary : ceres::SizedCostFunction<1, 1> {
  bool Evaluate(double const* const*, double*, double**) const override { return true; }
};
struct Binary : ceres::SizedCostFunction<2, 3, 1> {
  bool Evaluate(double const* const*, double*, double**) const override { return true; }
};
struct Scaled : ceres::SizedCostFunction<1, 1> {  // r = 3 x
  bool Evaluate(double const* const* p, double* r, double**) const override {
    r[0] = 3.0 * p[0][0];
    return true;
  }
};
int g_deleted = 0;
struct Counted : ceres::SizedCostFunction<1, 1> {
  ~Counted() override { ++g_deleted; }
  bool Evaluate(double const* const*, double*, double**) const override { return true; }
};
#define EXPECT(c)                                             \
  do {                                                        \
    if (!(c)) {                                               \
      std::fprintf(stderr, "FAILED %s:%d %s\n", __FILE__, __LINE__, #c); \
      return __LINE__;                                        \
    }                                                         \
  } while (0)
}  // namespace

extern "C" int swgn_ceres_selftest() {
  double x[3] = {1, 2, 3}, y[1] = {4}, z[1] = {5};
  {
    ceres::Problem p;
    p.AddParameterBlock(x, 3);
    EXPECT(p.HasParameterBlock(x) && !p.HasParameterBlock(y));
    ceres::ResidualBlockId r0 = p.AddResidualBlock(new Binary, nullptr, x, y);
    ceres::ResidualBlockId r1 = p.AddResidualBlock(new Unary, new ceres::CauchyLoss(1.0), y);
    ceres::ResidualBlockId r2 = p.AddResidualBlock(new Unary, nullptr, z);
    EXPECT(p.NumParameterBlocks() == 3 && p.NumResidualBlocks() == 3 && p.NumResiduals() == 4 && p.NumParameters() == 5);
    EXPECT(p.ParameterBlockSize(x) == 3 && p.ParameterBlockLocalSize(y) == 1);
    EXPECT(r0->is_use && r1->is_use);
    r1->is_use = false;  // the application toggles the mask through the id (swf_image.cpp:353-358)
    std::vector<ceres::ResidualBlockId> rs;
    p.GetResidualBlocksForParameterBlock(y, &rs);
    EXPECT(rs.size() == 2 && rs[0] == r0 && rs[1] == r1);
    std::vector<double*> ps;
    p.GetParameterBlocksForResidualBlock(r0, &ps);
    EXPECT(ps.size() == 2 && ps[0] == x && ps[1] == y);
    p.SetParameterBlockConstant(x);
    EXPECT(p.IsParameterBlockConstant(x) && !p.IsParameterBlockConstant(y));
    p.SetParameterBlockVariable(x);
    EXPECT(!p.IsParameterBlockConstant(x));
    // removing y cascades to r0 and r1 (problem_impl.cc:440-476)
    p.RemoveParameterBlock(y);
    EXPECT(!p.HasParameterBlock(y) && p.NumResidualBlocks() == 1 && p.NumParameterBlocks() == 2);
    p.GetResidualBlocks(&rs);
    EXPECT(rs.size() == 1 && rs[0] == r2);
    p.GetResidualBlocksForParameterBlock(x, &rs);
    EXPECT(rs.empty());
    p.RemoveResidualBlock(r2);
    EXPECT(p.NumResidualBlocks() == 0 && p.HasParameterBlock(z));
  }
  {  // ownership: a cost function shared by two residual blocks is deleted once, when the last goes
    ceres::Problem p;
    Counted* c = new Counted;
    g_deleted = 0;
    ceres::ResidualBlockId a = p.AddResidualBlock(c, nullptr, y);
    ceres::ResidualBlockId b = p.AddResidualBlock(c, nullptr, z);
    p.RemoveResidualBlock(a);
    EXPECT(g_deleted == 0);
    p.RemoveResidualBlock(b);
    EXPECT(g_deleted == 1);
  }
  {  // ordering groups (ordered_groups.h)
    ceres::ParameterBlockOrdering o;
    EXPECT(o.AddElementToGroup(x, 0) && o.AddElementToGroup(y, 2) && o.AddElementToGroup(z, 2));
    EXPECT(o.NumElements() == 3 && o.NumGroups() == 2 && o.GroupId(y) == 2 && o.GroupId(nullptr) == -1);
    EXPECT(o.AddElementToGroup(y, 0) && o.GroupSize(0) == 2 && o.GroupSize(2) == 1);
    EXPECT(o.Remove(z) && o.NumGroups() == 1);
    o.Clear();
    EXPECT(o.NumElements() == 0);
  }
  {  // Solve refuses configurations the device path does not implement, without touching the state
    ceres::Problem p;
    p.AddResidualBlock(new Unary, nullptr, y);
    ceres::Solver::Options opt;  // defaults: DENSE_QR + LM
    ceres::Solver::Summary s;
    ceres::Solve(opt, &p, &s);
    EXPECT(s.termination_type == ceres::FAILURE && y[0] == 4);
    opt.linear_solver_type = ceres::DENSE_SCHUR;
    opt.trust_region_strategy_type = ceres::DOGLEG;
    opt.jacobi_scaling = false;
    opt.linear_solver_ordering = std::make_shared<ceres::ParameterBlockOrdering>();
    opt.linear_solver_ordering->AddElementToGroup(y, 0);
    ceres::Solve(opt, &p, &s);  // Unary has no adapter: host-evaluated, so the call reaches the device layer
    EXPECT(s.termination_type != ceres::FAILURE || s.message.find("adapter") == std::string::npos);
  }
  {  // a problem without a variable parameter block converges at once with the fixed cost, evaluated by the user's own
     // cost and loss functions, like Ceres (solver.cc: "No non-constant parameter blocks found"); no device is touched
    ceres::Problem p;
    double a[1] = {2.0}, b[1] = {-1.0};
    p.AddResidualBlock(new Scaled, nullptr, a);
    p.AddResidualBlock(new Scaled, new ceres::CauchyLoss(1.0), b);
    p.SetParameterBlockConstant(a);
    p.SetParameterBlockConstant(b);
    ceres::Solver::Options opt;
    opt.linear_solver_type = ceres::DENSE_SCHUR;
    opt.linear_solver_ordering = std::make_shared<ceres::ParameterBlockOrdering>();
    opt.linear_solver_ordering->AddElementToGroup(a, 0);
    opt.linear_solver_ordering->AddElementToGroup(b, 1);
    ceres::Solver::Summary s;
    ceres::Solve(opt, &p, &s);
    const double want = 0.5 * 36.0 + 0.5 * std::log(1.0 + 9.0);
    EXPECT(s.termination_type == ceres::CONVERGENCE && s.IsSolutionUsable());
    EXPECT(std::fabs(s.fixed_cost - want) < 1e-14 && s.initial_cost == s.fixed_cost && s.final_cost == s.fixed_cost);
    EXPECT(a[0] == 2.0 && b[0] == -1.0);
  }
  EXPECT(ceres::internal::is_optimize == true && ceres::internal::parameter_head.empty());
  return 0;
}

// ---- ceres/small_blas.h against the cases of the reference's own small_blas_test.cc (:74-330): A(i,j) = B(i,j) =
// i + j + 1, C = ones of every size row_stride x col_stride from the product's size up to three times it, the
// product placed at every admissible (start_row_c, start_col_c), kOperation = +1 / -1 / 0, sizes (5,3,7), (1,1,1),
// (9,9,9), fixed and dynamic (-1 = Eigen::Dynamic) template arguments; the values are small integers, so the
// comparison with the triple loop is exact.  Matrix-vector forms: small_blas_test.cc:332-end.
namespace {
template <int kRowA, int kColA, int kColB, bool kDynamic, bool kTranspose>
int small_blas_case() {
  // plain product: A is kRowA x kColA, B is kColA x kColB, C block kRowA x kColB
  // transposed:    A is kRowA x kColA, B is kRowA x kColB, C block kColA x kColB (A' B)
  const int rows_b = kTranspose ? kRowA : kColA;
  const int rows_c = kTranspose ? kColA : kRowA;
  std::vector<double> A(kRowA * kColA), B(rows_b * kColB);
  for (int i = 0; i < kRowA; ++i)
    for (int j = 0; j < kColA; ++j) A[i * kColA + j] = i + j + 1;
  for (int i = 0; i < rows_b; ++i)
    for (int j = 0; j < kColB; ++j) B[i * kColB + j] = i + j + 1;
  std::vector<double> P(rows_c * kColB, 0.0);
  for (int i = 0; i < rows_c; ++i)
    for (int j = 0; j < kColB; ++j)
      for (int k = 0; k < (kTranspose ? kRowA : kColA); ++k)
        P[i * kColB + j] += (kTranspose ? A[k * kColA + i] : A[i * kColA + k]) * B[k * kColB + j];
  constexpr int D = -1;
  for (int rs = rows_c; rs < 3 * rows_c; ++rs)
    for (int cs = kColB; cs < 3 * kColB; ++cs) {
      std::vector<double> Cp(rs * cs, 1.0), Cm(rs * cs, 1.0), Ca(rs * cs, 1.0);
      std::vector<double> Rp = Cp, Rm = Cm, Ra = Ca;
      for (int r0 = 0; r0 + rows_c < rs; ++r0)
        for (int c0 = 0; c0 + kColB < cs; ++c0) {
          for (int i = 0; i < rows_c; ++i)
            for (int j = 0; j < kColB; ++j) {
              Rp[(r0 + i) * cs + c0 + j] += P[i * kColB + j];
              Rm[(r0 + i) * cs + c0 + j] -= P[i * kColB + j];
              Ra[(r0 + i) * cs + c0 + j] = P[i * kColB + j];
            }
          using namespace ceres::internal;
          if (!kTranspose) {
            if (kDynamic) {
              MatrixMatrixMultiply<D, D, D, D, 1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cp.data(), r0, c0, rs, cs);
              MatrixMatrixMultiply<D, D, D, D, -1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cm.data(), r0, c0, rs, cs);
              MatrixMatrixMultiply<D, D, D, D, 0>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Ca.data(), r0, c0, rs, cs);
            } else {
              MatrixMatrixMultiply<kRowA, kColA, kColA, kColB, 1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cp.data(), r0, c0, rs, cs);
              MatrixMatrixMultiply<kRowA, kColA, kColA, kColB, -1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cm.data(), r0, c0, rs, cs);
              MatrixMatrixMultiply<kRowA, kColA, kColA, kColB, 0>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Ca.data(), r0, c0, rs, cs);
            }
          } else {
            if (kDynamic) {
              MatrixTransposeMatrixMultiply<D, D, D, D, 1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cp.data(), r0, c0, rs, cs);
              MatrixTransposeMatrixMultiply<D, D, D, D, -1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cm.data(), r0, c0, rs, cs);
              MatrixTransposeMatrixMultiply<D, D, D, D, 0>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Ca.data(), r0, c0, rs, cs);
            } else {
              MatrixTransposeMatrixMultiply<kRowA, kColA, kRowA, kColB, 1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cp.data(), r0, c0, rs, cs);
              MatrixTransposeMatrixMultiply<kRowA, kColA, kRowA, kColB, -1>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Cm.data(), r0, c0, rs, cs);
              MatrixTransposeMatrixMultiply<kRowA, kColA, kRowA, kColB, 0>(A.data(), kRowA, kColA, B.data(), rows_b, kColB, Ca.data(), r0, c0, rs, cs);
            }
          }
          EXPECT(Cp == Rp);
          EXPECT(Cm == Rm);
          EXPECT(Ca == Ra);
        }
    }
  return 0;
}
template <int kRowA, int kColA>
int small_blas_vector_case() {
  std::vector<double> A(kRowA * kColA), b(kColA), bt(kRowA);
  for (int i = 0; i < kRowA; ++i)
    for (int j = 0; j < kColA; ++j) A[i * kColA + j] = i + j + 1;
  for (int j = 0; j < kColA; ++j) b[j] = 2 * j + 1;
  for (int i = 0; i < kRowA; ++i) bt[i] = 3 * i + 2;
  std::vector<double> Ab(kRowA, 0.0), Atb(kColA, 0.0);
  for (int i = 0; i < kRowA; ++i)
    for (int j = 0; j < kColA; ++j) {
      Ab[i] += A[i * kColA + j] * b[j];
      Atb[j] += A[i * kColA + j] * bt[i];
    }
  using namespace ceres::internal;
  std::vector<double> cp(kRowA, 1.0), cm(kRowA, 1.0), ca(kRowA, 1.0), dp(kColA, 1.0), dm(kColA, 1.0), da(kColA, 1.0);
  MatrixVectorMultiply<kRowA, kColA, 1>(A.data(), kRowA, kColA, b.data(), cp.data());
  MatrixVectorMultiply<-1, -1, -1>(A.data(), kRowA, kColA, b.data(), cm.data());
  MatrixVectorMultiply<kRowA, kColA, 0>(A.data(), kRowA, kColA, b.data(), ca.data());
  MatrixTransposeVectorMultiply<kRowA, kColA, 1>(A.data(), kRowA, kColA, bt.data(), dp.data());
  MatrixTransposeVectorMultiply<-1, -1, -1>(A.data(), kRowA, kColA, bt.data(), dm.data());
  MatrixTransposeVectorMultiply<kRowA, kColA, 0>(A.data(), kRowA, kColA, bt.data(), da.data());
  for (int i = 0; i < kRowA; ++i) EXPECT(cp[i] == 1.0 + Ab[i] && cm[i] == 1.0 - Ab[i] && ca[i] == Ab[i]);
  for (int j = 0; j < kColA; ++j) EXPECT(dp[j] == 1.0 + Atb[j] && dm[j] == 1.0 - Atb[j] && da[j] == Atb[j]);
  return 0;
}
}  // namespace

extern "C" int swgn_small_blas_selftest() {
  int rc = 0;
#define RUN(...)                     \
  do {                               \
    if ((rc = __VA_ARGS__())) return rc; \
  } while (0)
  RUN(small_blas_case<5, 3, 7, false, false>);
  RUN(small_blas_case<5, 3, 7, true, false>);
  RUN(small_blas_case<1, 1, 1, false, false>);
  RUN(small_blas_case<1, 1, 1, true, false>);
  RUN(small_blas_case<9, 9, 9, false, false>);
  RUN(small_blas_case<9, 9, 9, true, false>);
  RUN(small_blas_case<5, 3, 7, false, true>);
  RUN(small_blas_case<5, 3, 7, true, true>);
  RUN(small_blas_case<1, 1, 1, false, true>);
  RUN(small_blas_case<1, 1, 1, true, true>);
  RUN(small_blas_case<9, 9, 9, false, true>);
  RUN(small_blas_case<9, 9, 9, true, true>);
  RUN(small_blas_vector_case<5, 3>);
  RUN(small_blas_vector_case<1, 1>);
  RUN(small_blas_vector_case<9, 9>);
  RUN(small_blas_vector_case<15, 7>);
#undef RUN
  return 0;
}
