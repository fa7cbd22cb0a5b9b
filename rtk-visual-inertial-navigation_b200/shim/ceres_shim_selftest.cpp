// Host-only checks of the ceres::Problem bookkeeping (no device): the semantics the reference
// relies on (CERES/internal/ceres/problem_impl.cc:280-478,886; problem_test.cc): pointer identity,
// cascading RemoveParameterBlock, residual-block index reuse, constness, ordering groups, is_use.
#include <cstdio>
#include <vector>

#include "ceres/ceres.h"
#include "ceres/schur_complement_solver.h"

namespace {
struct Unary : ceres::SizedCostFunction<1, 1> {
  bool Evaluate(double const* const*, double*, double**) const override { return true; }
};
struct Binary : ceres::SizedCostFunction<2, 3, 1> {
  bool Evaluate(double const* const*, double*, double**) const override { return true; }
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
    ceres::Solve(opt, &p, &s);  // Unary has no adapter registered
    EXPECT(s.termination_type == ceres::FAILURE && s.message.find("adapter") != std::string::npos);
  }
  EXPECT(ceres::internal::is_optimize == true && ceres::internal::parameter_head.empty());
  return 0;
}
