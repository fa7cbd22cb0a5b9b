// Implementation of the ceres:: source-compatibility shim (include/ceres/*.h) on top of the swgn
// C ABI.  Host-side bookkeeping only: the factor graph store (CERES problem_impl.cc:280-478,886)
// and the flattening of one Solve() call into a swgn_graph.  All arithmetic of the solve runs in
// the CUDA library; if it is unavailable Solve() reports FAILURE -- there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <typeinfo>
#include <unordered_map>

#include "ceres/ceres.h"
#include "ceres/schur_complement_solver.h"
#include "ceres/swgn_adapter.h"
#include "swgn.h"

namespace ceres {
namespace internal {
double lhs_out[RHSROWLIMIT * RHSROWLIMIT], rhs_out[RHSROWLIMIT];
double lhs_out2[RHSROWLIMIT * RHSROWLIMIT];
int hs_row = 0;
bool is_optimize = true;
std::vector<double*> parameter_head;
std::vector<int> parameter_block_size;
}  // namespace internal

namespace swgn {
namespace {
std::unordered_map<std::type_index, Adapter>& registry() {
  static std::unordered_map<std::type_index, Adapter> r;
  return r;
}
Globals g_globals;
}  // namespace
void RegisterAdapter(const std::type_index& type, Adapter adapter) { registry()[type] = adapter; }
void SetGlobals(const Globals& g) { g_globals = g; }
const Globals& GetGlobals() { return g_globals; }
}  // namespace swgn

// ---------------------------------------------------------------------------------------------
// Problem
// ---------------------------------------------------------------------------------------------
void Problem::Fatal(const char* what) const {
  std::fprintf(stderr, "ceres (swgn shim) CHECK failed: %s\n", what);
  std::abort();
}
void Problem::Release(const CostFunction* c) {
  if (options_.cost_function_ownership != TAKE_OWNERSHIP || !c) return;
  if (--cost_refs_[c] == 0) {
    cost_refs_.erase(c);
    delete c;
  }
}
void Problem::Release(const LossFunction* l) {
  if (options_.loss_function_ownership != TAKE_OWNERSHIP || !l) return;
  if (--loss_refs_[l] == 0) {
    loss_refs_.erase(l);
    delete l;
  }
}
Problem::~Problem() {
  for (internal::ResidualBlock* rb : residual_blocks_) {
    Release(rb->cost_function());
    Release(rb->loss_function());
    delete rb;
  }
  if (options_.local_parameterization_ownership == TAKE_OWNERSHIP)
    for (auto& kv : param_refs_) delete static_cast<const LocalParameterization*>(kv.first);
}
void Problem::AddParameterBlock(double* values, int size) { AddParameterBlock(values, size, nullptr); }
void Problem::AddParameterBlock(double* values, int size, LocalParameterization* lp) {
  if (!values || size <= 0) Fatal("AddParameterBlock: null block or non-positive size");
  auto it = blocks_.find(values);
  if (it != blocks_.end()) {
    if (it->second.size != size) Fatal("AddParameterBlock: block re-added with a different size");
    if (lp) SetParameterization(values, lp);
    return;
  }
  ParameterBlockInfo info;
  info.size = size;
  info.index = next_block_index_++;
  blocks_[values] = info;
  if (lp) SetParameterization(values, lp);
}
void Problem::SetParameterization(double* values, LocalParameterization* lp) {
  auto it = blocks_.find(values);
  if (it == blocks_.end()) Fatal("SetParameterization: unknown parameter block");
  if (lp && lp->GlobalSize() != it->second.size) Fatal("SetParameterization: size mismatch");
  it->second.parameterization = lp;
  if (lp) param_refs_[lp] = 1;
}
const LocalParameterization* Problem::GetParameterization(const double* values) const {
  auto it = blocks_.find(const_cast<double*>(values));
  return it == blocks_.end() ? nullptr : it->second.parameterization;
}
ResidualBlockId Problem::AddResidualBlock(CostFunction* cost, LossFunction* loss, const std::vector<double*>& params) {
  if (!cost) Fatal("AddResidualBlock: null cost function");
  if (params.size() != cost->parameter_block_sizes().size()) Fatal("AddResidualBlock: wrong number of parameter blocks");
  for (size_t i = 0; i < params.size(); ++i) {
    for (size_t j = i + 1; j < params.size(); ++j)
      if (params[i] == params[j]) Fatal("AddResidualBlock: duplicate parameter block in a residual block");
    AddParameterBlock(params[i], cost->parameter_block_sizes()[i]);
  }
  internal::ResidualBlock* rb = new internal::ResidualBlock(cost, loss, params, (int)residual_blocks_.size());
  residual_blocks_.push_back(rb);
  for (double* p : params) blocks_[p].residual_blocks.insert(rb);
  if (options_.cost_function_ownership == TAKE_OWNERSHIP) ++cost_refs_[cost];
  if (options_.loss_function_ownership == TAKE_OWNERSHIP && loss) ++loss_refs_[loss];
  return rb;
}
void Problem::RemoveResidualBlock(ResidualBlockId rb) {
  if (!rb) Fatal("RemoveResidualBlock: null");
  const int i = rb->index();
  if (i < 0 || i >= (int)residual_blocks_.size() || residual_blocks_[i] != rb) Fatal("RemoveResidualBlock: unknown residual block");
  for (double* p : rb->parameter_blocks()) {
    auto it = blocks_.find(p);
    if (it != blocks_.end()) it->second.residual_blocks.erase(rb);
  }
  // problem_impl.cc DeleteBlockInVector: the last block takes the freed slot
  residual_blocks_[i] = residual_blocks_.back();
  residual_blocks_[i]->set_index(i);
  residual_blocks_.pop_back();
  Release(rb->cost_function());
  Release(rb->loss_function());
  delete rb;
}
void Problem::RemoveParameterBlock(const double* values) {
  auto it = blocks_.find(const_cast<double*>(values));
  if (it == blocks_.end()) Fatal("RemoveParameterBlock: unknown parameter block");
  std::vector<internal::ResidualBlock*> dependents(it->second.residual_blocks.begin(), it->second.residual_blocks.end());
  std::sort(dependents.begin(), dependents.end(), [](auto* a, auto* b) { return a->index() > b->index(); });
  for (internal::ResidualBlock* rb : dependents) RemoveResidualBlock(rb);
  blocks_.erase(const_cast<double*>(values));
}
void Problem::SetParameterBlockConstant(const double* values) {
  auto it = blocks_.find(const_cast<double*>(values));
  if (it == blocks_.end()) Fatal("SetParameterBlockConstant: unknown parameter block");
  it->second.constant = true;
}
void Problem::SetParameterBlockVariable(double* values) {
  auto it = blocks_.find(values);
  if (it == blocks_.end()) Fatal("SetParameterBlockVariable: unknown parameter block");
  it->second.constant = false;
}
bool Problem::IsParameterBlockConstant(const double* values) const {
  auto it = blocks_.find(const_cast<double*>(values));
  if (it == blocks_.end()) Fatal("IsParameterBlockConstant: unknown parameter block");
  return it->second.constant;
}
int Problem::ParameterBlockSize(const double* values) const {
  auto it = blocks_.find(const_cast<double*>(values));
  if (it == blocks_.end()) Fatal("ParameterBlockSize: unknown parameter block");
  return it->second.size;
}
int Problem::ParameterBlockLocalSize(const double* values) const {
  auto it = blocks_.find(const_cast<double*>(values));
  if (it == blocks_.end()) Fatal("ParameterBlockLocalSize: unknown parameter block");
  return it->second.parameterization ? it->second.parameterization->LocalSize() : it->second.size;
}
int Problem::NumParameters() const {
  int n = 0;
  for (auto& kv : blocks_) n += kv.second.size;
  return n;
}
int Problem::NumResiduals() const {
  int n = 0;
  for (auto* rb : residual_blocks_) n += rb->cost_function()->num_residuals();
  return n;
}
void Problem::GetParameterBlocks(std::vector<double*>* out) const {
  std::vector<std::pair<int, double*>> v;
  for (auto& kv : blocks_) v.push_back({kv.second.index, kv.first});
  std::sort(v.begin(), v.end());
  out->clear();
  for (auto& p : v) out->push_back(p.second);
}
void Problem::GetResidualBlocks(std::vector<ResidualBlockId>* out) const { *out = residual_blocks_; }
void Problem::GetParameterBlocksForResidualBlock(const ResidualBlockId rb, std::vector<double*>* out) const { *out = rb->parameter_blocks(); }
void Problem::GetResidualBlocksForParameterBlock(const double* values, std::vector<ResidualBlockId>* out) const {
  auto it = blocks_.find(const_cast<double*>(values));
  if (it == blocks_.end()) Fatal("GetResidualBlocksForParameterBlock: unknown parameter block");
  out->assign(it->second.residual_blocks.begin(), it->second.residual_blocks.end());
  std::sort(out->begin(), out->end(), [](auto* a, auto* b) { return a->index() < b->index(); });
}

// ---------------------------------------------------------------------------------------------
// Solve
// ---------------------------------------------------------------------------------------------
std::string Solver::Summary::BriefReport() const {
  char buf[256];
  static const char* names[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE", "USER_SUCCESS", "USER_FAILURE"};
  std::snprintf(buf, sizeof(buf), "Ceres(swgn) Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
                num_successful_steps + num_unsuccessful_steps, initial_cost, final_cost, names[termination_type]);
  return buf;
}

namespace {
// The device knows one non-Euclidean manifold: the reference's 7 -> 6 pose parameterization
// (RVI/factor/pose_local_parameterization.cpp:5-27).  Recognise it by behaviour, not by name:
// probe Plus() on a fixed vector and compare with p + dp, normalize(q * [1, dtheta/2]).
bool is_pose_parameterization(const LocalParameterization* lp) {
  if (!lp || lp->GlobalSize() != 7 || lp->LocalSize() != 6) return false;
  const double x[7] = {1.0, -2.0, 0.5, 0.1, -0.2, 0.3, std::sqrt(1.0 - 0.14)};
  const double d[6] = {0.01, 0.02, -0.03, 0.004, -0.005, 0.006};
  double out[7];
  if (!lp->Plus(x, d, out)) return false;
  const double qw = x[6], qx = x[3], qy = x[4], qz = x[5], dw = 1.0, dx = d[3] / 2, dy = d[4] / 2, dz = d[5] / 2;
  double r[4] = {qw * dw - qx * dx - qy * dy - qz * dz, qw * dx + qx * dw + qy * dz - qz * dy,
                 qw * dy + qy * dw + qz * dx - qx * dz, qw * dz + qz * dw + qx * dy - qy * dx};
  const double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
  const double expect[7] = {x[0] + d[0], x[1] + d[1], x[2] + d[2], r[1] / n, r[2] / n, r[3] / n, r[0] / n};
  for (int i = 0; i < 7; ++i)
    if (std::fabs(out[i] - expect[i]) > 1e-12) return false;
  return true;
}
void fail(Solver::Summary* s, const std::string& m) {
  s->termination_type = FAILURE;
  s->message = m;
}
}  // namespace

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  *summary = Solver::Summary();
  // what the device path implements: DENSE_SCHUR with a user ordering, monotonic steps, and either TRADITIONAL DOGLEG with
  // jacobi_scaling = false (the sliding-window solve, RVI/swf/swf.cpp:25-30) or LEVENBERG_MARQUARDT with or without
  // jacobi_scaling (Ceres' defaults: the per-epoch GNSS solves, RVI/swf/swf_gnss.cpp:204-215,563-573)
  if (options.linear_solver_type != DENSE_SCHUR || options.minimizer_type != TRUST_REGION || options.use_nonmonotonic_steps)
    return fail(summary, "swgn shim: only TRUST_REGION + DENSE_SCHUR with monotonic steps is implemented");
  if (options.trust_region_strategy_type == DOGLEG && (options.jacobi_scaling || options.dogleg_type != TRADITIONAL_DOGLEG))
    return fail(summary, "swgn shim: DOGLEG is implemented as TRADITIONAL_DOGLEG with jacobi_scaling = false "
                         "(the reference's configuration, RVI/swf/swf.cpp:25-30)");
  if (options.trust_region_strategy_type != DOGLEG && options.trust_region_strategy_type != LEVENBERG_MARQUARDT)
    return fail(summary, "swgn shim: unknown trust-region strategy");

  // ---- parameter blocks in insertion order
  std::vector<double*> blocks;
  problem->GetParameterBlocks(&blocks);
  const auto& bmap = problem->parameter_block_map();
  // No ordering given (the reference's per-epoch GNSS solves, RVI/swf/swf_gnss.cpp:204-215,563-573): Ceres picks the
  // elimination group itself, a maximal independent set of the variable blocks found greedily in order of increasing
  // degree (ReorderProgramForSchurTypeLinearSolver -> ComputeStableSchurOrdering, reorder_program.cc / graph_algorithms.h:
  // StableIndependentSetOrdering).  Which independent set is eliminated does not change the step of an exact solve.
  std::shared_ptr<ParameterBlockOrdering> auto_ordering;
  if (!options.linear_solver_ordering) {
    std::unordered_map<const double*, std::vector<const double*>> nbr;
    std::vector<double*> variable;
    for (double* p : blocks)
      if (!bmap.at(p).constant) variable.push_back(p);
    for (internal::ResidualBlock* rb : problem->residual_block_list())
      for (double* a : rb->parameter_blocks())
        for (double* c : rb->parameter_blocks())
          if (a != c && !bmap.at(a).constant && !bmap.at(c).constant) nbr[a].push_back(c);
    std::stable_sort(variable.begin(), variable.end(), [&](const double* a, const double* c) {
      auto deg = [&](const double* v) {
        auto it = nbr.find(v);
        if (it == nbr.end()) return (size_t)0;
        std::vector<const double*> u(it->second);
        std::sort(u.begin(), u.end());
        return (size_t)(std::unique(u.begin(), u.end()) - u.begin());
      };
      return deg(a) < deg(c);
    });
    auto_ordering = std::make_shared<ParameterBlockOrdering>();
    std::unordered_map<const double*, int> colour;  // 0 untouched, 1 in the set, 2 neighbour of the set
    for (double* v : variable) {
      if (colour[v] != 0) continue;
      colour[v] = 1;
      auto it = nbr.find(v);
      if (it != nbr.end())
        for (const double* u : it->second) colour[u] = 2;
    }
    for (double* pb : blocks) auto_ordering->AddElementToGroup(pb, colour[pb] == 1 ? 0 : 1);
  }
  const ParameterBlockOrdering* ordering = options.linear_solver_ordering ? options.linear_solver_ordering.get() : auto_ordering.get();
  std::unordered_map<const double*, int> index_of;
  std::vector<int32_t> bsize, bman, bconst, bgroup, boff;
  std::vector<double> state;
  for (double* p : blocks) {
    const Problem::ParameterBlockInfo& info = bmap.at(p);
    index_of[p] = (int)bsize.size();
    bsize.push_back(info.size);
    if (info.parameterization) {
      if (!is_pose_parameterization(info.parameterization))
        return fail(summary, "swgn shim: unsupported LocalParameterization (only the 7->6 pose parameterization runs on the device)");
      bman.push_back(SWGN_MANIFOLD_POSE);
    } else {
      bman.push_back(SWGN_MANIFOLD_EUCLIDEAN);
    }
    bconst.push_back(info.constant ? 1 : 0);
    bgroup.push_back(ordering->GroupId(p));
    boff.push_back((int32_t)state.size());
    state.insert(state.end(), p, p + info.size);
  }
  // ---- a problem without a variable parameter block never reaches the device: like Ceres (solver.cc Minimize:
  // "No non-constant parameter blocks found", reduced program empty) it converges at once with the fixed cost, which is
  // evaluated -- as Ceres does in Program::RemoveFixedBlocks, program.cc:304-411 -- by the user's own cost functions
  {
    bool any_variable = false;
    for (internal::ResidualBlock* rb : problem->residual_block_list())
      for (double* p : rb->parameter_blocks())
        if (!bmap.at(p).constant) any_variable = true;
    if (!any_variable) {
      double fixed = 0.0;
      std::vector<double> r;
      for (internal::ResidualBlock* rb : problem->residual_block_list()) {
        const CostFunction* cf = rb->cost_function();
        r.assign(cf->num_residuals(), 0.0);
        if (!cf->Evaluate(rb->parameter_blocks().data(), r.data(), nullptr)) return fail(summary, "swgn shim: evaluation of a fixed residual block failed");
        double sq = 0.0;
        for (double v : r) sq += v * v;
        if (rb->loss_function()) {
          double rho[3];
          rb->loss_function()->Evaluate(sq, rho);
          sq = rho[0];
        }
        fixed += 0.5 * sq;
      }
      summary->termination_type = CONVERGENCE;
      summary->message = "Function tolerance reached. No non-constant parameter blocks found.";
      summary->initial_cost = summary->final_cost = summary->fixed_cost = fixed;
      summary->num_successful_steps = summary->num_unsuccessful_steps = summary->num_linear_solves = 0;
      summary->num_parameter_blocks = (int)blocks.size();
      summary->num_residual_blocks = (int)problem->residual_block_list().size();
      summary->num_parameter_blocks_reduced = summary->num_residuals_reduced = 0;
      return;
    }
  }
  // ---- residual blocks through the adapters
  std::vector<int32_t> proj_blocks, imu_blocks, gnss_kind, gnss_blocks, prior_n, prior_blk_begin{0}, prior_blocks, prior_blk_idx, unit_block;
  std::vector<int64_t> prior_x0_begin, prior_J_begin, prior_r_begin;
  std::vector<double> proj_uv, imu_data, gnss_data, prior_x0, prior_J, prior_r0, unit_istd;
  std::vector<uint32_t> order;
  std::vector<uint8_t> use_by_kind[swgn::kNumKinds];
  std::vector<int32_t> chain_blk_begin{0}, chain_blocks, chain_frame_begin{0};
  std::vector<double> chain_frames, chain_frame_N, chain_N, chain_imu;
  std::vector<double*> chain_pose_ptr, chain_sb_ptr;
  std::vector<const CostFunction*> host_cf;  // cost functions without a device adapter
  std::vector<int32_t> host_nres, host_blk_begin{0}, host_blocks;
  double cauchy_a = -1.0;
  bool any_masked = false;
  for (internal::ResidualBlock* rb : problem->residual_block_list()) {
    const CostFunction* cf = rb->cost_function();
    auto it = swgn::registry().find(std::type_index(typeid(*cf)));
    std::vector<int32_t> ids;
    for (double* p : rb->parameter_blocks()) ids.push_back(index_of.at(p));
    if (it == swgn::registry().end()) {
      // no device kind for this cost function: it is evaluated on the host by its own Evaluate() (cost_function.h:116)
      if (rb->loss_function()) return fail(summary, std::string("swgn shim: a loss function on the host-evaluated cost function ") + typeid(*cf).name() + " is not supported");
      if (ids.size() != cf->parameter_block_sizes().size()) return fail(summary, "swgn shim: parameter block count mismatch");
      order.push_back(((uint32_t)swgn::kHost << 28) | (uint32_t)host_cf.size());
      use_by_kind[swgn::kHost].push_back(rb->is_use ? 1 : 0);
      any_masked |= !rb->is_use;
      host_cf.push_back(cf);
      host_nres.push_back(cf->num_residuals());
      host_blocks.insert(host_blocks.end(), ids.begin(), ids.end());
      host_blk_begin.push_back((int32_t)host_blocks.size());
      continue;
    }
    swgn::FactorRecord rec;
    if (!it->second(cf, &rec)) return fail(summary, std::string("swgn shim: adapter failed for ") + typeid(*cf).name());
    const LossFunction* loss = rb->loss_function();
    if (loss) {
      const CauchyLoss* cl = dynamic_cast<const CauchyLoss*>(loss);
      if (!cl || rec.kind != swgn::kProjection) return fail(summary, "swgn shim: only CauchyLoss on projection factors runs on the device");
      if (cauchy_a > 0 && cauchy_a != cl->a()) return fail(summary, "swgn shim: projection factors must share one CauchyLoss scale");
      cauchy_a = cl->a();
    } else if (rec.kind == swgn::kProjection && cauchy_a > 0) {
      return fail(summary, "swgn shim: projection factors must all carry the same loss");
    }
    uint32_t idx = 0;
    switch (rec.kind) {
      case swgn::kProjection:
        idx = (uint32_t)(proj_uv.size() / 2);
        proj_blocks.insert(proj_blocks.end(), ids.begin(), ids.end());
        proj_uv.insert(proj_uv.end(), rec.data.begin(), rec.data.begin() + 2);
        break;
      case swgn::kImu:
        idx = (uint32_t)(imu_blocks.size() / 4);
        imu_blocks.insert(imu_blocks.end(), ids.begin(), ids.end());
        rec.data.resize(SWGN_IMU_STRIDE, 0.0);
        imu_data.insert(imu_data.end(), rec.data.begin(), rec.data.end());
        break;
      case swgn::kGnss:
        idx = (uint32_t)gnss_kind.size();
        gnss_kind.push_back(rec.gnss_kind);
        ids.resize(3, -1);
        gnss_blocks.insert(gnss_blocks.end(), ids.begin(), ids.end());
        rec.data.resize(SWGN_GNSS_STRIDE, 0.0);
        gnss_data.insert(gnss_data.end(), rec.data.begin(), rec.data.end());
        break;
      case swgn::kPrior:
        idx = (uint32_t)prior_n.size();
        prior_n.push_back(rec.prior_n);
        prior_blocks.insert(prior_blocks.end(), ids.begin(), ids.end());
        prior_blk_idx.insert(prior_blk_idx.end(), rec.prior_blk_idx.begin(), rec.prior_blk_idx.end());
        prior_blk_begin.push_back((int32_t)prior_blocks.size());
        prior_x0_begin.push_back((int64_t)prior_x0.size());
        prior_J_begin.push_back((int64_t)prior_J.size());
        prior_r_begin.push_back((int64_t)prior_r0.size());
        prior_x0.insert(prior_x0.end(), rec.prior_x0.begin(), rec.prior_x0.end());
        prior_J.insert(prior_J.end(), rec.prior_J.begin(), rec.prior_J.end());
        prior_r0.insert(prior_r0.end(), rec.prior_r0.begin(), rec.prior_r0.end());
        break;
      case swgn::kUnit:
        idx = (uint32_t)unit_block.size();
        unit_block.push_back(ids[0]);
        unit_istd.push_back(rec.data[0]);
        break;
      case swgn::kChain: {
        idx = (uint32_t)(chain_blk_begin.size() - 1);
        const int k = (int)ids.size() - 4, m = rec.chain_m;
        if (k < 0 || m < 1 || rec.chain_frames.size() != (size_t)m * SWGN_CHAIN_FRAME_STRIDE || rec.chain_frame_N.size() != (size_t)m * 15 * k ||
            rec.chain_N.size() != (size_t)k * k + k || rec.chain_imu.size() != (size_t)(m + 1) * SWGN_IMU_STRIDE ||
            rec.chain_pose_ptr.size() != (size_t)m || rec.chain_sb_ptr.size() != (size_t)m)
          return fail(summary, "swgn shim: malformed IMUGNSSFactor chain record");
        chain_blocks.insert(chain_blocks.end(), ids.begin(), ids.end());
        chain_blk_begin.push_back((int32_t)chain_blocks.size());
        chain_frame_begin.push_back(chain_frame_begin.back() + m);
        chain_frames.insert(chain_frames.end(), rec.chain_frames.begin(), rec.chain_frames.end());
        chain_frame_N.insert(chain_frame_N.end(), rec.chain_frame_N.begin(), rec.chain_frame_N.end());
        chain_N.insert(chain_N.end(), rec.chain_N.begin(), rec.chain_N.end());
        chain_imu.insert(chain_imu.end(), rec.chain_imu.begin(), rec.chain_imu.end());
        chain_pose_ptr.insert(chain_pose_ptr.end(), rec.chain_pose_ptr.begin(), rec.chain_pose_ptr.end());
        chain_sb_ptr.insert(chain_sb_ptr.end(), rec.chain_sb_ptr.begin(), rec.chain_sb_ptr.end());
        break;
      }
      default:
        return fail(summary, "swgn shim: adapter produced an unknown factor kind");
    }
    order.push_back(((uint32_t)rec.kind << 28) | idx);
    use_by_kind[rec.kind].push_back(rb->is_use ? 1 : 0);
    any_masked |= !rb->is_use;
  }
  std::vector<uint8_t> is_use;
  for (int k = 0; k < swgn::kNumKinds; ++k) is_use.insert(is_use.end(), use_by_kind[k].begin(), use_by_kind[k].end());

  swgn_graph g;
  std::memset(&g, 0, sizeof(g));
  g.n_blocks = (int32_t)bsize.size();
  g.block_size = bsize.data();
  g.block_manifold = bman.data();
  g.block_const = bconst.data();
  g.block_group = bgroup.data();
  g.block_offset = boff.data();
  g.n_state = (int32_t)state.size();
  g.state = state.data();
  const swgn::Globals& gl = swgn::GetGlobals();
  std::memcpy(g.Pbg, gl.Pbg, sizeof(g.Pbg));
  std::memcpy(g.gravity, gl.gravity, sizeof(g.gravity));
  std::memcpy(g.proj_sqrt_info, gl.proj_sqrt_info, sizeof(g.proj_sqrt_info));
  g.proj_cauchy_a = cauchy_a;
  g.n_proj = (int32_t)(proj_uv.size() / 2);
  g.proj_blocks = proj_blocks.data();
  g.proj_uv = proj_uv.data();
  g.n_imu = (int32_t)(imu_blocks.size() / 4);
  g.imu_blocks = imu_blocks.data();
  g.imu_data = imu_data.data();
  g.n_gnss = (int32_t)gnss_kind.size();
  g.gnss_kind = gnss_kind.data();
  g.gnss_blocks = gnss_blocks.data();
  g.gnss_data = gnss_data.data();
  g.n_prior = (int32_t)prior_n.size();
  g.prior_n = prior_n.data();
  g.prior_blk_begin = prior_blk_begin.data();
  g.prior_blocks = prior_blocks.data();
  g.prior_blk_idx = prior_blk_idx.data();
  g.prior_x0_begin = prior_x0_begin.data();
  g.prior_x0 = prior_x0.data();
  g.prior_J_begin = prior_J_begin.data();
  g.prior_J = prior_J.data();
  g.prior_r_begin = prior_r_begin.data();
  g.prior_r0 = prior_r0.data();
  g.n_unit = (int32_t)unit_block.size();
  g.unit_block = unit_block.data();
  g.unit_istd = unit_istd.data();
  g.n_order = (int32_t)order.size();
  g.order = order.data();
  g.is_use = any_masked ? is_use.data() : nullptr;
  g.n_host = (int32_t)host_cf.size();
  if (g.n_host > 0) {
    g.host_nres = host_nres.data();
    g.host_blk_begin = host_blk_begin.data();
    g.host_blocks = host_blocks.data();
    g.host_user = &host_cf;
    g.host_eval = [](void* user, int32_t factor, double const* const* parameters, double* residuals, double** jacobians) -> int32_t {
      const auto& cfs = *static_cast<const std::vector<const CostFunction*>*>(user);
      return cfs[factor]->Evaluate(parameters, residuals, jacobians) ? 0 : 1;
    };
  }
  g.n_chain = (int32_t)chain_blk_begin.size() - 1;
  if (g.n_chain > 0) {
    g.chain_blk_begin = chain_blk_begin.data();
    g.chain_blocks = chain_blocks.data();
    g.chain_frame_begin = chain_frame_begin.data();
    g.chain_frame_data = chain_frames.data();
    g.chain_frame_N = chain_frame_N.data();
    g.chain_N = chain_N.data();
    g.chain_imu_data = chain_imu.data();
  }

  swgn_options o;
  swgn_default_options(&o);
  o.max_num_iterations = options.max_num_iterations;
  o.max_num_consecutive_invalid_steps = options.max_num_consecutive_invalid_steps;
  o.initial_trust_region_radius = options.initial_trust_region_radius;
  o.max_trust_region_radius = options.max_trust_region_radius;
  o.min_trust_region_radius = options.min_trust_region_radius;
  o.min_relative_decrease = options.min_relative_decrease;
  o.min_lm_diagonal = options.min_lm_diagonal;
  o.max_lm_diagonal = options.max_lm_diagonal;
  o.function_tolerance = options.function_tolerance;
  o.gradient_tolerance = options.gradient_tolerance;
  o.parameter_tolerance = options.parameter_tolerance;
  o.trust_region_strategy = options.trust_region_strategy_type == DOGLEG ? SWGN_DOGLEG : SWGN_LEVENBERG_MARQUARDT;
  o.jacobi_scaling = options.jacobi_scaling ? 1 : 0;
  o.is_optimize = internal::is_optimize ? 1 : 0;
  o.n_parameter_head = (int32_t)internal::parameter_head.size();
  o.device = options.device;
  // the head blocks must be the last groups of the ordering (RVI/swf/swf_gnss.cpp:775-782)
  if (o.n_parameter_head > 0) {
    int max_other = -1, min_head = 1 << 30;
    std::unordered_map<const double*, bool> is_head;
    for (double* h : internal::parameter_head) is_head[h] = true;
    for (size_t i = 0; i < blocks.size(); ++i) {
      if (bconst[i]) continue;
      if (is_head.count(blocks[i])) min_head = std::min(min_head, bgroup[i]);
      else max_other = std::max(max_other, bgroup[i]);
    }
    if (min_head <= max_other) return fail(summary, "swgn shim: parameter_head blocks must occupy the last groups of the ordering");
  }

  const swgn_graph* gp = &g;
  swgn_batch* batch = nullptr;
  swgn_status st = swgn_batch_create(&o, 1, &gp, &batch);
  if (st != SWGN_OK) return fail(summary, std::string("swgn_batch_create: ") + swgn_last_error());
  swgn_summary sm;
  st = swgn_batch_solve(batch, &sm);
  if (st != SWGN_OK) {
    swgn_batch_destroy(batch);
    return fail(summary, std::string("swgn_batch_solve: ") + swgn_last_error());
  }
  // ---- results back into user memory (Program::CopyParameterBlockStateToUserState, program.cc:105)
  std::vector<double> out(state.size());
  st = swgn_batch_get_state(batch, 0, out.data());
  if (st == SWGN_OK)
    for (size_t i = 0; i < blocks.size(); ++i)
      if (!bconst[i]) std::memcpy(blocks[i], out.data() + boff[i], sizeof(double) * bsize[i]);
  // ---- hidden GNSS frames of the IMUGNSSFactor chains back into the arrays the factors point at
  // (IMUGNSSBase::UpdateHiddenState writes gnss_poses[i] / gnss_speed_bias[i], gnss_imu_factor.cpp:636-645)
  if (st == SWGN_OK && !chain_pose_ptr.empty()) {
    int32_t nfr = 0;
    std::vector<double> fr(16 * chain_pose_ptr.size());
    st = swgn_batch_get_chain_frames(batch, 0, &nfr, fr.data());
    if (st == SWGN_OK && nfr == (int32_t)chain_pose_ptr.size())
      for (int i = 0; i < nfr; ++i) {
        std::memcpy(chain_pose_ptr[i], fr.data() + 16 * i, sizeof(double) * 7);
        std::memcpy(chain_sb_ptr[i], fr.data() + 16 * i + 7, sizeof(double) * 9);
      }
  }
  // ---- side channel (M2)
  if (st == SWGN_OK && o.n_parameter_head > 0) {
    int32_t n = 0;
    if (!internal::is_optimize) {
      st = swgn_batch_get_reduced(batch, 0, nullptr, nullptr, &n);
      if (st == SWGN_OK && n >= RHSROWLIMIT) st = SWGN_ERR_TOO_LARGE;
      if (st == SWGN_OK) st = swgn_batch_get_reduced(batch, 0, internal::lhs_out, internal::rhs_out, &n);
    } else {
      st = swgn_batch_get_cholesky(batch, 0, nullptr, &n);
      if (st == SWGN_OK && n >= RHSROWLIMIT) st = SWGN_ERR_TOO_LARGE;
      if (st == SWGN_OK) st = swgn_batch_get_cholesky(batch, 0, internal::lhs_out2, &n);
    }
    if (st == SWGN_OK) internal::hs_row = n;
  }
  double total_ms = 0;
  swgn_batch_last_timing(batch, &total_ms, nullptr, nullptr, nullptr);
  swgn_batch_destroy(batch);
  if (st != SWGN_OK) return fail(summary, std::string("swgn read-back: ") + swgn_last_error());
  summary->termination_type = sm.termination_type == SWGN_CONVERGENCE ? CONVERGENCE : (sm.termination_type == SWGN_NO_CONVERGENCE ? NO_CONVERGENCE : FAILURE);
  summary->message = "swgn device solve";
  summary->initial_cost = sm.initial_cost;
  summary->final_cost = sm.final_cost;
  summary->fixed_cost = sm.fixed_cost;
  summary->num_successful_steps = sm.num_successful_steps;
  summary->num_unsuccessful_steps = sm.num_unsuccessful_steps;
  summary->num_linear_solves = sm.num_linear_solves;
  summary->minimizer_time_in_seconds = total_ms * 1e-3;
  summary->total_time_in_seconds = total_ms * 1e-3;
  summary->preprocessor_time_in_seconds = 0.0;
  summary->num_parameter_blocks = problem->NumParameterBlocks();
  summary->num_residual_blocks = problem->NumResidualBlocks();
  summary->num_residuals = problem->NumResiduals();
  summary->num_residuals_reduced = sm.n_residuals;
}
}  // namespace ceres
