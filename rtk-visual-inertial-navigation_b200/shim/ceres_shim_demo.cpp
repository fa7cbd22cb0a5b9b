// Self-test of the ceres:: shim: application-style code that builds a sliding-window problem
// through ceres::Problem exactly the way the reference does (AddParameterBlock with a pose
// parameterization, AddResidualBlock per factor with CauchyLoss on the projection factors,
// SetParameterBlockConstant on the extrinsic, a predefined multi-group ordering, the
// parameter_head / is_optimize side channel) and calls ceres::Solve.
//
// The factor classes below stand in for the reference's (same names, same public data members,
// RVI/factor/*.h) because the real ones need Eigen, which this image does not have; their
// Evaluate() is never called on the device path.  The window content comes from the synthetic
// generator (synth/swgn_synth.h).  Exposed to the tests as one C function.
#include <array>
#include <cstring>
#include <memory>
#include <vector>

#include "../synth/swgn_synth.h"
#include "ceres/ceres.h"
#include "ceres/schur_complement_solver.h"
#include "reference_adapters.h"

namespace {
struct Vec3 : std::array<double, 3> {};
struct QuatXYZW {
  double v[4];
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double w() const { return v[3]; }
};
struct Mat15 {
  double a[225];
  double operator()(int i, int j) const { return a[i * 15 + j]; }
};

class projection_factor : public ceres::SizedCostFunction<2, 7, 7, 3> {
 public:
  explicit projection_factor(const Vec3& _pts) : pts(_pts) {}
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  Vec3 pts;
};
struct IntegrationBase {
  Vec3 delta_p, delta_v, linearized_ba, linearized_bg, gyri, gyrj;
  QuatXYZW delta_q;
  double sum_dt;
  Mat15 jacobian, sqrt_info;
};
class IMUFactor : public ceres::SizedCostFunction<15, 7, 9, 7, 9> {
 public:
  explicit IMUFactor(IntegrationBase* p) : pre_integration(p) {}
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  IntegrationBase* pre_integration;
};
class RTKCarrierPhaseFactor : public ceres::SizedCostFunction<1, 7, 1, 1> {
 public:
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  double* satelite_pos;
  double L1_lam, lam, el, base_rover_time_diff, mea_var;
  double* base_pos;
  bool use_istd = true;
};
class RTKPseudorangeFactor : public ceres::SizedCostFunction<1, 7, 1> {
 public:
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  double* satelite_pos;
  double P1, el, base_rover_time_diff, mea_var;
  double* base_pos;
};
class SppDopplerFactor : public ceres::SizedCostFunction<1, 9, 1, 7> {
 public:
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  double* satelitev1;
  double* satelite_pos;
  double D1_lam, istd;
  double* base_pos;
};
class InitialBlackFactor : public ceres::SizedCostFunction<1, 1> {
 public:
  explicit InitialBlackFactor(double s) : istd(s) {}
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  double istd;
};
struct MarginalizationInfo {
  int m = 0, n = 0;
  std::vector<int> keep_block_size, keep_block_idx;
  std::vector<double*> keep_block_data;
  struct M {
    std::vector<double> a;
    int n;
    double operator()(int i, int j) const { return a[(size_t)i * n + j]; }
  } linearized_jacobians;
  struct V {
    std::vector<double> a;
    double operator()(int i) const { return a[i]; }
  } linearized_residuals;
};
class MarginalizationFactor : public ceres::CostFunction {
 public:
  explicit MarginalizationFactor(MarginalizationInfo* info) : marginalization_info(info) {
    for (int s : info->keep_block_size) mutable_parameter_block_sizes()->push_back(s);
    set_num_residuals(info->n);
  }
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  MarginalizationInfo* marginalization_info;
};
// RVI/factor/gnss_imu_factor.h:18-151 (the public members the device adapter reads)
struct DynMat {
  std::vector<double> a;
  int cols = 0;
  double operator()(int i, int j) const { return a[(size_t)i * cols + j]; }
  double operator()(int i) const { return a[i]; }
};
struct MarginalizationInfoStub;
struct IMUGNSSBase {
  std::vector<double*> gnss_phase_biases, gnss_speed_bias, gnss_speed_bias_lin, gnss_poses, gnss_poses_lin;
  DynMat phase_biases_hessians, phase_biases_rhs;
  std::vector<Mat15> pose_hessians;
  std::vector<DynMat> pose_phase_biases_hessians, pose_rhses;
  std::vector<IMUFactor*> imu_factors;
  IMUFactor* last_imu_factor = nullptr;
  void* gnss_middle_marginfo = nullptr;
  std::vector<std::unique_ptr<IMUFactor>> owned;
  std::vector<std::unique_ptr<IntegrationBase>> pre;
  std::vector<std::array<double, 16>> lin;
};
class IMUGNSSFactor : public ceres::CostFunction {
 public:
  explicit IMUGNSSFactor(IMUGNSSBase* info) : IMUGNSS_info(info) {
    for (int s : {7, 9, 7, 9}) mutable_parameter_block_sizes()->push_back(s);
    for (size_t i = 0; i < info->gnss_phase_biases.size(); ++i) mutable_parameter_block_sizes()->push_back(1);
    set_num_residuals(30 + (int)info->gnss_phase_biases.size());
  }
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  IMUGNSSBase* IMUGNSS_info;
};
void fill_preintegration(IntegrationBase& p, const double* r) {
  for (int k = 0; k < 3; ++k) {
    p.delta_p[k] = r[SWGN_IMU_DELTA_P + k]; p.delta_v[k] = r[SWGN_IMU_DELTA_V + k];
    p.linearized_ba[k] = r[SWGN_IMU_LIN_BA + k]; p.linearized_bg[k] = r[SWGN_IMU_LIN_BG + k];
    p.gyri[k] = r[SWGN_IMU_GYRI + k]; p.gyrj[k] = r[SWGN_IMU_GYRJ + k];
  }
  for (int k = 0; k < 4; ++k) p.delta_q.v[k] = r[SWGN_IMU_DELTA_Q + k];
  p.sum_dt = r[SWGN_IMU_SUM_DT];
  std::memcpy(p.jacobian.a, r + SWGN_IMU_JACOBIAN, sizeof(p.jacobian.a));
  std::memcpy(p.sqrt_info.a, r + SWGN_IMU_SQRT_INFO, sizeof(p.sqrt_info.a));
}
// RVI/factor/pose_local_parameterization.cpp:5-27
class PoseLocalParameterization : public ceres::LocalParameterization {
  bool Plus(const double* x, const double* d, double* o) const override {
    const double qw = x[6], qx = x[3], qy = x[4], qz = x[5], dx = d[3] / 2, dy = d[4] / 2, dz = d[5] / 2;
    double r[4] = {qw - qx * dx - qy * dy - qz * dz, qw * dx + qx + qy * dz - qz * dy, qw * dy + qy + qz * dx - qx * dz,
                   qw * dz + qz + qx * dy - qy * dx};
    const double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    for (int i = 0; i < 3; ++i) o[i] = x[i] + d[i];
    o[3] = r[1] / n; o[4] = r[2] / n; o[5] = r[3] / n; o[6] = r[0] / n;
    return true;
  }
  bool ComputeJacobian(const double*, double* j) const override {
    std::memset(j, 0, sizeof(double) * 42);
    for (int i = 0; i < 6; ++i) j[i * 6 + i] = 1.0;
    return true;
  }
  int GlobalSize() const override { return 7; }
  int LocalSize() const override { return 6; }
};

void register_adapters() {
  static bool done = false;
  if (done) return;
  done = true;
  using namespace swgn_adapters;
  ceres::swgn::RegisterAdapter(typeid(projection_factor), &projection<projection_factor>);
  ceres::swgn::RegisterAdapter(typeid(IMUFactor), &imu<IMUFactor>);
  ceres::swgn::RegisterAdapter(typeid(RTKCarrierPhaseFactor), &rtk_carrier_phase<RTKCarrierPhaseFactor>);
  ceres::swgn::RegisterAdapter(typeid(RTKPseudorangeFactor), &rtk_pseudorange<RTKPseudorangeFactor>);
  ceres::swgn::RegisterAdapter(typeid(SppDopplerFactor), &spp_doppler<SppDopplerFactor>);
  ceres::swgn::RegisterAdapter(typeid(InitialBlackFactor), &unit_prior<InitialBlackFactor>);
  ceres::swgn::RegisterAdapter(typeid(MarginalizationFactor), &marginalization<MarginalizationFactor>);
  ceres::swgn::RegisterAdapter(typeid(IMUGNSSFactor), &imu_gnss<IMUGNSSFactor>);
}
}  // namespace

// Builds the synthetic window `window_id` of config `which` through the ceres:: API, solves it and
// writes the user-memory state (graph layout) to state_out.  export_mode != 0 runs the reference's
// "export" solve (is_optimize = false, one iteration) and returns hs_row / lhs_out / rhs_out;
// otherwise lhs_out2 (Cholesky factor) is returned in mat_out when the window has a parameter head.
// Returns the termination type, or -1 when Solve reports an unsupported / failed call.
extern "C" int swgn_ceres_demo_solve(int which, uint64_t window_id, int export_mode, int device, double* state_out,
                                     double* cost_out /* initial, final */, int* steps_out /* successful, unsuccessful */,
                                     int* hs_row_out, double* mat_out, double* rhs_out, char* message, int message_len) {
  register_adapters();
  swgn_synth_config cfg;
  swgn_synth_default_config(which, &cfg);
  swgn_synth* S = swgn_synth_create(&cfg, window_id);
  if (!S) return -1;
  const swgn_graph* g = swgn_synth_graph(S);
  swgn_options so;
  swgn_synth_options(S, &so);
  // ---- "application" state: one heap array per parameter block, like para_pose[i] etc.
  std::vector<std::unique_ptr<double[]>> mem(g->n_blocks);
  for (int b = 0; b < g->n_blocks; ++b) {
    mem[b].reset(new double[g->block_size[b]]);
    std::memcpy(mem[b].get(), g->state + g->block_offset[b], sizeof(double) * g->block_size[b]);
  }
  ceres::swgn::Globals gl;
  std::memcpy(gl.Pbg, g->Pbg, sizeof(gl.Pbg));
  std::memcpy(gl.gravity, g->gravity, sizeof(gl.gravity));
  std::memcpy(gl.proj_sqrt_info, g->proj_sqrt_info, sizeof(gl.proj_sqrt_info));
  ceres::swgn::SetGlobals(gl);

  std::vector<std::unique_ptr<IntegrationBase>> pre(g->n_imu);
  std::vector<std::unique_ptr<MarginalizationInfo>> marg(g->n_prior);
  std::vector<std::array<double, 9>> gnss_store(g->n_gnss);  // sat pos, sat vel, base pos
  // the hidden GNSS frames of the IMUGNSSFactor chains live in application arrays that are NOT part of
  // the problem (the reference removes them: gnss_imu_factor.cpp:110-113)
  std::vector<std::unique_ptr<IMUGNSSBase>> chains(g->n_chain);
  std::vector<int32_t> hid_pose(g->n_chain ? g->chain_frame_begin[g->n_chain] : 0), hid_sb(hid_pose.size());
  swgn_synth_chain_frame_blocks(S, hid_pose.data(), hid_sb.data());
  std::vector<char> hidden(g->n_blocks, 0);
  for (size_t i = 0; i < hid_pose.size(); ++i) hidden[hid_pose[i]] = hidden[hid_sb[i]] = 1;
  int result = -1;
  {
    ceres::Problem problem;
    for (int b = 0; b < g->n_blocks; ++b) {
      if (hidden[b]) continue;
      if (g->block_manifold[b] == SWGN_MANIFOLD_POSE) problem.AddParameterBlock(mem[b].get(), 7, new PoseLocalParameterization());
      else problem.AddParameterBlock(mem[b].get(), g->block_size[b]);
    }
    for (int i = 0; i < g->n_proj; ++i) {
      Vec3 pts;
      pts[0] = g->proj_uv[2 * i]; pts[1] = g->proj_uv[2 * i + 1]; pts[2] = 1.0;
      problem.AddResidualBlock(new projection_factor(pts), new ceres::CauchyLoss(g->proj_cauchy_a), mem[g->proj_blocks[3 * i]].get(),
                               mem[g->proj_blocks[3 * i + 1]].get(), mem[g->proj_blocks[3 * i + 2]].get());
    }
    for (int i = 0; i < g->n_imu; ++i) {
      const double* r = g->imu_data + (size_t)SWGN_IMU_STRIDE * i;
      pre[i].reset(new IntegrationBase());
      IntegrationBase& p = *pre[i];
      fill_preintegration(p, r);
      problem.AddResidualBlock(new IMUFactor(&p), nullptr, mem[g->imu_blocks[4 * i]].get(), mem[g->imu_blocks[4 * i + 1]].get(),
                               mem[g->imu_blocks[4 * i + 2]].get(), mem[g->imu_blocks[4 * i + 3]].get());
    }
    for (int i = 0; i < g->n_gnss; ++i) {
      const double* r = g->gnss_data + (size_t)SWGN_GNSS_STRIDE * i;
      std::memcpy(gnss_store[i].data(), r, sizeof(double) * 9);
      double* sat = gnss_store[i].data();
      double* vel = sat + 3;
      double* base = sat + 6;
      const int32_t* bl = g->gnss_blocks + 3 * i;
      if (g->gnss_kind[i] == SWGN_GNSS_RTK_CARRIER) {
        auto* f = new RTKCarrierPhaseFactor();
        f->satelite_pos = sat; f->base_pos = base; f->L1_lam = r[SWGN_GNSS_MEAS]; f->lam = r[SWGN_GNSS_LAM];
        f->el = r[SWGN_GNSS_EL]; f->base_rover_time_diff = r[SWGN_GNSS_DT]; f->mea_var = r[SWGN_GNSS_VAR];
        problem.AddResidualBlock(f, nullptr, mem[bl[0]].get(), mem[bl[1]].get(), mem[bl[2]].get());
      } else if (g->gnss_kind[i] == SWGN_GNSS_RTK_PSEUDORANGE) {
        auto* f = new RTKPseudorangeFactor();
        f->satelite_pos = sat; f->base_pos = base; f->P1 = r[SWGN_GNSS_MEAS];
        f->el = r[SWGN_GNSS_EL]; f->base_rover_time_diff = r[SWGN_GNSS_DT]; f->mea_var = r[SWGN_GNSS_VAR];
        problem.AddResidualBlock(f, nullptr, mem[bl[0]].get(), mem[bl[1]].get());
      } else if (g->gnss_kind[i] == SWGN_GNSS_DOPPLER) {
        auto* f = new SppDopplerFactor();
        f->satelite_pos = sat; f->satelitev1 = vel; f->base_pos = base; f->D1_lam = r[SWGN_GNSS_MEAS]; f->istd = r[SWGN_GNSS_WEIGHT];
        problem.AddResidualBlock(f, nullptr, mem[bl[0]].get(), mem[bl[1]].get(), mem[bl[2]].get());
      }
    }
    for (int i = 0; i < g->n_prior; ++i) {
      marg[i].reset(new MarginalizationInfo());
      MarginalizationInfo& m = *marg[i];
      m.n = g->prior_n[i];
      m.m = 0;
      std::vector<double*> params;
      const double* x0 = g->prior_x0 + g->prior_x0_begin[i];
      for (int k = g->prior_blk_begin[i]; k < g->prior_blk_begin[i + 1]; ++k) {
        const int b = g->prior_blocks[k];
        m.keep_block_size.push_back(g->block_size[b]);
        m.keep_block_idx.push_back(g->prior_blk_idx[k]);
        m.keep_block_data.push_back(const_cast<double*>(x0));
        x0 += g->block_size[b];
        params.push_back(mem[b].get());
      }
      m.linearized_jacobians.n = m.n;
      m.linearized_jacobians.a.assign(g->prior_J + g->prior_J_begin[i], g->prior_J + g->prior_J_begin[i] + (size_t)m.n * m.n);
      m.linearized_residuals.a.assign(g->prior_r0 + g->prior_r_begin[i], g->prior_r0 + g->prior_r_begin[i] + m.n);
      problem.AddResidualBlock(new MarginalizationFactor(&m), nullptr, params);
    }
    for (int i = 0; i < g->n_unit; ++i) problem.AddResidualBlock(new InitialBlackFactor(g->unit_istd[i]), nullptr, mem[g->unit_block[i]].get());
    // chains last: the default program order of the flat graph is proj, imu, gnss, prior, unit, chain
    {
      size_t fn_off = 0, cn_off = 0, imu_off = 0;
      for (int c = 0; c < g->n_chain; ++c) {
        const int b0 = g->chain_blk_begin[c], k = g->chain_blk_begin[c + 1] - b0 - 4;
        const int f0 = g->chain_frame_begin[c], m = g->chain_frame_begin[c + 1] - f0;
        chains[c].reset(new IMUGNSSBase());
        IMUGNSSBase& B = *chains[c];
        B.lin.resize(m);
        for (int q = 0; q < k; ++q) B.gnss_phase_biases.push_back(mem[g->chain_blocks[b0 + 4 + q]].get());
        B.phase_biases_hessians.cols = k;
        B.phase_biases_hessians.a.assign(g->chain_N + cn_off, g->chain_N + cn_off + (size_t)k * k);
        B.phase_biases_rhs.a.assign(g->chain_N + cn_off + (size_t)k * k, g->chain_N + cn_off + (size_t)k * k + k);
        for (int i = 0; i <= m; ++i) {
          B.pre.emplace_back(new IntegrationBase());
          fill_preintegration(*B.pre.back(), g->chain_imu_data + imu_off + (size_t)SWGN_IMU_STRIDE * i);
          B.owned.emplace_back(new IMUFactor(B.pre.back().get()));
          if (i < m) B.imu_factors.push_back(B.owned.back().get());
          else B.last_imu_factor = B.owned.back().get();
        }
        for (int i = 0; i < m; ++i) {
          const double* fr = g->chain_frame_data + (size_t)SWGN_CHAIN_FRAME_STRIDE * (f0 + i);
          // current hidden state in application memory, linearisation point kept by the factor
          std::memcpy(mem[hid_pose[f0 + i]].get(), fr + SWGN_CHAIN_POSE, sizeof(double) * 7);
          std::memcpy(mem[hid_sb[f0 + i]].get(), fr + SWGN_CHAIN_SB, sizeof(double) * 9);
          std::memcpy(B.lin[i].data(), fr + SWGN_CHAIN_POSE_LIN, sizeof(double) * 16);
          B.gnss_poses.push_back(mem[hid_pose[f0 + i]].get());
          B.gnss_speed_bias.push_back(mem[hid_sb[f0 + i]].get());
          B.gnss_poses_lin.push_back(B.lin[i].data());
          B.gnss_speed_bias_lin.push_back(B.lin[i].data() + 7);
          Mat15 ph;
          std::memcpy(ph.a, fr + SWGN_CHAIN_HESSIAN, sizeof(ph.a));
          B.pose_hessians.push_back(ph);
          DynMat pr, pn;
          pr.a.assign(fr + SWGN_CHAIN_RHS, fr + SWGN_CHAIN_RHS + 15);
          pn.cols = k;
          pn.a.assign(g->chain_frame_N + fn_off + (size_t)15 * k * i, g->chain_frame_N + fn_off + (size_t)15 * k * (i + 1));
          B.pose_rhses.push_back(pr);
          B.pose_phase_biases_hessians.push_back(pn);
        }
        std::vector<double*> params;
        for (int q = b0; q < g->chain_blk_begin[c + 1]; ++q) params.push_back(mem[g->chain_blocks[q]].get());
        problem.AddResidualBlock(new IMUGNSSFactor(&B), nullptr, params);
        fn_off += (size_t)m * 15 * k;
        cn_off += (size_t)k * k + k;
        imu_off += (size_t)(m + 1) * SWGN_IMU_STRIDE;
      }
    }
    for (int b = 0; b < g->n_blocks; ++b)
      if (g->block_const[b] && !hidden[b]) problem.SetParameterBlockConstant(mem[b].get());

    ceres::Solver::Options options;
    options.linear_solver_type = ceres::DENSE_SCHUR;
    options.trust_region_strategy_type = ceres::DOGLEG;
    options.max_num_iterations = export_mode ? 1 : so.max_num_iterations;
    options.num_threads = 4;
    options.jacobi_scaling = 0;
    options.device = device;
    options.linear_solver_ordering = std::make_shared<ceres::ParameterBlockOrdering>();
    for (int b = 0; b < g->n_blocks; ++b)
      if (g->block_group[b] >= 0 && !hidden[b]) options.linear_solver_ordering->AddElementToGroup(mem[b].get(), g->block_group[b]);
    // the ambiguities to be resolved go last and are announced through the side channel
    ceres::internal::parameter_head.clear();
    int32_t info[8];
    swgn_synth_info(S, info);
    for (int k = 0; k < so.n_parameter_head; ++k) ceres::internal::parameter_head.push_back(mem[info[5] + k].get());
    ceres::internal::is_optimize = !export_mode;
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    ceres::internal::is_optimize = true;
    if (message && message_len > 0) std::snprintf(message, message_len, "%s | %s", summary.message.c_str(), summary.BriefReport().c_str());
    if (summary.termination_type != ceres::FAILURE || summary.num_successful_steps >= 0) {
      result = summary.termination_type;
      if (cost_out) { cost_out[0] = summary.initial_cost; cost_out[1] = summary.final_cost; }
      if (steps_out) { steps_out[0] = summary.num_successful_steps; steps_out[1] = summary.num_unsuccessful_steps; }
    }
    if (hs_row_out) *hs_row_out = so.n_parameter_head > 0 ? ceres::internal::hs_row : 0;
    if (so.n_parameter_head > 0 && result >= 0) {
      const int n = ceres::internal::hs_row;
      if (mat_out) std::memcpy(mat_out, export_mode ? ceres::internal::lhs_out : ceres::internal::lhs_out2, sizeof(double) * n * n);
      if (rhs_out && export_mode) std::memcpy(rhs_out, ceres::internal::rhs_out, sizeof(double) * n);
    }
    ceres::internal::parameter_head.clear();
  }  // problem destroyed here (owns the factors)
  if (state_out)
    for (int b = 0; b < g->n_blocks; ++b) std::memcpy(state_out + g->block_offset[b], mem[b].get(), sizeof(double) * g->block_size[b]);
  swgn_synth_destroy(S);
  return result;
}
