// Drop-in evidence for the ceres:: shim: a sliding-window problem built from the REFERENCE'S OWN factor classes --
// RVI/factor/{projection_factor, imu_factor + integration_base, gnss_factor, marginalization_factor,
// pose_local_parameterization}.cpp, compiled
// unmodified where they lie under /root/reference (oracle/build_ref.sh; Eigen is not installed, so against the minimal
// stand-in of oracle/ref_stubs/) -- registered with the device adapters of shim/reference_adapters.h and solved by
// ceres::Solve of the shim, i.e. on the GPU.  The costs the device reports are then checked against the costs the
// reference's own Evaluate() methods give on the CPU at the same states.  TEST INFRASTRUCTURE: it links reference
// code, is built only where /root/reference exists and lives in oracle/_ref/.
// The window's prior is the reference's own MarginalizationFactor over a MarginalizationInfo filled through its public
// members (marginalization_factor.h:40-103), evaluated on the CPU by its own Evaluate() (marginalization_factor.cpp:410-446).
// With host_factors != 0 the window also gets the reference's INITIALISATION factors -- InitialPoseFactor,
// InitialBiasFactor, InitialFactor11 (RVI/factor/initial_factor.cpp) and InitPose0Factor (pose0_factor.cpp) -- for
// which no device adapter exists: the shim evaluates them on the host through their own Evaluate() (generic
// CostFunction contract, CERES/include/ceres/cost_function.h:116).
#include <array>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "../synth/swgn_synth.h"
#include "ceres/ceres.h"
#include "ceres/schur_complement_solver.h"
#include "factor/gnss_factor.h"
#include "factor/gnss_imu_factor.h"
#include "factor/imu_factor.h"
#include "factor/initial_factor.h"
#include "factor/marginalization_factor.h"
#include "factor/pose0_factor.h"
#include "factor/pose_local_parameterization.h"
#include "factor/projection_factor.h"
#include "reference_adapters.h"

extern Eigen::Vector3d Pbg;  // defined next to the factor trampolines (oracle/ref_shim.cpp)
extern Eigen::Matrix3d Rwgw;
extern Eigen::Vector3d G;

namespace {
struct Block {
  ceres::CostFunction* f;
  bool cauchy;
  std::vector<double*> params;
};
// 1/2 sum rho(|r|^2) with the reference's own Evaluate() (Cauchy: rho = a^2 log(1 + s / a^2), loss_function.cc:73-80);
// blocks whose parameters are all constant are the fixed cost
double cpu_cost(const std::vector<Block>& blocks, double a) {
  double total = 0.0;
  std::vector<double> r;
  for (const Block& b : blocks) {
    r.assign(b.f->num_residuals(), 0.0);
    b.f->Evaluate(b.params.data(), r.data(), nullptr);
    double s = 0.0;
    for (double v : r) s += v * v;
    total += 0.5 * (b.cauchy && a > 0 ? a * a * std::log1p(s / (a * a)) : s);
  }
  return total;
}

void register_adapters() {
  static bool done = false;
  if (done) return;
  done = true;
  using namespace swgn_adapters;
  ceres::swgn::RegisterAdapter(typeid(projection_factor), &projection<projection_factor>);
  ceres::swgn::RegisterAdapter(typeid(IMUFactor), &imu<IMUFactor>);
  ceres::swgn::RegisterAdapter(typeid(RTKCarrierPhaseFactor), &rtk_carrier_phase<RTKCarrierPhaseFactor>);
  ceres::swgn::RegisterAdapter(typeid(RTKPseudorangeFactor), &rtk_pseudorange<RTKPseudorangeFactor>);
  ceres::swgn::RegisterAdapter(typeid(SppPseudorangeFactor), &spp_pseudorange<SppPseudorangeFactor>);
  ceres::swgn::RegisterAdapter(typeid(SppCarrierPhaseFactor), &spp_carrier_phase<SppCarrierPhaseFactor>);
  ceres::swgn::RegisterAdapter(typeid(SppDopplerFactor), &spp_doppler<SppDopplerFactor>);
  ceres::swgn::RegisterAdapter(typeid(FixedIntegerFactor), &fixed_integer<FixedIntegerFactor>);
  ceres::swgn::RegisterAdapter(typeid(InitialBlackFactor), &unit_prior<InitialBlackFactor>);
  ceres::swgn::RegisterAdapter(typeid(MarginalizationFactor), &marginalization<MarginalizationFactor>);
  ceres::swgn::RegisterAdapter(typeid(IMUGNSSFactor), &imu_gnss<IMUGNSSFactor>);
}
// test hooks: run the solve in export mode (ceres::internal::is_optimize = false) and / or look at the problem and the
// shim's exported arrays right after ceres::Solve, before parameter_head is cleared -- where the reference calls UpdateSchur /
// UpdateSchurHessianOnly (RVI/swf/swf_image.cpp:232-236, swf_core.cpp:445-460)
int g_refdemo_is_optimize = 1;
void (*g_refdemo_after_solve)(ceres::Problem*) = nullptr;
double* (*g_refdemo_block_memory)(int block, int size) = nullptr;
void (*g_refdemo_before_solve)(ceres::Problem*, ceres::Solver::Options*, const swgn_graph*, double* const*) = nullptr;
// hidden GNSS-frame states of the last refdemo solve (16 doubles per frame, graph order), as the shim wrote them back
// into the arrays IMUGNSSBase::gnss_poses / gnss_speed_bias point at
std::vector<double> g_last_chain_frames;
IntegrationBase* integration_from_record(const double* r) {
  IntegrationBase* ib = new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero());
  for (int k = 0; k < 3; ++k) {
    ib->delta_p(k) = r[SWGN_IMU_DELTA_P + k];
    ib->delta_v(k) = r[SWGN_IMU_DELTA_V + k];
    ib->linearized_ba(k) = r[SWGN_IMU_LIN_BA + k];
    ib->linearized_bg(k) = r[SWGN_IMU_LIN_BG + k];
    ib->gyri(k) = r[SWGN_IMU_GYRI + k];
    ib->gyrj(k) = r[SWGN_IMU_GYRJ + k];
  }
  ib->delta_q = Eigen::Quaterniond(r[SWGN_IMU_DELTA_Q + 3], r[SWGN_IMU_DELTA_Q], r[SWGN_IMU_DELTA_Q + 1], r[SWGN_IMU_DELTA_Q + 2]);
  ib->sum_dt = r[SWGN_IMU_SUM_DT];
  for (int a = 0; a < 15; ++a)
    for (int c = 0; c < 15; ++c) {
      ib->jacobian(a, c) = r[SWGN_IMU_JACOBIAN + a * 15 + c];
      ib->sqrt_info(a, c) = r[SWGN_IMU_SQRT_INFO + a * 15 + c];
    }
  ib->covariance_update = false;
  return ib;
}
}  // namespace

// Builds the composition-B synthetic window through the ceres:: API out of reference factor objects, solves it on the
// device and reports cost_out = {device initial, device final, reference-CPU cost at the initial state, reference-CPU
// cost at the returned state}.  strategy: 0 DOGLEG (jacobi_scaling false), 1 LEVENBERG_MARQUARDT with Ceres' default
// jacobi_scaling = true.  Returns the termination type or -1.
extern "C" int swgn_ceres_refdemo_solve(int which, uint64_t window_id, int variant, int strategy, int host_factors, int device,
                                        double* state_out, double* cost_out, int* steps_out, char* message, int message_len) {
  register_adapters();
  swgn_synth_config cfg;
  swgn_synth_default_config(which, &cfg);
  cfg.variant = variant;
  swgn_synth* S = swgn_synth_create(&cfg, window_id);
  if (!S) return -1;
  const swgn_graph* g = swgn_synth_graph(S);
  swgn_options so;
  swgn_synth_options(S, &so);
  // parameter-block memory: this function's own, or the caller's (test hook: the estimator's para_pose / para_speed_bias /
  // feature / ambiguity storage, so that the reference's MyOrdering can recognise the blocks by address)
  std::vector<std::unique_ptr<double[]>> mem(g->n_blocks);
  std::vector<double*> ptr(g->n_blocks);
  for (int b = 0; b < g->n_blocks; ++b) {
    if (g_refdemo_block_memory) {
      ptr[b] = g_refdemo_block_memory(b, g->block_size[b]);
    } else {
      mem[b].reset(new double[g->block_size[b]]);
      ptr[b] = mem[b].get();
    }
    std::memcpy(ptr[b], g->state + g->block_offset[b], sizeof(double) * g->block_size[b]);
  }
  // the application globals, for the reference's Evaluate() on the CPU and for the device
  Pbg = Eigen::Vector3d(g->Pbg[0], g->Pbg[1], g->Pbg[2]);
  Rwgw = Eigen::Matrix3d::Identity();
  G = Eigen::Vector3d(g->gravity[0], g->gravity[1], g->gravity[2]);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) projection_factor::sqrt_info(i, j) = g->proj_sqrt_info[2 * i + j];
  ceres::swgn::Globals gl;
  std::memcpy(gl.Pbg, g->Pbg, sizeof(gl.Pbg));
  std::memcpy(gl.gravity, g->gravity, sizeof(gl.gravity));
  std::memcpy(gl.proj_sqrt_info, g->proj_sqrt_info, sizeof(gl.proj_sqrt_info));
  ceres::swgn::SetGlobals(gl);

  std::vector<std::unique_ptr<IntegrationBase>> pre(g->n_imu);
  std::vector<std::unique_ptr<MarginalizationInfo>> marg(g->n_prior);
  std::vector<std::array<double, 12>> gnss_store(g->n_gnss);  // sat pos, sat vel, base pos, xyzt
  std::vector<Block> blocks;
  // composition A: the reference's own IMUGNSSBase objects, members filled the way AddMargInfo / SetLastImuFactor leave
  // them (gnss_imu_factor.cpp:96-117,245-352); the hidden frames live in `hidden`, which the shim updates after the solve
  const int n_hidden = g->n_chain > 0 ? g->chain_frame_begin[g->n_chain] : 0;
  std::vector<double> hidden((size_t)16 * n_hidden), hidden_lin((size_t)16 * n_hidden);
  std::vector<std::unique_ptr<IMUGNSSBase>> bases;
  std::vector<std::unique_ptr<IntegrationBase>> chain_pre;
  int result = -1;
  {
    ceres::Problem problem;
    for (int b = 0; b < g->n_blocks; ++b) {
      if (g->block_manifold[b] == SWGN_MANIFOLD_POSE) problem.AddParameterBlock(ptr[b], 7, new PoseLocalParameterization());
      else problem.AddParameterBlock(ptr[b], g->block_size[b]);
    }
    auto add = [&](ceres::CostFunction* f, ceres::LossFunction* loss, std::vector<double*> params) {
      problem.AddResidualBlock(f, loss, params);
      blocks.push_back(Block{f, loss != nullptr, params});
    };
    for (int i = 0; i < g->n_proj; ++i)
      add(new projection_factor(Eigen::Vector3d(g->proj_uv[2 * i], g->proj_uv[2 * i + 1], 1.0)), new ceres::CauchyLoss(g->proj_cauchy_a),
          {ptr[g->proj_blocks[3 * i]], ptr[g->proj_blocks[3 * i + 1]], ptr[g->proj_blocks[3 * i + 2]]});
    for (int i = 0; i < g->n_imu; ++i) {
      const double* r = g->imu_data + (size_t)SWGN_IMU_STRIDE * i;
      pre[i].reset(new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero()));
      IntegrationBase& ib = *pre[i];
      for (int k = 0; k < 3; ++k) {
        ib.delta_p(k) = r[SWGN_IMU_DELTA_P + k];
        ib.delta_v(k) = r[SWGN_IMU_DELTA_V + k];
        ib.linearized_ba(k) = r[SWGN_IMU_LIN_BA + k];
        ib.linearized_bg(k) = r[SWGN_IMU_LIN_BG + k];
        ib.gyri(k) = r[SWGN_IMU_GYRI + k];
        ib.gyrj(k) = r[SWGN_IMU_GYRJ + k];
      }
      ib.delta_q = Eigen::Quaterniond(r[SWGN_IMU_DELTA_Q + 3], r[SWGN_IMU_DELTA_Q], r[SWGN_IMU_DELTA_Q + 1], r[SWGN_IMU_DELTA_Q + 2]);
      ib.sum_dt = r[SWGN_IMU_SUM_DT];
      for (int a = 0; a < 15; ++a)
        for (int c = 0; c < 15; ++c) {
          ib.jacobian(a, c) = r[SWGN_IMU_JACOBIAN + a * 15 + c];
          ib.sqrt_info(a, c) = r[SWGN_IMU_SQRT_INFO + a * 15 + c];
        }
      ib.covariance_update = false;
      add(new IMUFactor(&ib), nullptr,
          {ptr[g->imu_blocks[4 * i]], ptr[g->imu_blocks[4 * i + 1]], ptr[g->imu_blocks[4 * i + 2]], ptr[g->imu_blocks[4 * i + 3]]});
    }
    for (int i = 0; i < g->n_gnss; ++i) {
      const double* r = g->gnss_data + (size_t)SWGN_GNSS_STRIDE * i;
      std::memcpy(gnss_store[i].data(), r, sizeof(double) * 9);
      double* sat = gnss_store[i].data();
      double* vel = sat + 3;
      double* base = sat + 6;
      double* xyzt = sat + 9;
      const int32_t* bl = g->gnss_blocks + 3 * i;
      const double meas = r[SWGN_GNSS_MEAS], lam = r[SWGN_GNSS_LAM], wgt = r[SWGN_GNSS_WEIGHT];
      const double el = r[SWGN_GNSS_EL], dt = r[SWGN_GNSS_DT], var = r[SWGN_GNSS_VAR];
      std::vector<double*> p = {ptr[bl[0]], ptr[bl[1]]};
      if (bl[2] >= 0) p.push_back(ptr[bl[2]]);
      switch (g->gnss_kind[i]) {
        case SWGN_GNSS_SPP_PSEUDORANGE: add(new SppPseudorangeFactor(sat, meas, wgt, base), nullptr, p); break;
        case SWGN_GNSS_SPP_CARRIER: add(new SppCarrierPhaseFactor(sat, meas, wgt, base, lam), nullptr, p); break;
        case SWGN_GNSS_RTK_CARRIER: add(new RTKCarrierPhaseFactor(sat, meas, lam, el, dt, var, base, true, 0, 0), nullptr, p); break;
        case SWGN_GNSS_RTK_PSEUDORANGE: add(new RTKPseudorangeFactor(sat, meas, el, dt, var, base), nullptr, p); break;
        case SWGN_GNSS_DOPPLER: add(new SppDopplerFactor(vel, sat, xyzt, meas, wgt, base), nullptr, p); break;
        case SWGN_GNSS_FIXED_INTEGER: add(new FixedIntegerFactor(meas, wgt), nullptr, p); break;
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
        params.push_back(ptr[b]);
      }
      m.linearized_jacobians.resize(m.n, m.n);
      m.linearized_residuals = Eigen::VectorXd(m.n);
      for (int a = 0; a < m.n; ++a) {
        m.linearized_residuals(a) = g->prior_r0[g->prior_r_begin[i] + a];
        for (int c = 0; c < m.n; ++c) m.linearized_jacobians(a, c) = g->prior_J[g->prior_J_begin[i] + (size_t)a * m.n + c];
      }
      add(new MarginalizationFactor(&m), nullptr, params);
    }
    for (int i = 0; i < g->n_unit; ++i) add(new InitialBlackFactor(g->unit_istd[i]), nullptr, {ptr[g->unit_block[i]]});
    if (host_factors) {
      // the reference's initialisation factors, anchored at the generator's ground truth; no adapter is registered for
      // them, so the shim evaluates them on the host
      const double* truth = swgn_synth_truth(S);
      int32_t info8[8];
      swgn_synth_info(S, info8);
      const int F = info8[0];  // frames: poses are blocks 0..F-1, speed-biases F..2F-1
      const double* p1 = truth + g->block_offset[1];
      Eigen::Matrix<double, 6, 6> w6 = Eigen::Matrix<double, 6, 6>::Identity();
      for (int k = 0; k < 6; ++k) w6(k, k) = k < 3 ? 20.0 : 200.0;
      w6(0, 4) = 3.0;  // (not diagonal on purpose)
      add(new InitialPoseFactor(Eigen::Vector3d(p1[0], p1[1], p1[2]), Eigen::Quaterniond(p1[6], p1[3], p1[4], p1[5]), w6), nullptr, {ptr[1]});
      const double* s2 = truth + g->block_offset[F + 2];
      Eigen::Matrix<double, 9, 9> w9 = Eigen::Matrix<double, 9, 9>::Identity();
      for (int k = 0; k < 9; ++k) w9(k, k) = 5.0 + k;
      add(new InitialBiasFactor(Eigen::Vector3d(s2[0], s2[1], s2[2]), Eigen::Vector3d(s2[3], s2[4], s2[5]), Eigen::Vector3d(s2[6], s2[7], s2[8]), w9), nullptr,
          {ptr[F + 2]});
      const double* p3 = truth + g->block_offset[3];
      const Eigen::Matrix3d R3 = Eigen::Quaterniond(p3[6], p3[3], p3[4], p3[5]).toRotationMatrix();
      add(new InitPose0Factor(Eigen::MatrixXd(R3), Eigen::Vector3d(p3[0], p3[1], p3[2]), true, true, 30.0), nullptr, {ptr[3]});
    }
    {
      size_t fN = 0, cN = 0, imu = 0;
      for (int c = 0; c < g->n_chain; ++c) {
        const int b0 = g->chain_blk_begin[c], k = g->chain_blk_begin[c + 1] - b0 - 4;
        const int f0 = g->chain_frame_begin[c], m = g->chain_frame_begin[c + 1] - f0;
        IMUGNSSBase* B = new IMUGNSSBase(ptr[g->chain_blocks[b0]], ptr[g->chain_blocks[b0 + 1]], &problem);
        bases.emplace_back(B);
        for (int i = 0; i < m; ++i) {
          const double* f = g->chain_frame_data + (size_t)SWGN_CHAIN_FRAME_STRIDE * (f0 + i);
          double* h = &hidden[(size_t)16 * (f0 + i)];
          double* hl = &hidden_lin[(size_t)16 * (f0 + i)];
          std::memcpy(h, f + SWGN_CHAIN_POSE, sizeof(double) * 16);
          std::memcpy(hl, f + SWGN_CHAIN_POSE_LIN, sizeof(double) * 16);
          B->gnss_poses.push_back(h);
          B->gnss_speed_bias.push_back(h + 7);
          B->gnss_poses_lin.push_back(hl);
          B->gnss_speed_bias_lin.push_back(hl + 7);
          Eigen::Matrix<double, 15, 15, Eigen::RowMajor> H;
          Eigen::Matrix<double, 15, 1, Eigen::ColMajor> rhs;
          Eigen::Matrix<double, 15, Eigen::Dynamic, Eigen::RowMajor> HN(15, k);
          for (int a = 0; a < 15; ++a) {
            rhs(a) = f[SWGN_CHAIN_RHS + a];
            for (int q = 0; q < 15; ++q) H(a, q) = f[SWGN_CHAIN_HESSIAN + 15 * a + q];
            for (int q = 0; q < k; ++q) HN(a, q) = g->chain_frame_N[fN + ((size_t)i * 15 + a) * k + q];
          }
          B->pose_hessians.push_back(H);
          B->pose_rhses.push_back(rhs);
          B->pose_phase_biases_hessians.push_back(HN);
          chain_pre.emplace_back(integration_from_record(g->chain_imu_data + (size_t)SWGN_IMU_STRIDE * (imu + i)));
          B->imu_factors.push_back(new IMUFactor(chain_pre.back().get()));
        }
        chain_pre.emplace_back(integration_from_record(g->chain_imu_data + (size_t)SWGN_IMU_STRIDE * (imu + m)));
        B->last_imu_factor = new IMUFactor(chain_pre.back().get());
        B->pose1_pose2_hessians.setZero();
        B->phase_biases_hessians.resize(k, k);
        B->phase_biases_rhs.resize(k);
        std::vector<double*> params;
        for (int q = 0; q < 4 + k; ++q) params.push_back(ptr[g->chain_blocks[b0 + q]]);
        for (int a = 0; a < k; ++a) {
          B->phase_biases_rhs(a) = g->chain_N[cN + (size_t)k * k + a];
          for (int q = 0; q < k; ++q) B->phase_biases_hessians(a, q) = g->chain_N[cN + (size_t)a * k + q];
          B->gnss_phase_biases.push_back(params[4 + a]);
        }
        B->gnss_Index = m;
        B->Init();
        add(new IMUGNSSFactor(B), nullptr, params);
        fN += (size_t)m * 15 * k;
        cN += (size_t)k * k + k;
        imu += m + 1;
      }
    }
    for (int b = 0; b < g->n_blocks; ++b)
      if (g->block_const[b]) problem.SetParameterBlockConstant(ptr[b]);

    ceres::Solver::Options options;  // Ceres' defaults: LEVENBERG_MARQUARDT, jacobi_scaling = true
    options.linear_solver_type = ceres::DENSE_SCHUR;
    if (strategy == 0) {
      options.trust_region_strategy_type = ceres::DOGLEG;
      options.jacobi_scaling = false;
    }
    options.max_num_iterations = so.max_num_iterations;
    options.device = device;
    options.linear_solver_ordering = std::make_shared<ceres::ParameterBlockOrdering>();
    for (int b = 0; b < g->n_blocks; ++b)
      if (g->block_group[b] >= 0) options.linear_solver_ordering->AddElementToGroup(ptr[b], g->block_group[b]);
    ceres::internal::parameter_head.clear();
    int32_t info[8];
    swgn_synth_info(S, info);
    for (int k = 0; k < so.n_parameter_head; ++k) ceres::internal::parameter_head.push_back(ptr[info[5] + k]);
    ceres::internal::is_optimize = g_refdemo_is_optimize != 0;
    if (g_refdemo_before_solve) g_refdemo_before_solve(&problem, &options, g, ptr.data());  // (may replace the ordering)
    const double cpu_initial = cpu_cost(blocks, g->proj_cauchy_a);
    ceres::Solver::Summary summary;
    ceres::Solve(options, &problem, &summary);
    if (g_refdemo_after_solve) g_refdemo_after_solve(&problem);
    // the reference's stateful chain factors answer a cost-only call with the linearisation of their last Jacobian
    // evaluation (the CPU call above); forgetting that history makes them eliminate afresh at the returned states, hidden
    // frames included (the shim wrote those back into `hidden`)
    for (auto& B : bases) B->history_flag = false;
    const double cpu_final = cpu_cost(blocks, g->proj_cauchy_a);
    g_last_chain_frames = hidden;
    if (message && message_len > 0) std::snprintf(message, message_len, "%s | %s", summary.message.c_str(), summary.BriefReport().c_str());
    if (summary.termination_type != ceres::FAILURE || summary.num_successful_steps >= 0) {
      result = summary.termination_type;
      if (cost_out) {
        cost_out[0] = summary.initial_cost;
        cost_out[1] = summary.final_cost;
        cost_out[2] = cpu_initial;
        cost_out[3] = cpu_final;
      }
      if (steps_out) {
        steps_out[0] = summary.num_successful_steps;
        steps_out[1] = summary.num_unsuccessful_steps;
      }
    }
    ceres::internal::parameter_head.clear();
  }
  if (state_out)
    for (int b = 0; b < g->n_blocks; ++b) std::memcpy(state_out + g->block_offset[b], ptr[b], sizeof(double) * g->block_size[b]);
  swgn_synth_destroy(S);
  return result;
}

// hidden GNSS-frame states after the last swgn_ceres_refdemo_solve of a composition-A window (16 doubles per frame)
extern "C" int swgn_ceres_refdemo_chain_frames(double* out, int cap_frames) {
  const int n = (int)(g_last_chain_frames.size() / 16);
  if (out && cap_frames >= n) std::memcpy(out, g_last_chain_frames.data(), sizeof(double) * g_last_chain_frames.size());
  return n;
}
extern "C" void swgn_ceres_refdemo_set_hooks(int is_optimize, void (*after_solve)(ceres::Problem*)) {
  g_refdemo_is_optimize = is_optimize;
  g_refdemo_after_solve = after_solve;
}
extern "C" void swgn_ceres_refdemo_set_build_hooks(double* (*block_memory)(int, int),
                                                   void (*before_solve)(ceres::Problem*, ceres::Solver::Options*, const swgn_graph*, double* const*)) {
  g_refdemo_block_memory = block_memory;
  g_refdemo_before_solve = before_solve;
}
// pieces of the window builder for callers that assemble the problem themselves (oracle/ref_estimator_shim.cpp lets the
// reference's own AddAllResidual do it): adapters + application globals of graph g, and a pre-integration object from a record
extern "C" void swgn_ceres_refdemo_prepare(const swgn_graph* g) {
  register_adapters();
  Pbg = Eigen::Vector3d(g->Pbg[0], g->Pbg[1], g->Pbg[2]);
  Rwgw = Eigen::Matrix3d::Identity();
  G = Eigen::Vector3d(g->gravity[0], g->gravity[1], g->gravity[2]);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) projection_factor::sqrt_info(i, j) = g->proj_sqrt_info[2 * i + j];
  ceres::swgn::Globals gl;
  std::memcpy(gl.Pbg, g->Pbg, sizeof(gl.Pbg));
  std::memcpy(gl.gravity, g->gravity, sizeof(gl.gravity));
  std::memcpy(gl.proj_sqrt_info, g->proj_sqrt_info, sizeof(gl.proj_sqrt_info));
  ceres::swgn::SetGlobals(gl);
}
extern "C" void* swgn_ceres_refdemo_integration(const double* record) { return integration_from_record(record); }
