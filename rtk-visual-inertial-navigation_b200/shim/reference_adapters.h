// Device adapters for the reference's factor classes (RVI/factor/*.h).  Templates over the factor
// type so that the application's own classes plug in unchanged: they only read PUBLIC data members
//   projection_factor::pts                       (projection_factor.h:16)
//   IMUFactor::pre_integration -> IntegrationBase {delta_p, delta_q, delta_v, linearized_ba/bg,
//                                  gyr_0i? / gyri, gyrj, sum_dt, jacobian, sqrt_info}  (imu_factor.h:17)
//   RTK*/Spp* {satelite_pos, base_pos, L1_lam | P1 | D1_lam, lam, el, base_rover_time_diff, mea_var, istd}
//                                                 (gnss_factor.h:31-37,63-68,82-85,102-106,124-129)
//   FixedIntegerFactor {N21, istd}, InitialBlackFactor {istd}
//   MarginalizationFactor::marginalization_info -> {n, keep_block_size/idx/data, linearized_jacobians,
//                                                   linearized_residuals}   (marginalization_factor.h:109)
//   IMUGNSSFactor::IMUGNSS_info -> IMUGNSSBase {gnss_poses, gnss_speed_bias, gnss_poses_lin,
//                                  gnss_speed_bias_lin, gnss_phase_biases, pose_hessians, pose_phase_biases_hessians,
//                                  pose_rhses, phase_biases_hessians, phase_biases_rhs, imu_factors, last_imu_factor}
//                                                 (gnss_imu_factor.h:52-77, 134)
// Vector members are read through operator[] / operator(), so Eigen types and plain arrays both
// work.  Register once per process, e.g.
//   ceres::swgn::RegisterAdapter(typeid(projection_factor), &swgn_adapters::projection<projection_factor>);
#ifndef SWGN_REFERENCE_ADAPTERS_H_
#define SWGN_REFERENCE_ADAPTERS_H_
#include <cmath>

#include "ceres/swgn_adapter.h"

namespace swgn_adapters {
using ceres::CostFunction;
using ceres::swgn::FactorRecord;

template <class F>
bool projection(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  out->kind = ceres::swgn::kProjection;
  out->data = {f->pts[0], f->pts[1]};  // pts.z is 1 on the normalised plane (projection_factor.cpp:27)
  return true;
}

// the weight the factor multiplies in: 1/sqrt(varerr2(el, dt, var)) with the reference's
// single-precision sinf (gnss_factor.cpp:98-103)
inline double rtk_weight(double el, double dt, double mea_var) {
  const double b = 299792458.0 * 5e-12 * dt;
  const double sinel = sinf(el);
  return 1.0 / std::sqrt((mea_var / sinel / sinel) + b * b);
}
inline void gnss_common(FactorRecord* out, int kind, const double* sat, const double* vel, const double* base, double meas,
                        double lam, double w) {
  out->kind = ceres::swgn::kGnss;
  out->gnss_kind = kind;
  out->data.assign(SWGN_GNSS_STRIDE, 0.0);
  for (int i = 0; i < 3; ++i) {
    out->data[SWGN_GNSS_SAT_POS + i] = sat[i];
    out->data[SWGN_GNSS_SAT_VEL + i] = vel ? vel[i] : 0.0;
    out->data[SWGN_GNSS_BASE_POS + i] = base[i];
  }
  out->data[SWGN_GNSS_MEAS] = meas;
  out->data[SWGN_GNSS_LAM] = lam;
  out->data[SWGN_GNSS_WEIGHT] = w;
}
template <class F>
bool rtk_carrier_phase(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  // use_istd = false: sqrt_info is 1 (gnss_factor.cpp:116-117), as the per-epoch phase-bias initialisation builds the
  // factor (swf_gnss.cpp:355,367: use_istd = false, mea_var = 0)
  gnss_common(out, SWGN_GNSS_RTK_CARRIER, f->satelite_pos, nullptr, f->base_pos, f->L1_lam, f->lam,
              f->use_istd ? rtk_weight(f->el, f->base_rover_time_diff, f->mea_var) : 1.0);
  out->data[SWGN_GNSS_EL] = f->el;
  out->data[SWGN_GNSS_DT] = f->base_rover_time_diff;
  out->data[SWGN_GNSS_VAR] = f->mea_var;
  return true;
}
template <class F>
bool rtk_pseudorange(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  gnss_common(out, SWGN_GNSS_RTK_PSEUDORANGE, f->satelite_pos, nullptr, f->base_pos, f->P1, 0.0,
              rtk_weight(f->el, f->base_rover_time_diff, f->mea_var));
  out->data[SWGN_GNSS_EL] = f->el;
  out->data[SWGN_GNSS_DT] = f->base_rover_time_diff;
  out->data[SWGN_GNSS_VAR] = f->mea_var;
  return true;
}
template <class F>
bool spp_pseudorange(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  gnss_common(out, SWGN_GNSS_SPP_PSEUDORANGE, f->satelite_pos, nullptr, f->base_pos, f->P1, 0.0, f->istd);
  return true;
}
template <class F>
bool spp_carrier_phase(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  gnss_common(out, SWGN_GNSS_SPP_CARRIER, f->satelite_pos, nullptr, f->base_pos, f->L1_lam, f->lam, f->istd);
  return true;
}
template <class F>
bool spp_doppler(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  gnss_common(out, SWGN_GNSS_DOPPLER, f->satelite_pos, f->satelitev1, f->base_pos, f->D1_lam, 0.0, f->istd);
  return true;
}
template <class F>
bool fixed_integer(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  const double zero[3] = {0, 0, 0};
  gnss_common(out, SWGN_GNSS_FIXED_INTEGER, zero, nullptr, zero, f->N21, 0.0, f->istd);
  return true;
}
template <class F>
bool unit_prior(const CostFunction* cf, FactorRecord* out) {  // InitialBlackFactor
  const F* f = static_cast<const F*>(cf);
  out->kind = ceres::swgn::kUnit;
  out->data = {f->istd};
  return true;
}
// IntegrationBase-like object -> SWGN_IMU_STRIDE record
template <class P>
void imu_record(const P* p, double* r) {
  for (int i = 0; i < 3; ++i) {
    r[SWGN_IMU_DELTA_P + i] = p->delta_p[i];
    r[SWGN_IMU_DELTA_V + i] = p->delta_v[i];
    r[SWGN_IMU_LIN_BA + i] = p->linearized_ba[i];
    r[SWGN_IMU_LIN_BG + i] = p->linearized_bg[i];
    r[SWGN_IMU_GYRI + i] = p->gyri[i];
    r[SWGN_IMU_GYRJ + i] = p->gyrj[i];
  }
  r[SWGN_IMU_DELTA_Q] = p->delta_q.x();
  r[SWGN_IMU_DELTA_Q + 1] = p->delta_q.y();
  r[SWGN_IMU_DELTA_Q + 2] = p->delta_q.z();
  r[SWGN_IMU_DELTA_Q + 3] = p->delta_q.w();
  r[SWGN_IMU_SUM_DT] = p->sum_dt;
  for (int i = 0; i < 15; ++i)
    for (int j = 0; j < 15; ++j) {
      r[SWGN_IMU_JACOBIAN + i * 15 + j] = p->jacobian(i, j);
      r[SWGN_IMU_SQRT_INFO + i * 15 + j] = p->sqrt_info(i, j);
    }
}
// IMUFactor: P = IntegrationBase-like object reachable as f->pre_integration
template <class F>
bool imu(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  out->kind = ceres::swgn::kImu;
  out->data.assign(SWGN_IMU_STRIDE, 0.0);
  imu_record(f->pre_integration, out->data.data());
  return true;
}
// IMUGNSSFactor: B = IMUGNSSBase-like object reachable as f->IMUGNSS_info.  The middle-marginalisation
// link (gnss_middle_marginfo, pose1_pose2_hessians) has no device representation: such a factor is
// reported as unsupported instead of being evaluated differently.
template <class F>
bool imu_gnss(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  const auto* b = f->IMUGNSS_info;
  if (b->gnss_middle_marginfo) return false;
  const int m = (int)b->gnss_poses.size(), k = (int)b->gnss_phase_biases.size();
  if (m < 1 || (int)b->imu_factors.size() != m || !b->last_imu_factor) return false;
  out->kind = ceres::swgn::kChain;
  out->chain_m = m;
  out->chain_frames.assign((size_t)m * SWGN_CHAIN_FRAME_STRIDE, 0.0);
  out->chain_frame_N.assign((size_t)m * 15 * k, 0.0);
  out->chain_N.assign((size_t)k * k + k, 0.0);
  out->chain_imu.assign((size_t)(m + 1) * SWGN_IMU_STRIDE, 0.0);
  for (int i = 0; i < m; ++i) {
    double* fr = out->chain_frames.data() + (size_t)SWGN_CHAIN_FRAME_STRIDE * i;
    for (int q = 0; q < 7; ++q) {
      fr[SWGN_CHAIN_POSE + q] = b->gnss_poses[i][q];
      fr[SWGN_CHAIN_POSE_LIN + q] = b->gnss_poses_lin[i][q];
    }
    for (int q = 0; q < 9; ++q) {
      fr[SWGN_CHAIN_SB + q] = b->gnss_speed_bias[i][q];
      fr[SWGN_CHAIN_SB_LIN + q] = b->gnss_speed_bias_lin[i][q];
    }
    for (int r = 0; r < 15; ++r) {
      fr[SWGN_CHAIN_RHS + r] = b->pose_rhses[i](r);
      for (int c = 0; c < 15; ++c) fr[SWGN_CHAIN_HESSIAN + r * 15 + c] = b->pose_hessians[i](r, c);
      for (int c = 0; c < k; ++c) out->chain_frame_N[((size_t)i * 15 + r) * k + c] = b->pose_phase_biases_hessians[i](r, c);
    }
    imu_record(b->imu_factors[i]->pre_integration, out->chain_imu.data() + (size_t)SWGN_IMU_STRIDE * i);
    out->chain_pose_ptr.push_back(b->gnss_poses[i]);
    out->chain_sb_ptr.push_back(b->gnss_speed_bias[i]);
  }
  imu_record(b->last_imu_factor->pre_integration, out->chain_imu.data() + (size_t)SWGN_IMU_STRIDE * m);
  for (int r = 0; r < k; ++r) {
    out->chain_N[(size_t)k * k + r] = b->phase_biases_rhs(r);
    for (int c = 0; c < k; ++c) out->chain_N[(size_t)r * k + c] = b->phase_biases_hessians(r, c);
  }
  return true;
}
// MarginalizationFactor: M = MarginalizationInfo-like object reachable as f->marginalization_info
template <class F>
bool marginalization(const CostFunction* cf, FactorRecord* out) {
  const F* f = static_cast<const F*>(cf);
  const auto* m = f->marginalization_info;
  out->kind = ceres::swgn::kPrior;
  const int n = m->n;
  out->prior_n = n;
  const int nb = (int)m->keep_block_size.size();
  for (int b = 0; b < nb; ++b) {
    out->prior_blk_idx.push_back(m->keep_block_idx[b] - m->m);
    for (int k = 0; k < m->keep_block_size[b]; ++k) out->prior_x0.push_back(m->keep_block_data[b][k]);
  }
  out->prior_J.resize((size_t)n * n);
  out->prior_r0.resize(n);
  for (int i = 0; i < n; ++i) {
    out->prior_r0[i] = m->linearized_residuals(i);
    for (int j = 0; j < n; ++j) out->prior_J[(size_t)i * n + j] = m->linearized_jacobians(i, j);
  }
  return true;
}
}  // namespace swgn_adapters
#endif
