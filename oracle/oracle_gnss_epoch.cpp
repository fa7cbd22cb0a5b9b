// TEST INFRASTRUCTURE -- see oracle_core.h.
//
// CPU restatement of the reference's per-epoch GNSS preprocessing, written the way the reference is: raw pointers
// into parameter storage, std::list<PB> per satellite and frequency, a MarginalizationInfo that keys parameter
// blocks by ADDRESS.
//   update_azel / ecef2pos / satazel / xyz2enu   RVI/gnss/src/common_function.cpp:84-124,142-162,394-408
//   SWFOptimization::GnssPreprocess               RVI/swf/swf_gnss.cpp:265-587
//   SWFOptimization::AddGnssResidual              RVI/swf/swf_core.cpp:87-205 (+ the ADDRESIDUAL macro :10-48)
//   ResidualBlockInfo::Evaluate, MarginalizationInfo::{addResidualBlockInfo, marginalize, getParameterBlocks}
//                                                 RVI/factor/marginalization_factor.cpp:7-70,260-400
// Pin status: update_azel is checked bit-exact against the reference's own common_function.cpp compiled into
// oracle/_ref/libref_gnss.so; the factors it evaluates are pinned bit-exact on the reference's classes
// (tests/test_oracle_ref_factors.py); the epoch's prior is pinned on the reference's own MarginalizationInfo::marginalize
// (marginalization_factor.cpp compiled into oracle/_ref, fed with the reference's own factor classes through
// oracle/ref_marg_shim.cpp) and on an independent numpy Schur complement (tests/test_gnss_epoch.py).  GnssPreprocess itself
// is a member of SWFOptimization (ROS / OpenCV / whole estimator) and cannot be compiled here: its bookkeeping is restated,
// not pinned by execution -- PARITY UNPINNED for :265-500.
//
// The one liberty: the reference keeps parameter blocks in std::unordered_map<long, ...> keyed by ADDRESS, so the order of
// its keep blocks is whatever that container iterates in (implementation-defined); here an ordered map is used and all
// parameters of the epoch are laid out in one array (pose, speed-bias, blackvalue, clocks, ambiguities: RTK, SPP,
// pseudorange correction in observation order) so that the order is that order.  Only a permutation of J0's columns.
#include <algorithm>
#include <list>
#include <numeric>

#include "../include/swgn_gnss.h"
#include "oracle_core.h"

namespace oracle {
namespace {
const double kPI = 3.1415926535897932;
const double kRE_WGS84 = 6378137.0;
const double kFE_WGS84 = 1.0 / 298.257223563;
const double kEps = 1e-8;  // marginalization_factor.h: eps

void ecef2pos(const double* r, double* pos) {  // common_function.cpp:111-123
  double e2 = kFE_WGS84 * (2.0 - kFE_WGS84), r2 = dot_rtk(r, r, 2), z, zk, v = kRE_WGS84, sinp;
  for (z = r[2], zk = 0.0; std::fabs(z - zk) >= 1E-4;) {
    zk = z;
    sinp = z / std::sqrt(r2 + z * z);
    v = kRE_WGS84 / std::sqrt(1.0 - e2 * sinp * sinp);
    z = r[2] + v * e2 * sinp;
  }
  pos[0] = r2 > 1E-12 ? std::atan(z / std::sqrt(r2)) : (r[2] > 0.0 ? kPI / 2.0 : -kPI / 2.0);
  pos[1] = r2 > 1E-12 ? std::atan2(r[1], r[0]) : 0.0;
  pos[2] = std::sqrt(r2 + z * z) - v;
}
void xyz2enu(const double* pos, double* E) {  // :150-162, column-major
  double sinp = std::sin(pos[0]), cosp = std::cos(pos[0]), sinl = std::sin(pos[1]), cosl = std::cos(pos[1]);
  E[0] = -sinl;
  E[3] = cosl;
  E[6] = 0.0;
  E[1] = -sinp * cosl;
  E[4] = -sinp * sinl;
  E[7] = cosp;
  E[2] = cosp * cosl;
  E[5] = cosp * sinl;
  E[8] = sinp;
}
void ecef2enu(const double* pos, const double* r, double* e) {  // :142-147, matmul("NN", 3, 1, 3, ...)
  double E[9];
  xyz2enu(pos, E);
  for (int i = 0; i < 3; ++i) {
    double d = 0.0;
    for (int x = 0; x < 3; ++x) d += E[i + x * 3] * r[x];
    e[i] = 1.0 * d;
  }
}
double satazel(const double* pos, const double* e, double* azel) {  // :84-100
  double az = 0.0, el = kPI / 2.0, enu[3];
  if (pos[2] > -kRE_WGS84) {
    ecef2enu(pos, e, enu);
    az = dot_rtk(enu, enu, 2) < 1E-12 ? 0.0 : std::atan2(enu[0], enu[1]);
    if (az < 0.0) az += 2 * kPI;
    el = std::asin(enu[2]);
  }
  if (azel) {
    azel[0] = az;
    azel[1] = el;
  }
  return el;
}
}  // namespace

void update_azel(const double globalxyz[3], swgn_epoch* rover) {  // :394-408
  for (int i = 0; i < rover->n_obs; i++) {
    swgn_obs* d = rover->obs + i;
    if (d->svh != 0) continue;
    double pos[3], e2[3], azel[2], e[3];
    ecef2pos(globalxyz, pos);
    distance_rtk(globalxyz, d->sat_pos, e);
    e2[0] = -e[0];
    e2[1] = -e[1];
    e2[2] = -e[2];
    satazel(pos, e2, azel);
    d->el = azel[1];
  }
}

// PBtype, common_function.h:47-69
struct PB {
  double value = 0;
  uint8_t SLIP_COUNT = 0, half_flag = 0, sys = 0, f = 0;
  int continue_count = 0;
  double last_update_time = 0;
  int handle = -1, sat = 0;
};
struct GnssTracker {
  swgn_gnss_config cfg;
  std::list<PB> lists[3][SWGN_MAXSAT * 2];  // rtk_phase_bias_variables, spp_phase_bias_variables, pseudorange_correction_variables
  std::vector<PB*> by_handle[3];
  PB* push(int fam, int sat, int sys, int f) {
    PB n;
    n.sys = (uint8_t)sys;
    n.f = (uint8_t)f;
    n.value = 0;
    n.continue_count = 0;
    n.sat = sat;
    n.handle = (int)by_handle[fam].size();
    lists[fam][sat * 2 + f].push_back(n);
    PB* p = &lists[fam][sat * 2 + f].back();
    by_handle[fam].push_back(p);
    return p;
  }
};

// ResidualBlockInfo + MarginalizationInfo, marginalization_factor.cpp
struct ResidualInfo {
  std::shared_ptr<CostFunction> cost;
  std::vector<double*> parameter_blocks;
  std::vector<int> drop_set;
  std::vector<double> residuals;
  std::vector<Mat> jacobians;
  double cauchy_a = 0.0;  // > 0: loss_function = CauchyLoss(a)
  void Evaluate() {  // :7-43
    residuals.assign(cost->num_residuals, 0.0);
    std::vector<std::vector<double>> raw(cost->block_sizes.size());
    std::vector<double*> ptr(cost->block_sizes.size());
    for (size_t i = 0; i < raw.size(); ++i) {
      raw[i].assign((size_t)cost->num_residuals * cost->block_sizes[i], 0.0);
      ptr[i] = raw[i].data();
    }
    cost->Evaluate(parameter_blocks.data(), residuals.data(), ptr.data());
    jacobians.clear();
    for (size_t i = 0; i < raw.size(); ++i) {
      Mat J(cost->num_residuals, cost->block_sizes[i]);
      J.a = raw[i];
      jacobians.push_back(J);
    }
    if (cauchy_a > 0.0) {  // :21-41
      double sq_norm = 0.0, rho[3];
      for (double v : residuals) sq_norm += v * v;
      cauchy_loss(cauchy_a, sq_norm, rho);
      const double sqrt_rho1_ = std::sqrt(rho[1]);
      double residual_scaling_, alpha_sq_norm_;
      if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
        residual_scaling_ = sqrt_rho1_;
        alpha_sq_norm_ = 0.0;
      } else {
        const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
        residual_scaling_ = sqrt_rho1_ / (1 - alpha);
        alpha_sq_norm_ = alpha / sq_norm;
      }
      const int nr = cost->num_residuals;
      for (Mat& J : jacobians) {  // J = sqrt_rho1 (J - alpha_sq_norm r (r' J))
        for (int c = 0; c < J.c; ++c) {
          double rtJ = 0.0;
          for (int r = 0; r < nr; ++r) rtJ += residuals[r] * J(r, c);
          for (int r = 0; r < nr; ++r) J(r, c) = sqrt_rho1_ * (J(r, c) - alpha_sq_norm_ * residuals[r] * rtJ);
        }
      }
      for (double& v : residuals) v *= residual_scaling_;
    }
  }
};
struct MargInfo {
  std::vector<ResidualInfo> factors;
  std::map<long, int> parameter_block_size, parameter_block_idx, parameter_block_drop_idx;
  std::map<long, std::vector<double>> parameter_block_data;
  int m = 0, n = 0;
  Mat A, linearized_jacobians;
  std::vector<double> b, linearized_residuals;
  std::vector<int> keep_block_size, keep_block_idx;
  std::vector<std::vector<double>> keep_block_data;
  std::vector<double*> keep_block_addr;
  static int localSize(int size) { return size == 7 ? 6 : size; }

  void addResidualBlockInfo(const ResidualInfo& info) {  // :58-79
    factors.push_back(info);
    for (size_t i = 0; i < info.parameter_blocks.size(); ++i)
      parameter_block_size[reinterpret_cast<long>(info.parameter_blocks[i])] = info.cost->block_sizes[i];
    for (int d : info.drop_set) parameter_block_drop_idx[reinterpret_cast<long>(info.parameter_blocks[d])] = 0;
  }
  void marginalize() {  // :260-377 (initialinformation = true; the multithread flag is overwritten with false)
    int pos = 0;
    for (auto& it : parameter_block_drop_idx)
      if (parameter_block_idx.find(it.first) == parameter_block_idx.end()) {
        parameter_block_idx[it.first] = pos;
        pos += localSize(parameter_block_size[it.first]);
      }
    m = pos;
    for (auto& it : parameter_block_size)
      if (parameter_block_idx.find(it.first) == parameter_block_idx.end()) {
        parameter_block_idx[it.first] = pos;
        pos += localSize(it.second);
      }
    n = pos - m;
    if (n == 0) return;
    Mat Afull(pos, pos);
    std::vector<double> bfull(pos, 0.0);
    for (ResidualInfo& it : factors) {  // ThreadsConstructA :92-117
      it.Evaluate();
      const int nr = it.cost->num_residuals;
      for (size_t i = 0; i < it.parameter_blocks.size(); ++i) {
        const int idx_i = parameter_block_idx[reinterpret_cast<long>(it.parameter_blocks[i])];
        const int size_i = localSize(parameter_block_size[reinterpret_cast<long>(it.parameter_blocks[i])]);
        for (size_t j = i; j < it.parameter_blocks.size(); ++j) {
          const int idx_j = parameter_block_idx[reinterpret_cast<long>(it.parameter_blocks[j])];
          const int size_j = localSize(parameter_block_size[reinterpret_cast<long>(it.parameter_blocks[j])]);
          for (int a = 0; a < size_i; ++a)
            for (int c = 0; c < size_j; ++c) {
              double s = 0.0;
              for (int r = 0; r < nr; ++r) s += it.jacobians[i](r, a) * it.jacobians[j](r, c);
              Afull(idx_i + a, idx_j + c) += s;
              if (i != j) Afull(idx_j + c, idx_i + a) = Afull(idx_i + a, idx_j + c);
            }
        }
        for (int a = 0; a < size_i; ++a) {
          double s = 0.0;
          for (int r = 0; r < nr; ++r) s += it.jacobians[i](r, a) * it.residuals[r];
          bfull[idx_i + a] += s;
        }
      }
    }
    A = Mat(n, n);
    b.assign(n, 0.0);
    if (m != 0) {  // :331-345
      Mat Amm(m, m);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Amm(i, j) = 0.5 * (Afull(i, j) + Afull(j, i));
      std::vector<double> w;
      Mat V;
      eig_sym(Amm, &w, &V);
      Mat Amm_inv(m, m);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
          double s = 0.0;
          for (int k = 0; k < m; ++k) s += V(i, k) * (w[k] > kEps ? 1.0 / w[k] : 0.0) * V(j, k);
          Amm_inv(i, j) = s;
        }
      Mat T(n, m);  // Arm * Amm_inv
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j) {
          double s = 0.0;
          for (int k = 0; k < m; ++k) s += Afull(m + i, k) * Amm_inv(k, j);
          T(i, j) = s;
        }
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
          double s = 0.0;
          for (int k = 0; k < m; ++k) s += T(i, k) * Afull(k, m + j);
          A(i, j) = Afull(m + i, m + j) - s;
        }
        double s = 0.0;
        for (int k = 0; k < m; ++k) s += T(i, k) * bfull[k];
        b[i] = bfull[m + i] - s;
      }
    } else {
      A = Afull;
      b = bfull;
    }
    std::vector<double> w;  // :348-358
    Mat V;
    eig_sym(A, &w, &V);
    linearized_jacobians = Mat(n, n);
    linearized_residuals.assign(n, 0.0);
    for (int k = 0; k < n; ++k) {
      const double S = w[k] > kEps ? w[k] : 0.0, S_inv = w[k] > kEps ? 1.0 / w[k] : 0.0;
      const double S_sqrt = std::sqrt(S), S_inv_sqrt = std::sqrt(S_inv);
      double s = 0.0;
      for (int i = 0; i < n; ++i) {
        linearized_jacobians(k, i) = S_sqrt * V(i, k);
        s += V(i, k) * b[i];
      }
      linearized_residuals[k] = S_inv_sqrt * s;
    }
    for (ResidualInfo& it : factors)  // :360-372
      for (size_t i = 0; i < it.parameter_blocks.size(); ++i) {
        const long addr = reinterpret_cast<long>(it.parameter_blocks[i]);
        if (parameter_block_data.find(addr) == parameter_block_data.end())
          parameter_block_data[addr].assign(it.parameter_blocks[i], it.parameter_blocks[i] + it.cost->block_sizes[i]);
      }
  }
  void getParameterBlocks() {  // :381-399
    keep_block_size.clear();
    keep_block_idx.clear();
    keep_block_data.clear();
    keep_block_addr.clear();
    for (const auto& it : parameter_block_idx)
      if (it.second >= m) {
        keep_block_size.push_back(parameter_block_size[it.first]);
        keep_block_idx.push_back(it.second);
        keep_block_data.push_back(parameter_block_data[it.first]);
        keep_block_addr.push_back(reinterpret_cast<double*>(it.first));
      }
  }
};

namespace {
enum { NormalMode = 0, MargeExcludeMode = 1 };
struct Added {  // what AddGnssResidual produced, for the second (ceres::Problem) use
  int gnss_kind;  // -1: InitialBlackFactor
  std::vector<double> record;
  std::vector<double*> parameter_blocks;
};

struct EpochContext {
  GnssTracker* T;
  swgn_epoch* rover;
  swgn_gnss_frame* frame;
  std::vector<PB*> RTK_Npoint, SPP_Npoint, PC_Npoint;  // [obs * NFREQ + f]
  // parameter storage in address order: pose 7 | speed-bias 9 | blackvalue | gnss_dt 13 | ambiguity values
  std::vector<double> store;
  double* para_pose() { return &store[0]; }
  double* para_speed_bias() { return &store[7]; }
  double* blackvalue() { return &store[16]; }
  double* para_gnss_dt() { return &store[17]; }
  std::vector<std::pair<int, PB*>> amb_order;  // (family, PB) in storage order
  double* amb_value(int fam, PB* p) {
    for (size_t i = 0; i < amb_order.size(); ++i)
      if (amb_order[i].first == fam && amb_order[i].second == p) return &store[30 + i];
    return nullptr;
  }
};

void gnss_record(double* rec, const swgn_obs* d, const double* base, double meas, double lam, double w, bool doppler) {
  std::fill(rec, rec + SWGN_GNSS_STRIDE, 0.0);
  for (int i = 0; i < 3; ++i) {
    rec[SWGN_GNSS_SAT_POS + i] = d->sat_pos[i];
    rec[SWGN_GNSS_SAT_VEL + i] = doppler ? d->sat_vel[i] : 0.0;
    rec[SWGN_GNSS_BASE_POS + i] = base[i];
  }
  rec[SWGN_GNSS_MEAS] = meas;
  rec[SWGN_GNSS_LAM] = lam;
  rec[SWGN_GNSS_WEIGHT] = w;
}

// swf_core.cpp:87-205 with the ADDRESIDUAL macro (:10-48) for the two modes GnssPreprocess uses
void AddGnssResidual(int mode, const std::set<double*>& MargePoint, MargInfo* marginalization_info, std::vector<Added>* problem,
                     EpochContext& C) {
  const swgn_gnss_config& cfg = C.T->cfg;
  swgn_epoch* rover = C.rover;
  bool have_base = false;
  double globalxyz[3] = {C.para_pose()[0] + rover->base_xyz[0], C.para_pose()[1] + rover->base_xyz[1],
                         C.para_pose()[2] + rover->base_xyz[2]};
  update_azel(globalxyz, rover);
  auto ADDRESIDUAL = [&](std::vector<double*> parameter_block_vector, CostFunction* factor, int kind, const double* rec) {
    if (mode == NormalMode) {
      Added a;
      a.gnss_kind = kind;
      if (rec) a.record.assign(rec, rec + SWGN_GNSS_STRIDE);
      a.parameter_blocks = parameter_block_vector;
      problem->push_back(a);
      delete factor;
      return;
    }
    std::vector<int> dropset, keepset;
    for (int vi = 0; vi < (int)parameter_block_vector.size(); vi++) {
      if (MargePoint.find(parameter_block_vector[vi]) != MargePoint.end())
        dropset.push_back(vi);
      else
        keepset.push_back(vi);
    }
    dropset = keepset;  // MargeExcludeMode: everything NOT in the set is dropped (:41)
    ResidualInfo info;
    info.cost.reset(factor);
    info.parameter_blocks = parameter_block_vector;
    info.drop_set = dropset;
    marginalization_info->addResidualBlockInfo(info);
  };
  double rec[SWGN_GNSS_STRIDE];
  ADDRESIDUAL({C.blackvalue()}, make_unit_factor(1), -1, nullptr);  // InitialBlackFactor(1), :101-103
  if (cfg.use_rtk) {
    for (int i = 0; i < rover->n_obs; i++) {
      swgn_obs* d = rover->obs + i;
      const int sys = d->sys;
      for (int f = 0; f < SWGN_NFREQ; f++) {
        PB* N = C.RTK_Npoint[i * SWGN_NFREQ + f];
        if (!N) continue;
        if (d->el < cfg.azelmin) continue;
        have_base = true;
        const double lam = cfg.lams[d->sys][f];
        gnss_record(rec, d, rover->base_xyz, d->rtk_l[f] * lam, lam,
                    1 / std::sqrt(varerr2(d->el, rover->br_time_diff, std::pow(d->rtk_lstd[f] * lam, 2))), false);
        rec[SWGN_GNSS_EL] = d->el, rec[SWGN_GNSS_DT] = rover->br_time_diff, rec[SWGN_GNSS_VAR] = std::pow(d->rtk_lstd[f] * lam, 2);
        ADDRESIDUAL({C.para_pose(), C.amb_value(SWGN_AMB_RTK, N), C.para_gnss_dt() + sys * 2 + f},
                    make_gnss_factor(SWGN_GNSS_RTK_CARRIER, rec), SWGN_GNSS_RTK_CARRIER, rec);
      }
    }
  }
  if (cfg.use_rtd) {
    for (int i = 0; i < rover->n_obs; i++) {
      swgn_obs* d = rover->obs + i;
      const int sys = d->sys;
      for (int f = 0; f < SWGN_NFREQ; f++) {
        if (d->rtk_p[f] == 0.0 || d->svh != 0 || d->rtk_pstd[f] > 2) continue;
        if (d->el < cfg.azelmin) continue;
        have_base = true;
        gnss_record(rec, d, rover->base_xyz, d->rtk_p[f], 0.0,
                    1 / std::sqrt(varerr2(d->el, rover->br_time_diff, std::pow(d->rtk_pstd[f], 2))), false);
        rec[SWGN_GNSS_EL] = d->el, rec[SWGN_GNSS_DT] = rover->br_time_diff, rec[SWGN_GNSS_VAR] = std::pow(d->rtk_pstd[f], 2);
        ADDRESIDUAL({C.para_pose(), C.para_gnss_dt() + sys * 2 + f}, make_gnss_factor(SWGN_GNSS_RTK_PSEUDORANGE, rec),
                    SWGN_GNSS_RTK_PSEUDORANGE, rec);
      }
    }
  }
  for (int i = 0; i < rover->n_obs; i++) {
    swgn_obs* d = rover->obs + i;
    if (d->svh != 0) continue;
    if (d->el < cfg.azelmin) continue;
    if (d->spp_p[0] != 0.0 && d->spp_pstd[0] < 2 && !have_base) {
      double sin_el = std::sin(d->el);
      double istd = sin_el * sin_el /
                    std::sqrt(d->spp_pstd[0] * d->spp_pstd[0] +
                              (d->ion_var * 0.125 * 0.125 + d->trop_var * 0.7 * 0.7 + d->sat_var * 0.35 * 0.35 + 1));
      if (C.frame->epochs_since_start < 100) istd *= 10;
      gnss_record(rec, d, rover->base_xyz, d->spp_p[0], 0.0, istd, false);
      ADDRESIDUAL({C.para_pose(), C.para_gnss_dt() + 6 + d->sys * 2 + 0}, make_gnss_factor(SWGN_GNSS_SPP_PSEUDORANGE, rec),
                  SWGN_GNSS_SPP_PSEUDORANGE, rec);
    }
    if (cfg.use_spp_phase && d->spp_l[0] != 0.0 && C.SPP_Npoint[i * SWGN_NFREQ + 0]) {
      double lam = cfg.lams[d->sys][0];
      double sin_el = std::sin(d->el);
      double x = d->spp_lstd[0] * lam;
      double istd = sin_el * sin_el / std::sqrt(x * x + (d->ion_var * 0.125 * 0.125 + d->trop_var * 0.7 * 0.7 + d->sat_var * 0.35 * 0.35));
      gnss_record(rec, d, rover->base_xyz, d->spp_l[0] * lam, lam, istd, false);
      ADDRESIDUAL({C.para_pose(), C.para_gnss_dt() + 6 + d->sys * 2 + 0, C.amb_value(SWGN_AMB_SPP, C.SPP_Npoint[i * SWGN_NFREQ + 0])},
                  make_gnss_factor(SWGN_GNSS_SPP_CARRIER, rec), SWGN_GNSS_SPP_CARRIER, rec);
    }
    if (cfg.use_spp_correction && d->spp_p0[0] != 0.0 && C.PC_Npoint[i * SWGN_NFREQ + 0]) {
      double lam = cfg.lams[d->sys][0];
      double sin_el = std::sin(d->el);
      double istd = sin_el * sin_el /
                    std::sqrt(d->spp_pstd[0] * d->spp_pstd[0] + (d->ion_var * 0.125 * 0.125 + d->trop_var * 0.7 * 0.7 + d->sat_var * 0.35 * 0.35));
      gnss_record(rec, d, rover->base_xyz, d->spp_p0[0], lam, istd, false);
      ADDRESIDUAL({C.para_pose(), C.para_gnss_dt() + 6 + d->sys * 2 + 0, C.amb_value(SWGN_AMB_PCORR, C.PC_Npoint[i * SWGN_NFREQ + 0])},
                  make_gnss_factor(SWGN_GNSS_SPP_CARRIER, rec), SWGN_GNSS_SPP_CARRIER, rec);
    }
  }
  if (cfg.use_doppler) {
    for (int i = 0; i < rover->n_obs; i++) {
      swgn_obs* d = rover->obs + i;
      if (d->spp_d[0] == 0.0 || d->svh != 0) continue;
      if (d->spp_dstd[0] > 2) continue;
      if (d->el < cfg.azelmin) continue;
      double istd = std::sin(d->el) * std::sin(d->el) / (d->spp_dstd[0] * cfg.lams[d->sys][0]);
      gnss_record(rec, d, rover->base_xyz, d->spp_d[0] * cfg.lams[d->sys][0], 0.0, istd, true);
      ADDRESIDUAL({C.para_speed_bias(), C.para_gnss_dt() + 12, C.para_pose()}, make_gnss_factor(SWGN_GNSS_DOPPLER, rec),
                  SWGN_GNSS_DOPPLER, rec);
    }
  }
}
}  // namespace

namespace {
// the epoch's parameter storage: pose | speed-bias | blackvalue | gnss_dt | ambiguity values (at 0: PhaseBiasSaveAndReset)
void layout_store(EpochContext& C, swgn_epoch* data, swgn_gnss_frame* frame) {
  const int NF = SWGN_NFREQ;
  for (int fam = 0; fam < 3; ++fam)
    for (int i = 0; i < data->n_obs; i++)
      for (int f = 0; f < NF; f++) {
        PB* p = fam == 0 ? C.RTK_Npoint[i * NF + f] : fam == 1 ? C.SPP_Npoint[i * NF + f] : C.PC_Npoint[i * NF + f];
        if (!p) continue;
        bool dup = false;
        for (auto& a : C.amb_order) dup |= a.first == fam && a.second == p;
        if (!dup) C.amb_order.push_back({fam, p});
      }
  C.store.assign(30 + C.amb_order.size(), 0.0);
  std::copy(frame->pose, frame->pose + 7, C.para_pose());
  std::copy(frame->speed_bias, frame->speed_bias + 9, C.para_speed_bias());
  *C.blackvalue() = frame->blackvalue;
  std::copy(frame->gnss_dt, frame->gnss_dt + SWGN_GNSS_NCLK, C.para_gnss_dt());
}
}  // namespace

// GnssPreprocess, swf_gnss.cpp:265-587.  Returns 0, or a negative code when the output buffers are too small.
int gnss_preprocess(GnssTracker* T, swgn_epoch* data, swgn_gnss_frame* frame, swgn_gnss_output* out) {
  const swgn_gnss_config& cfg = T->cfg;
  swgn_obs* d;
  int i;
  const int NF = SWGN_NFREQ;
  EpochContext C;
  C.T = T;
  C.rover = data;
  C.frame = frame;
  C.RTK_Npoint.assign(data->n_obs * NF, nullptr);
  C.SPP_Npoint.assign(data->n_obs * NF, nullptr);
  C.PC_Npoint.assign(data->n_obs * NF, nullptr);
  out->n_new[0] = out->n_new[1] = out->n_new[2] = out->n_slip_rtk = out->n_slip_spp = 0;

  {  // GnssProcess :177-183: elevations at the current position
    double globalxyz[3] = {frame->pose[0] + data->base_xyz[0], frame->pose[1] + data->base_xyz[1], frame->pose[2] + data->base_xyz[2]};
    update_azel(globalxyz, data);
  }
  if (cfg.use_spp_correction) {  // :271-293
    for (i = 0; i < data->n_obs; i++) {
      d = data->obs + i;
      if (d->spp_p[0] != 0) {
        d->spp_p0[0] = d->spp_p[0];
        auto& lst = T->lists[SWGN_AMB_PCORR][d->sat * 2 + 0];
        if (lst.size()) {
          auto it = lst.end();
          it--;
          it->last_update_time = data->ros_time;
          if (it->continue_count > cfg.estimate_pcorrection_period) {
            d->spp_p0[0] = 0;
            d->spp_p[0] += it->value * cfg.lams[d->sys][0];
          }
        }
      } else {
        d->spp_p0[0] = 0;
      }
    }
  }
  for (i = 0; i < data->n_obs; i++) {  // :296-325
    d = data->obs + i;
    if (d->svh) continue;
    for (int f = 0; f < NF; f++) {
      auto last_recent = [&](int fam) -> PB* {
        auto& lst = T->lists[fam][d->sat * 2 + f];
        if (lst.size()) {
          auto it = lst.end();
          it--;
          if (data->ros_time - it->last_update_time < cfg.ambiguity_timeout) return &(*it);
        }
        return nullptr;
      };
      if (d->rtk_l[f] != 0) C.RTK_Npoint[i * NF + f] = last_recent(SWGN_AMB_RTK);
      if (d->spp_l[f] != 0) C.SPP_Npoint[i * NF + f] = last_recent(SWGN_AMB_SPP);
      if (d->spp_p0[f] != 0) C.PC_Npoint[i * NF + f] = last_recent(SWGN_AMB_PCORR);
    }
  }

  std::vector<double> error1_rtk(data->n_obs * 2, 0), error2_rtk[6], error1_spp(data->n_obs * 2, 0), error2_spp[6];
  double median_error_rtk[6] = {0}, median_error_spp[6] = {0};
  for (i = 0; i < data->n_obs; i++) {  // :346-377
    d = data->obs + i;
    if (d->svh) continue;
    const double* lam = cfg.lams[d->sys];
    const int sys = d->sys;
    for (int f = 0; f < NF; f++) {
      if (d->el < cfg.azelmin) d->rtk_l[f] = d->spp_l[f] = d->spp_p0[f] = 0;
      auto phase_residual = [&](double L, double Nvalue, double clock) {
        // RTKCarrierPhaseFactor(sat, L * lam, lam, el, 0, 0, base, use_istd = false, sys, f): weight 1
        double rec[SWGN_GNSS_STRIDE];
        gnss_record(rec, d, data->base_xyz, L * lam[f], lam[f], 1.0, false);
        std::unique_ptr<CostFunction> factor(make_gnss_factor(SWGN_GNSS_RTK_CARRIER, rec));
        double residuals;
        const double* parameter_blocks[3] = {frame->pose, &Nvalue, &clock};
        factor->Evaluate(parameter_blocks, &residuals, 0);
        return residuals;
      };
      if (PB* N = C.RTK_Npoint[i * NF + f]) {
        const double residuals = phase_residual(d->rtk_l[f], N->value, frame->gnss_dt[sys * 2 + f]);
        error1_rtk[i * 2 + f] = residuals;
        if (N->SLIP_COUNT == d->rtk_slip_count[f]) error2_rtk[d->sys * 2 + f].push_back(residuals);
      }
      if (PB* N = C.SPP_Npoint[i * NF + f]) {
        const double residuals = phase_residual(d->spp_l[f], N->value, frame->gnss_dt[6 + d->sys * 2 + 0]);
        error1_spp[i * 2 + f] = residuals;
        if (N->SLIP_COUNT == d->spp_slip_count[f]) error2_spp[d->sys * 2 + f].push_back(residuals);
      }
    }
  }
  for (int s = 0; s < 6; s++) {  // :378-390
    if (error2_rtk[s].size()) {
      std::sort(error2_rtk[s].begin(), error2_rtk[s].end());
      median_error_rtk[s] = error2_rtk[s][error2_rtk[s].size() / 2];
    }
    if (error2_spp[s].size()) {
      std::sort(error2_spp[s].begin(), error2_spp[s].end());
      median_error_spp[s] = error2_spp[s][error2_spp[s].size() / 2];
    }
  }
  for (i = 0; i < data->n_obs; i++) {  // :393-500
    d = data->obs + i;
    if (d->svh) continue;
    const double* lam = cfg.lams[d->sys];
    const int sys = d->sys;
    for (int f = 0; f < NF; f++) {
      bool condition3 = false, condition4 = false;
      PB*& RN = C.RTK_Npoint[i * NF + f];
      PB*& SN = C.SPP_Npoint[i * NF + f];
      PB*& PN = C.PC_Npoint[i * NF + f];
      if (d->rtk_l[f] != 0) {
        if (cfg.use_imu && cfg.use_rtk && frame->nonlinear && frame->rover_count > 1 && RN && RN->SLIP_COUNT == d->rtk_slip_count[f]) {
          double residuals = error1_rtk[i * 2 + f];
          if (std::fabs(residuals - median_error_rtk[sys * 2 + f]) > lam[f] * cfg.slip_fraction_rtk) {
            condition3 = true;
            out->n_slip_rtk++;
          }
        }
      }
      if (d->spp_l[f] != 0) {
        if (cfg.use_imu && cfg.use_spp_phase && frame->nonlinear && frame->rover_count > 1 && SN && SN->SLIP_COUNT == d->spp_slip_count[f]) {
          double residuals = error1_spp[i * 2 + f];
          if (std::abs((d->spp_l[f] + SN->value) * lam[f] - d->spp_p[f]) * std::sin(d->el) * std::sin(d->el) > 10) condition4 = true;
          if (std::fabs(residuals - median_error_spp[sys * 2 + f]) > lam[f]) condition4 = true;
          if (condition4) out->n_slip_spp++;
        }
      }
      if (d->rtk_l[f] != 0) {
        if ((!RN) || (RN->SLIP_COUNT != d->rtk_slip_count[f]) || condition3 || frame->not_fix_count > cfg.phase_all_reset_count) {
          RN = T->push(SWGN_AMB_RTK, d->sat, d->sys, f);
          RN->SLIP_COUNT = d->rtk_slip_count[f];
          RN->half_flag = d->half_flag[f];
          out->n_new[SWGN_AMB_RTK]++;
        }
        if (RN) RN->last_update_time = data->ros_time;
      }
      if (d->spp_l[f] != 0) {
        if ((!SN) || (SN->SLIP_COUNT != d->spp_slip_count[f]) || condition3 || condition4) {
          SN = T->push(SWGN_AMB_SPP, d->sat, d->sys, f);
          SN->SLIP_COUNT = d->spp_slip_count[f];
          SN->half_flag = d->half_flag[f];
          out->n_new[SWGN_AMB_SPP]++;
        }
        if (SN) SN->last_update_time = data->ros_time;
      }
      if (d->spp_p0[f] != 0) {
        if (!PN) {
          PN = T->push(SWGN_AMB_PCORR, d->sat, d->sys, f);
          out->n_new[SWGN_AMB_PCORR]++;
        }
        if (PN) PN->last_update_time = data->ros_time;
      }
      if (RN) RN->continue_count++;
      if (SN) SN->continue_count++;
      if (PN) PN->continue_count++;
    }
  }
  for (i = 0; i < data->n_obs; i++)
    for (int f = 0; f < NF; f++) {
      d = data->obs + i;
      d->rtk_n[f] = C.RTK_Npoint[i * NF + f] ? C.RTK_Npoint[i * NF + f]->handle : -1;
      d->spp_n[f] = C.SPP_Npoint[i * NF + f] ? C.SPP_Npoint[i * NF + f]->handle : -1;
      d->pcorr_n[f] = C.PC_Npoint[i * NF + f] ? C.PC_Npoint[i * NF + f]->handle : -1;
    }

  // ---- parameter storage (Vector2Double) and RemainPoint, :504-521 -----------------------------------------
  layout_store(C, data, frame);
  std::set<double*> RemainPoint{C.para_pose(), C.para_speed_bias(), C.blackvalue()};
  for (size_t a = 0; a < C.amb_order.size(); ++a) RemainPoint.insert(&C.store[30 + a]);

  // PhaseBiasSaveAndReset: all phase biases at 0 while the epoch is linearised (:523-531)
  MargInfo marg;
  AddGnssResidual(MargeExcludeMode, RemainPoint, &marg, nullptr, C);
  marg.marginalize();
  marg.getParameterBlocks();
  out->n_factors = (int)marg.factors.size();
  out->n_keep = (int)marg.keep_block_addr.size();
  out->n = marg.n;
  if (out->n_keep > out->cap_keep || out->n > out->cap_n) return -1;
  {
    int xo = 0;
    for (int k = 0; k < out->n_keep; ++k) {
      double* addr = marg.keep_block_addr[k];
      const long off = addr - C.store.data();
      int kind, handle = -1;
      if (off == 0) kind = SWGN_KEEP_POSE;
      else if (off == 7) kind = SWGN_KEEP_SPEED_BIAS;
      else if (off == 16) kind = SWGN_KEEP_BLACK;
      else {
        kind = SWGN_KEEP_AMB_RTK + C.amb_order[off - 30].first;
        handle = C.amb_order[off - 30].second->handle;
      }
      out->keep_kind[k] = kind;
      out->keep_handle[k] = handle;
      out->keep_idx[k] = marg.keep_block_idx[k] - marg.m;
      for (double v : marg.keep_block_data[k]) out->x0[xo++] = v;
    }
    for (int r = 0; r < marg.n; ++r) {
      for (int c = 0; c < marg.n; ++c) out->J0[(size_t)r * marg.n + c] = marg.linearized_jacobians(r, c);
      out->r0[r] = marg.linearized_residuals[r];
    }
  }
  // PhaseBiasRestore
  for (size_t a = 0; a < C.amb_order.size(); ++a) C.store[30 + a] = C.amb_order[a].second->value;

  std::memset(&out->init_summary, 0, sizeof(out->init_summary));
  if (cfg.use_spp_phase || cfg.use_rtk) {  // :532-571
    std::vector<Added> problem;
    AddGnssResidual(NormalMode, std::set<double*>{}, nullptr, &problem, C);
    // the ceres::Problem holds the blocks the residual blocks mention; flatten it into a swgn_graph
    std::map<double*, int> block_of;
    std::vector<double*> block_ptr;
    for (const Added& a : problem)
      for (double* p : a.parameter_blocks) block_of[p] = 0;
    for (auto& it : block_of) {
      it.second = (int)block_ptr.size();
      block_ptr.push_back(it.first);
    }
    const int nb = (int)block_ptr.size();
    std::vector<int32_t> size(nb, 1), manifold(nb, SWGN_MANIFOLD_EUCLIDEAN), konst(nb, 0), group(nb, 1), offset(nb, 0);
    std::vector<double> state;
    for (int b = 0; b < nb; ++b) {
      const long off = block_ptr[b] - C.store.data();
      if (off == 0) size[b] = 7, manifold[b] = SWGN_MANIFOLD_POSE, konst[b] = 1;  // SetParameterBlockConstant(para_pose[last])
      else if (off == 7) size[b] = 9, konst[b] = 1;
      else if (off >= 17 && off < 30) group[b] = 0;  // Ceres picks an independent set itself; any exact ordering gives the same step
      else if (off >= 30 && C.amb_order[off - 30].second->continue_count > cfg.init_constant_after) konst[b] = 1;
      offset[b] = (int32_t)state.size();
      state.insert(state.end(), block_ptr[b], block_ptr[b] + size[b]);
    }
    std::vector<int32_t> gkind, gblocks, ublock;
    std::vector<double> gdata, uistd;
    for (const Added& a : problem) {
      if (a.gnss_kind < 0) {
        ublock.push_back(block_of[a.parameter_blocks[0]]);
        uistd.push_back(1.0);
        continue;
      }
      gkind.push_back(a.gnss_kind);
      for (int k = 0; k < 3; ++k) gblocks.push_back(k < (int)a.parameter_blocks.size() ? block_of[a.parameter_blocks[k]] : -1);
      gdata.insert(gdata.end(), a.record.begin(), a.record.end());
    }
    if (!gkind.empty()) {
      swgn_graph g;
      std::memset(&g, 0, sizeof(g));
      g.n_blocks = nb;
      g.block_size = size.data();
      g.block_manifold = manifold.data();
      g.block_const = konst.data();
      g.block_group = group.data();
      g.block_offset = offset.data();
      g.n_state = (int32_t)state.size();
      g.state = state.data();
      g.proj_sqrt_info[0] = g.proj_sqrt_info[3] = 1.0;
      g.n_gnss = (int32_t)gkind.size();
      g.gnss_kind = gkind.data();
      g.gnss_blocks = gblocks.data();
      g.gnss_data = gdata.data();
      g.n_unit = (int32_t)ublock.size();
      g.unit_block = ublock.data();
      g.unit_istd = uistd.data();
      swgn_options opt;  // ceres::Solver::Options defaults + :563-567
      std::memset(&opt, 0, sizeof(opt));
      opt.max_num_iterations = cfg.init_max_iterations;
      opt.max_num_consecutive_invalid_steps = 5;
      opt.initial_trust_region_radius = opt.max_trust_region_radius = cfg.init_radius;
      opt.min_trust_region_radius = 1e-32;
      opt.min_relative_decrease = 1e-3;
      opt.min_lm_diagonal = 1e-6;
      opt.max_lm_diagonal = 1e32;
      opt.function_tolerance = 1e-6;
      opt.gradient_tolerance = 1e-10;
      opt.parameter_tolerance = 1e-8;
      opt.dogleg_min_mu = 1e-12;
      opt.is_optimize = 1;
      opt.trust_region_strategy = SWGN_LEVENBERG_MARQUARDT;
      opt.jacobi_scaling = 1;
      Solver s;
      if (!s.Build(&g, &opt) || !s.Preprocess()) return -2;
      s.Minimize(&out->init_summary);
      for (int b = 0; b < nb; ++b) std::copy(s.state.begin() + offset[b], s.state.begin() + offset[b] + size[b], block_ptr[b]);
    }
  }
  // Double2Vector for what this function owns
  frame->blackvalue = *C.blackvalue();
  std::copy(C.para_gnss_dt(), C.para_gnss_dt() + SWGN_GNSS_NCLK, frame->gnss_dt);
  for (size_t a = 0; a < C.amb_order.size(); ++a) C.amb_order[a].second->value = C.store[30 + a];
  return 0;
}

// The residual blocks AddGnssResidual builds for an epoch that has been preprocessed already (ambiguity handles in its
// observations), ambiguities at 0 as during the linearisation: kind (-1 = InitialBlackFactor, record[0] = istd), up to three
// offsets into the parameter storage per factor (-1 pad) and the record; store_out receives the storage
// (pose 0 | speed-bias 7 | blackvalue 16 | gnss_dt 17 | ambiguities 30..).  Returns the number of factors.
int gnss_epoch_factors(GnssTracker* T, swgn_epoch* data, swgn_gnss_frame* frame, int cap, int32_t* kind, int32_t* store_off,
                       double* records, double* store_out, int32_t* n_store) {
  const int NF = SWGN_NFREQ;
  EpochContext C;
  C.T = T;
  C.rover = data;
  C.frame = frame;
  C.RTK_Npoint.assign(data->n_obs * NF, nullptr);
  C.SPP_Npoint.assign(data->n_obs * NF, nullptr);
  C.PC_Npoint.assign(data->n_obs * NF, nullptr);
  for (int i = 0; i < data->n_obs; ++i)
    for (int f = 0; f < NF; ++f) {
      const swgn_obs& d = data->obs[i];
      if (d.rtk_n[f] >= 0) C.RTK_Npoint[i * NF + f] = T->by_handle[SWGN_AMB_RTK][d.rtk_n[f]];
      if (d.spp_n[f] >= 0) C.SPP_Npoint[i * NF + f] = T->by_handle[SWGN_AMB_SPP][d.spp_n[f]];
      if (d.pcorr_n[f] >= 0) C.PC_Npoint[i * NF + f] = T->by_handle[SWGN_AMB_PCORR][d.pcorr_n[f]];
    }
  layout_store(C, data, frame);
  std::vector<Added> problem;
  AddGnssResidual(NormalMode, std::set<double*>{}, nullptr, &problem, C);
  if ((int)problem.size() > cap) return -1;
  for (size_t i = 0; i < problem.size(); ++i) {
    kind[i] = problem[i].gnss_kind;
    for (int k = 0; k < 3; ++k)
      store_off[3 * i + k] = k < (int)problem[i].parameter_blocks.size() ? (int32_t)(problem[i].parameter_blocks[k] - C.store.data()) : -1;
    double* rec = records + (size_t)SWGN_GNSS_STRIDE * i;
    std::fill(rec, rec + SWGN_GNSS_STRIDE, 0.0);
    if (problem[i].gnss_kind < 0) rec[0] = 1.0;
    else std::copy(problem[i].record.begin(), problem[i].record.end(), rec);
  }
  std::copy(C.store.begin(), C.store.end(), store_out);
  *n_store = (int32_t)C.store.size();
  return (int)problem.size();
}

// The prior rebuild after FIX_CONTINUE_THRESHOLD accepted fixes, RVI/swf/swf_lambda.cpp:249-355: marginalization_info2 =
// { MarginalizationFactor(last_marg_info) over its keep blocks, FixedIntegerFactor(0, istd) on (tf[sys*2+f], reference
// ambiguity) once per system / frequency, FixedIntegerFactor(round(F), istd) on (tf[sys*2+f], ambiguity) per fixed double
// difference }, drop set = the tf dummies, marginalize(false, true), getParameterBlocks().
int fixed_integer_prior(const swgn_fixed_integer_job* J) {
  int nx = 0;
  std::vector<int> sizes(J->keep_size, J->keep_size + J->n_keep), idx(J->keep_idx, J->keep_idx + J->n_keep), xoff;
  for (int s : sizes) {
    xoff.push_back(nx);
    nx += s;
  }
  std::vector<double> store(J->x, J->x + nx);  // keep blocks, then double tf[6] = {0} (:247)
  store.resize(nx + 6, 0.0);
  double* tf = store.data() + nx;
  bool tfb[6] = {false};
  MargInfo marginalization_info2;
  {
    ResidualInfo info;
    info.cost.reset(make_prior_factor(J->n, sizes, idx, J->x0, J->J0, J->r0));
    for (int k = 0; k < J->n_keep; ++k) info.parameter_blocks.push_back(store.data() + xoff[k]);
    marginalization_info2.addResidualBlockInfo(info);
  }
  auto add_fixed = [&](double* dummy, double* point, double n21) {
    double rec[SWGN_GNSS_STRIDE] = {0};
    rec[SWGN_GNSS_MEAS] = n21;
    rec[SWGN_GNSS_WEIGHT] = J->istd;
    ResidualInfo info;
    info.cost.reset(make_gnss_factor(SWGN_GNSS_FIXED_INTEGER, rec));
    info.parameter_blocks = {dummy, point};
    info.drop_set = {0};
    marginalization_info2.addResidualBlockInfo(info);
  };
  for (int i = 0; i < J->n_dd; i++) {
    double* ppoint = store.data() + xoff[J->dd_keep[2 * i]];
    double* npoint = store.data() + xoff[J->dd_keep[2 * i + 1]];
    const int sf = J->dd_sysfreq[i];
    if (tfb[sf] == false) {
      add_fixed(&tf[sf], npoint, 0.0);
      tfb[sf] = true;
    }
    add_fixed(&tf[sf], ppoint, J->F[i]);
  }
  marginalization_info2.marginalize();
  marginalization_info2.getParameterBlocks();
  if (marginalization_info2.n != J->n) return -1;
  for (int r = 0; r < J->n; ++r) {
    for (int c = 0; c < J->n; ++c) J->J0_out[(size_t)r * J->n + c] = marginalization_info2.linearized_jacobians(r, c);
    J->r0_out[r] = marginalization_info2.linearized_residuals[r];
  }
  return 0;
}

// MarginalizationInfo over the factors of a whole swgn_graph with an arbitrary drop set (what MargFrames builds through
// AddAllResidual(MargeIncludeMode, ...), swf.cpp:329-341, swf_core.cpp:372-390): addResidualBlockInfo per factor,
// marginalize(false, true), getParameterBlocks.  Keep blocks come out in block-index order (= address order of the state).
int marginalize_graph(const swgn_graph* g, const uint8_t* drop, int cap_keep, int cap_n, int32_t* n_keep, int32_t* n_out, int32_t* m_out,
                      int32_t* keep_block, int32_t* keep_idx, double* J0, double* r0) {
  swgn_options opt;
  std::memset(&opt, 0, sizeof(opt));
  Solver s;
  if (!s.Build(g, &opt)) return -2;
  MargInfo info;
  for (auto& rb : s.residual_blocks) {
    ResidualInfo ri;
    ri.cost = std::shared_ptr<CostFunction>(rb->cost.get(), [](CostFunction*) {});  // owned by the Solver
    ri.cauchy_a = rb->cauchy_a;
    for (size_t k = 0; k < rb->blocks.size(); ++k) {
      ri.parameter_blocks.push_back(rb->blocks[k]->user_state);
      if (drop[rb->blocks[k]->graph_index]) ri.drop_set.push_back((int)k);
    }
    info.addResidualBlockInfo(ri);
  }
  info.marginalize();
  info.getParameterBlocks();
  *n_out = info.n;
  *m_out = info.m;
  *n_keep = (int)info.keep_block_addr.size();
  if (*n_keep > cap_keep || info.n > cap_n) return -1;
  for (int k = 0; k < *n_keep; ++k) {
    keep_block[k] = -1;
    for (auto& b : s.blocks)
      if (b.user_state == info.keep_block_addr[k]) keep_block[k] = b.graph_index;
    keep_idx[k] = info.keep_block_idx[k] - info.m;
  }
  for (int r = 0; r < info.n; ++r) {
    for (int c = 0; c < info.n; ++c) J0[(size_t)r * info.n + c] = info.linearized_jacobians(r, c);
    r0[r] = info.linearized_residuals[r];
  }
  return 0;
}

}  // namespace oracle

using namespace oracle;
extern "C" {
int oracle_marginalize_graph(const swgn_graph* g, const uint8_t* drop, int cap_keep, int cap_n, int32_t* n_keep, int32_t* n_out,
                             int32_t* m_out, int32_t* keep_block, int32_t* keep_idx, double* J0, double* r0) {
  return marginalize_graph(g, drop, cap_keep, cap_n, n_keep, n_out, m_out, keep_block, keep_idx, J0, r0);
}
int oracle_fixed_integer_prior(const swgn_fixed_integer_job* job) { return fixed_integer_prior(job); }
int oracle_gnss_epoch_factors(void* t, swgn_epoch* e, swgn_gnss_frame* f, int cap, int32_t* kind, int32_t* store_off, double* records,
                              double* store_out, int32_t* n_store) {
  return gnss_epoch_factors((GnssTracker*)t, e, f, cap, kind, store_off, records, store_out, n_store);
}
void* oracle_gnss_tracker_new(const swgn_gnss_config* cfg) {
  GnssTracker* t = new GnssTracker();
  t->cfg = *cfg;
  return t;
}
void oracle_gnss_tracker_free(void* t) { delete (GnssTracker*)t; }
int oracle_gnss_tracker_count(void* t, int fam) { return (int)((GnssTracker*)t)->by_handle[fam].size(); }
int oracle_gnss_tracker_get(void* t, int fam, int handle, swgn_ambiguity* out) {
  GnssTracker* T = (GnssTracker*)t;
  if (handle < 0 || handle >= (int)T->by_handle[fam].size()) return 1;
  const PB* p = T->by_handle[fam][handle];
  out->value = p->value;
  out->last_update_time = p->last_update_time;
  out->continue_count = p->continue_count;
  out->slip_count = p->SLIP_COUNT;
  out->half_flag = p->half_flag;
  out->sys = p->sys;
  out->f = p->f;
  out->sat = p->sat;
  out->alive = 1;
  return 0;
}
void oracle_gnss_tracker_set_value(void* t, int fam, int handle, double v) { ((GnssTracker*)t)->by_handle[fam][handle]->value = v; }
int oracle_gnss_preprocess(void* t, swgn_epoch* e, swgn_gnss_frame* f, swgn_gnss_output* out) {
  return gnss_preprocess((GnssTracker*)t, e, f, out);
}
void oracle_update_azel(const double* globalxyz, swgn_epoch* e) { update_azel(globalxyz, e); }
}
