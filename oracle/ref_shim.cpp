// TEST INFRASTRUCTURE.  extern "C" trampolines onto the reference's own C++-mangled functions
// (compiled from /root/reference by build_ref.sh) so ctypes can call them.
#include "common_function.h"
#include "lambda.h"
extern "C" {
int ref_lambda(int n, int m, const double* a, const double* Q, double* F, double* s) {
  return lambda(n, m, a, Q, F, s);
}
int ref_matinv(double* A, int n) { return matinv(A, n); }
double ref_distance(const double* rr, const double* rs, double* e) { return distance(rr, rs, e); }
double ref_velecitydistance(const double* rr, const double* rs, const double* vr, const double* vs,
                            double* e) {
  return velecitydistance(rr, rs, vr, vs, e);
}
double ref_dot(const double* a, const double* b, int n) { return dot(a, b, n); }
void ref_xyz2enu(const double* pos, double* E) { xyz2enu(pos, E); }
void ref_ecef2pos(const double* r, double* pos) { ecef2pos(r, pos); }
// update_azel (common_function.cpp:394-408) on a mea_t filled from flat arrays; el is in/out
void ref_update_azel(const double* globalxyz, int n, const double* satpos3, const unsigned char* svh, double* el) {
  static mea_t m;
  double xyz[3] = {globalxyz[0], globalxyz[1], globalxyz[2]};
  m.obs_count = n;
  for (int i = 0; i < n; ++i) {
    for (int c = 0; c < 3; ++c) m.obs_data[i].satellite_pos[c] = satpos3[3 * i + c];
    m.obs_data[i].SVH = svh[i];
    m.obs_data[i].el = el[i];
  }
  update_azel(xyz, &m);
  for (int i = 0; i < n; ++i) el[i] = m.obs_data[i].el;
}
}

// ---- the reference's own factor classes (RVI/factor/*.cpp compiled where they lie, against oracle/ref_stubs' minimal
// Eigen stand-in and the repository's include/ceres/ headers) behind the same C signature as oracle_factor_eval ----
#include <memory>
#include <vector>

#include "../include/swgn.h"
#include "factor/gnss_factor.h"
#include "factor/imu_factor.h"
#include "factor/pose_local_parameterization.h"
#include "factor/projection_factor.h"
#include "parameter/parameters.h"

double varerr2(double el, double dt, double mea_var);  // RVI/factor/gnss_factor.cpp:98-103 (no header declares it)

// (the application globals the factor sources read -- Pbg, Rwgw, G, ACC_N ... -- are defined in oracle/ref_globals.cpp)

namespace {
void set_globals(const double* g /* Pbg3, gravity3, proj sqrt_info 4 */) {
  Pbg = Eigen::Vector3d(g[0], g[1], g[2]);
  Rwgw = Eigen::Matrix3d::Identity();
  G = Eigen::Vector3d(g[3], g[4], g[5]);  // the factors use Rwgw * G
  projection_factor::sqrt_info(0, 0) = g[6];
  projection_factor::sqrt_info(0, 1) = g[7];
  projection_factor::sqrt_info(1, 0) = g[8];
  projection_factor::sqrt_info(1, 1) = g[9];
}
void load_record(IntegrationBase& ib, const double* rec) {
  for (int i = 0; i < 3; ++i) {
    ib.delta_p(i) = rec[SWGN_IMU_DELTA_P + i];
    ib.delta_v(i) = rec[SWGN_IMU_DELTA_V + i];
    ib.linearized_ba(i) = rec[SWGN_IMU_LIN_BA + i];
    ib.linearized_bg(i) = rec[SWGN_IMU_LIN_BG + i];
    ib.gyri(i) = rec[SWGN_IMU_GYRI + i];
    ib.gyrj(i) = rec[SWGN_IMU_GYRJ + i];
  }
  ib.delta_q = Eigen::Quaterniond(rec[SWGN_IMU_DELTA_Q + 3], rec[SWGN_IMU_DELTA_Q], rec[SWGN_IMU_DELTA_Q + 1], rec[SWGN_IMU_DELTA_Q + 2]);
  ib.sum_dt = rec[SWGN_IMU_SUM_DT];
  for (int i = 0; i < 15; ++i)
    for (int j = 0; j < 15; ++j) {
      ib.jacobian(i, j) = rec[SWGN_IMU_JACOBIAN + i * 15 + j];
      ib.sqrt_info(i, j) = rec[SWGN_IMU_SQRT_INFO + i * 15 + j];
    }
  ib.covariance_update = false;  // get_sqrtinfo() returns sqrt_info as loaded
}
void store_record(IntegrationBase& ib, double* rec) {
  for (int k = 0; k < SWGN_IMU_STRIDE; ++k) rec[k] = 0.0;
  const Eigen::Matrix<double, 15, 15> si = ib.get_sqrtinfo();
  for (int i = 0; i < 3; ++i) {
    rec[SWGN_IMU_DELTA_P + i] = ib.delta_p(i);
    rec[SWGN_IMU_DELTA_V + i] = ib.delta_v(i);
    rec[SWGN_IMU_LIN_BA + i] = ib.linearized_ba(i);
    rec[SWGN_IMU_LIN_BG + i] = ib.linearized_bg(i);
    rec[SWGN_IMU_GYRI + i] = ib.gyri(i);
    rec[SWGN_IMU_GYRJ + i] = ib.gyrj(i);
  }
  rec[SWGN_IMU_DELTA_Q + 0] = ib.delta_q.x();
  rec[SWGN_IMU_DELTA_Q + 1] = ib.delta_q.y();
  rec[SWGN_IMU_DELTA_Q + 2] = ib.delta_q.z();
  rec[SWGN_IMU_DELTA_Q + 3] = ib.delta_q.w();
  rec[SWGN_IMU_SUM_DT] = ib.sum_dt;
  for (int i = 0; i < 15; ++i)
    for (int j = 0; j < 15; ++j) {
      rec[SWGN_IMU_JACOBIAN + i * 15 + j] = ib.jacobian(i, j);
      rec[SWGN_IMU_SQRT_INFO + i * 15 + j] = si(i, j);
    }
}
}  // namespace

extern "C" {
// kind 0 projection_factor (record = uv), 1 IMUFactor (record = SWGN_IMU_STRIDE), 2 GNSS (kind2 = SWGN_GNSS_*, record =
// SWGN_GNSS_STRIDE; the RTK factors take el / dt / var and weigh with their own varerr2: a carrier factor with var <= 0 is
// built with use_istd = false).  params: concatenated global blocks in factor order; jac_out: concatenated row-major
// global Jacobians (may be null).
int ref_factor_eval(int kind, int kind2, const double* globals, const double* record, const double* params, double* residuals,
                    double* jac_out) {
  set_globals(globals);
  std::unique_ptr<ceres::CostFunction> f;
  std::unique_ptr<IntegrationBase> ib;
  double sat[3], satv[3], base[3], xyzt[3] = {0, 0, 0};
  if (kind == 0) {
    f.reset(new projection_factor(Eigen::Vector3d(record[0], record[1], 1.0)));
  } else if (kind == 1) {
    ib.reset(new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero()));
    load_record(*ib, record);
    f.reset(new IMUFactor(ib.get()));
  } else if (kind == 2) {
    for (int i = 0; i < 3; ++i) {
      sat[i] = record[SWGN_GNSS_SAT_POS + i];
      satv[i] = record[SWGN_GNSS_SAT_VEL + i];
      base[i] = record[SWGN_GNSS_BASE_POS + i];
    }
    const double meas = record[SWGN_GNSS_MEAS], lam = record[SWGN_GNSS_LAM], wgt = record[SWGN_GNSS_WEIGHT];
    const double el = record[SWGN_GNSS_EL], dt = record[SWGN_GNSS_DT], var = record[SWGN_GNSS_VAR];
    switch (kind2) {
      case SWGN_GNSS_SPP_PSEUDORANGE: f.reset(new SppPseudorangeFactor(sat, meas, wgt, base)); break;
      case SWGN_GNSS_SPP_CARRIER: f.reset(new SppCarrierPhaseFactor(sat, meas, wgt, base, lam)); break;
      case SWGN_GNSS_RTK_CARRIER: f.reset(new RTKCarrierPhaseFactor(sat, meas, lam, el, dt, var, base, var > 0.0, 0, 0)); break;
      case SWGN_GNSS_RTK_PSEUDORANGE: f.reset(new RTKPseudorangeFactor(sat, meas, el, dt, var, base)); break;
      case SWGN_GNSS_DOPPLER: f.reset(new SppDopplerFactor(satv, sat, xyzt, meas, wgt, base)); break;
      case SWGN_GNSS_FIXED_INTEGER: f.reset(new FixedIntegerFactor(meas, wgt)); break;
      default: return 2;
    }
  } else {
    return 2;
  }
  std::vector<const double*> p;
  std::vector<double*> J;
  const double* pp = params;
  double* jp = jac_out;
  for (int sz : f->parameter_block_sizes()) {
    p.push_back(pp);
    pp += sz;
    J.push_back(jp);
    if (jp) jp += (size_t)f->num_residuals() * sz;
  }
  return f->Evaluate(p.data(), residuals, jac_out ? J.data() : nullptr) ? 0 : 1;
}

// IntegrationBase: constructor + one push_back per further sample (RVI/factor/integration_base.cpp:5-142), then
// get_sqrtinfo(); samples = 7 doubles each (dt, acc, gyr), the first one is (acc_0, gyr_0).
int ref_preintegrate(int n_samples, const double* samples, const double* bias6, const double* noise4, double* record) {
  ACC_N = noise4[0];
  GYR_N = noise4[1];
  ACC_W = noise4[2];
  GYR_W = noise4[3];
  IntegrationBase ib(Eigen::Vector3d(samples[1], samples[2], samples[3]), Eigen::Vector3d(samples[4], samples[5], samples[6]),
                     Eigen::Vector3d(bias6[0], bias6[1], bias6[2]), Eigen::Vector3d(bias6[3], bias6[4], bias6[5]));
  for (int k = 1; k < n_samples; ++k) {
    const double* s = samples + 7 * k;
    ib.push_back(s[0], Eigen::Vector3d(s[1], s[2], s[3]), Eigen::Vector3d(s[4], s[5], s[6]));
  }
  store_record(ib, record);
  return 0;
}

void ref_pose_plus(const double* x, const double* delta, double* out) {
  PoseLocalParameterization p;
  static_cast<const ceres::LocalParameterization&>(p).Plus(x, delta, out);  // (the reference declares its overrides private)
}
double ref_varerr2(double el, double dt, double var) { return varerr2(el, dt, var); }
}

// ---- the reference's own IMUGNSSBase / IMUGNSSFactor (RVI/factor/gnss_imu_factor.cpp compiled where it lies, against the
// Eigen stand-in and this repository's include/ceres/{small_blas,invert_psd_matrix}.h) ------------------------------------
#include "factor/gnss_imu_factor.h"

namespace {
struct RefChain {
  int m = 0, k = 0;
  std::vector<double> pose0, sb0;                       // para_pose0 / para_speed_bias0 (never read by Evaluate)
  std::vector<double> hidden, hidden_lin;               // m x 16: pose 7 | speed-bias 9
  std::vector<double> dummyN;                           // gnss_phase_biases[i] targets
  std::vector<std::unique_ptr<IntegrationBase>> pre;    // m + 1
  IMUGNSSBase* base = nullptr;
  std::unique_ptr<IMUGNSSFactor> factor;
  ~RefChain() {
    factor.reset();
    delete base;  // deletes its IMUFactor objects, not the pre-integrations
  }
};
}  // namespace

extern "C" {
// Arrays exactly as the chain_* fields of swgn_graph (include/swgn.h): frames m x SWGN_CHAIN_FRAME_STRIDE, frameN m x 15 x k,
// chainN k x k + k, imu (m + 1) x SWGN_IMU_STRIDE.  The members of IMUGNSSBase are filled directly, the way AddMargInfo /
// SetLastImuFactor leave them (gnss_imu_factor.cpp:96-117,245-352), then Init().
void* ref_chain_create(const double* globals, int m, int k, const double* frames, const double* frameN, const double* chainN,
                       const double* imu) {
  set_globals(globals);
  RefChain* C = new RefChain();
  C->m = m;
  C->k = k;
  C->pose0.assign(7, 0.0);
  C->sb0.assign(9, 0.0);
  C->hidden.assign((size_t)16 * m, 0.0);
  C->hidden_lin.assign((size_t)16 * m, 0.0);
  C->dummyN.assign(std::max(k, 1), 0.0);
  IMUGNSSBase* B = new IMUGNSSBase(C->pose0.data(), C->sb0.data(), nullptr);
  C->base = B;
  for (int i = 0; i <= m; ++i) {
    C->pre.emplace_back(new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero()));
    load_record(*C->pre.back(), imu + (size_t)SWGN_IMU_STRIDE * i);
  }
  for (int i = 0; i < m; ++i) {
    const double* f = frames + (size_t)SWGN_CHAIN_FRAME_STRIDE * i;
    for (int q = 0; q < 7; ++q) {
      C->hidden[16 * i + q] = f[SWGN_CHAIN_POSE + q];
      C->hidden_lin[16 * i + q] = f[SWGN_CHAIN_POSE_LIN + q];
    }
    for (int q = 0; q < 9; ++q) {
      C->hidden[16 * i + 7 + q] = f[SWGN_CHAIN_SB + q];
      C->hidden_lin[16 * i + 7 + q] = f[SWGN_CHAIN_SB_LIN + q];
    }
    B->gnss_poses.push_back(&C->hidden[16 * i]);
    B->gnss_speed_bias.push_back(&C->hidden[16 * i + 7]);
    B->gnss_poses_lin.push_back(&C->hidden_lin[16 * i]);
    B->gnss_speed_bias_lin.push_back(&C->hidden_lin[16 * i + 7]);
    Eigen::Matrix<double, 15, 15, Eigen::RowMajor> H;
    Eigen::Matrix<double, 15, 1, Eigen::ColMajor> rhs;
    Eigen::Matrix<double, 15, Eigen::Dynamic, Eigen::RowMajor> HN(15, k);
    for (int a = 0; a < 15; ++a) {
      rhs(a) = f[SWGN_CHAIN_RHS + a];
      for (int b = 0; b < 15; ++b) H(a, b) = f[SWGN_CHAIN_HESSIAN + 15 * a + b];
      for (int b = 0; b < k; ++b) HN(a, b) = frameN[((size_t)i * 15 + a) * k + b];
    }
    B->pose_hessians.push_back(H);
    B->pose_rhses.push_back(rhs);
    B->pose_phase_biases_hessians.push_back(HN);
    B->imu_factors.push_back(new IMUFactor(C->pre[i].get()));
  }
  B->last_imu_factor = new IMUFactor(C->pre[m].get());
  B->pose1_pose2_hessians.setZero();
  B->phase_biases_hessians.resize(k, k);
  B->phase_biases_rhs.resize(k);
  for (int a = 0; a < k; ++a) {
    B->phase_biases_rhs(a) = chainN[(size_t)k * k + a];
    for (int b = 0; b < k; ++b) B->phase_biases_hessians(a, b) = chainN[(size_t)a * k + b];
    B->gnss_phase_biases.push_back(&C->dummyN[a]);
  }
  B->gnss_Index = m;
  B->Init();
  C->factor.reset(new IMUGNSSFactor(B));
  return C;
}
// params: pose_i 7 | sb_i 9 | pose_j 7 | sb_j 9 | N k.  jac (may be null = cost-only evaluation): the 4 + k row-major
// (30 + k) x global-size Jacobians back to back.
int ref_chain_evaluate(void* h, const double* params, double* residuals, double* jac) {
  RefChain* C = (RefChain*)h;
  const int n = 30 + C->k;
  std::vector<const double*> p = {params, params + 7, params + 16, params + 23};
  for (int a = 0; a < C->k; ++a) p.push_back(params + 32 + a);
  std::vector<double*> J;
  if (jac) {
    double* q = jac;
    const int sizes[4] = {7, 9, 7, 9};
    for (int b = 0; b < 4; ++b) {
      J.push_back(q);
      q += (size_t)n * sizes[b];
    }
    for (int a = 0; a < C->k; ++a) {
      J.push_back(q);
      q += n;
    }
  }
  return C->factor->Evaluate(p.data(), residuals, jac ? J.data() : nullptr) ? 0 : 1;
}
void ref_chain_frames(void* h, double* out) {
  RefChain* C = (RefChain*)h;
  std::copy(C->hidden.begin(), C->hidden.end(), out);
}
void ref_chain_destroy(void* h) { delete (RefChain*)h; }
}

// IMUGNSSBase::AddMargInfo (gnss_imu_factor.cpp:245-352) executed on real MarginalizationInfo objects: one call per
// epoch, keep blocks identified across epochs by keep_id (a shared id = the same address, e.g. an ambiguity or the
// blackvalue seen by several epochs).  Per epoch e: n_keep[e] blocks (sizes / first columns / ids / x0 concatenated over
// the epochs), n[e], A (n x n row-major) and b concatenated.  Outputs: k, slot_of_id[max_ids] (position in
// gnss_phase_biases or -1), pose_hessians (E x 225), pose_rhses (E x 15), pose_N (E x 15 x k), NN (k x k), Nrhs (k).
extern "C" int ref_add_marg_info(int n_epochs, const int32_t* n_keep, const int32_t* keep_size, const int32_t* keep_idx,
                                 const int32_t* keep_id, const double* x0, const int32_t* n, const double* A, const double* b, int max_ids,
                                 int32_t* k_out, int32_t* slot_of_id, double* pose_hessians, double* pose_rhses, double* pose_N, double* NN,
                                 double* Nrhs) {
  std::vector<std::vector<double>> storage(max_ids, std::vector<double>(9, 0.0));  // user memory of every block
  std::vector<double> pose0(7, 0.0), sb0(9, 0.0);
  IMUGNSSBase B(pose0.data(), sb0.data(), nullptr);
  std::vector<std::unique_ptr<MarginalizationInfo>> infos;
  std::vector<std::unique_ptr<IntegrationBase>> pres;
  std::vector<std::vector<double>> lin;
  size_t ko = 0, xo = 0, ao = 0, bo = 0;
  for (int e = 0; e < n_epochs; ++e) {
    MarginalizationInfo* M = new MarginalizationInfo();
    infos.emplace_back(M);
    M->n = n[e];
    M->m = 0;
    for (int i = 0; i < n_keep[e]; ++i) {
      const int s = keep_size[ko + i];
      lin.emplace_back(x0 + xo, x0 + xo + s);
      xo += s;
    }
    for (int i = 0; i < n_keep[e]; ++i) {
      M->keep_block_size.push_back(keep_size[ko + i]);
      M->keep_block_idx.push_back(keep_idx[ko + i]);
      M->keep_block_addr.push_back(storage[keep_id[ko + i]].data());
      M->keep_block_data.push_back(lin[lin.size() - n_keep[e] + i].data());
    }
    M->A.resize(n[e], n[e]);
    M->b = Eigen::VectorXd(n[e]);
    for (int r = 0; r < n[e]; ++r) {
      M->b(r) = b[bo + r];
      for (int c = 0; c < n[e]; ++c) M->A(r, c) = A[ao + (size_t)r * n[e] + c];
    }
    pres.emplace_back(new IntegrationBase(Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero(), Eigen::Vector3d::Zero()));
    std::vector<double> dummy_pose(7, 0.0), dummy_sb(9, 0.0);
    B.AddMargInfo(M, pres.back().get(), dummy_pose.data(), dummy_sb.data());
    ko += n_keep[e];
    ao += (size_t)n[e] * n[e];
    bo += n[e];
  }
  const int k = (int)B.gnss_phase_biases.size();
  *k_out = k;
  for (int id = 0; id < max_ids; ++id) {
    slot_of_id[id] = -1;
    for (int q = 0; q < k; ++q)
      if (B.gnss_phase_biases[q] == storage[id].data()) slot_of_id[id] = q;
  }
  for (int e = 0; e < n_epochs; ++e) {
    for (int a = 0; a < 15; ++a) {
      pose_rhses[15 * e + a] = B.pose_rhses[e](a);
      for (int c = 0; c < 15; ++c) pose_hessians[225 * e + 15 * a + c] = B.pose_hessians[e](a, c);
      for (int c = 0; c < k; ++c) pose_N[((size_t)e * 15 + a) * k + c] = c < B.pose_phase_biases_hessians[e].cols() ? B.pose_phase_biases_hessians[e](a, c) : 0.0;
    }
  }
  for (int a = 0; a < k; ++a) {
    Nrhs[a] = B.phase_biases_rhs(a);
    for (int c = 0; c < k; ++c) NN[(size_t)a * k + c] = B.phase_biases_hessians(a, c);
  }
  // the marginalization infos stay owned here (IMUGNSSBase keeps pointers into them only through the vectors above)
  B.gnss_poses.clear();
  return 0;
}
