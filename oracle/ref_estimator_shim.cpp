// TEST INFRASTRUCTURE.  The reference's ESTIMATOR code for the per-epoch GNSS path -- RVI/swf/swf_gnss.cpp (GnssPreprocess,
// MyOrdering, UpdateSchur*, PhaseBias*), swf_core.cpp (AddGnssResidual, AddAllResidual) and swf_lambda.cpp (LambdaSearch),
// compiled UNMODIFIED where they lie -- executed on this repository's ceres:: shim, i.e. with every ceres::Solve inside
// them running on the device.  Built by oracle/build_ref.sh into oracle/_ref/libref_estimator.so.
//
// The estimator class has members and methods that live in translation units which need ROS and OpenCV (swf.cpp,
// swf_imu.cpp, swf_image.cpp, feature_*.cpp); those are not compiled.  What the executed code needs from them is defined
// here instead: the constructor (swf.cpp:13-34 + the allocations of ClearState :55-82), Vector2Double (swf.cpp:140-164,
// restated), empty constructors of the two feature classes, the application globals of parameters.cpp, and abort() bodies
// for five methods the executed paths never reach.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "swf/swf.h"
#include "../include/swgn_gnss.h"
#include "../rtk-visual-inertial-navigation_b200/shim/reference_gnss_binding.h"

// ---- application globals (RVI/parameter/parameters.cpp reads them from the yaml; Pbg, Rwgw, G, ACC_N .. and
// USE_GLOBAL_OPTIMIZATION / MAX_TRUST_REGION_RADIUS are in ref_globals.cpp) -------------------------------------------------
bool USE_IMAGE = false, USE_GNSS = true, USE_IMU = true, USE_RTK = true, USE_RTD = true, USE_DOPPLER = true, USE_SPP_PHASE = false;
bool USE_MAG_INIT_YAW = false, USE_MAG_CORRECT_YAW = false, USE_DIRECT_N_RESOLVE = false, USE_N_RESOLVE = true, USE_SPP_CORRECTION = false;
bool USE_STEREO = false;
Eigen::Vector3d ANCHOR_POINT;
std::vector<Eigen::Matrix3d> RIC;
std::vector<Eigen::Vector3d> TIC;
double USE_FEATURE = 0;
int CARRIER_PHASE_CONTINUE_THRESHOLD = 0, FIX_CONTINUE_THRESHOLD = 0, Phase_ALL_RESET_COUNT = 10;
int NUM_OF_CAM = 0, ESTIMATE_EXTRINSIC = 0, MAX_NUM_ITERATIONS = 8;

// ---- members defined in translation units that are not compiled ------------------------------------------------------
FeatureTracker::FeatureTracker() {}
FeatureManager::FeatureManager(Eigen::Matrix3d _Rs[]) : Rs(_Rs) {}
SWFOptimization::SWFOptimization() : f_manager{Rs} {
  for (int i = 0; i < FEATURE_WINDOW_SIZE + GNSS_WINDOW_SIZE + 1; i++) {  // ClearState, swf.cpp:56-72
    para_gnss_dt[i] = new double[13];
    para_pose[i] = new double[SIZE_POSE];
    para_speed_bias[i] = new double[SIZE_SPEEDBIAS];
    Rs[i].setIdentity();
    Ps[i].setZero();
    Vs[i].setZero();
    Bas[i].setZero();
    Bgs[i].setZero();
    for (int s = 0; s < 13; s++) para_gnss_dt[i][s] = 0;
    rovers[i] = nullptr;
    i2f[i] = g2f[i] = f2i[i] = f2g[i] = 0;
    frame_types[i] = ErroFrame;
  }
  solver_flag = Initial;
  last_marg_info = nullptr;
  rover_count = 0;
  image_count = 0;
  fix = false;
  prev_time = prev_time2 = -1;
  cur_time = 0;
  open_ex_estimation = 0;
  imu_initialize = false;
  init_gnss = true;
  pub_init = false;
  first_imu = true;
  rtk_fix = false;
  not_fix_count = 0;
  gnss_fix_solution_count = 0;
  my_options.linear_solver_type = ceres::DENSE_SCHUR;
  my_options.max_num_iterations = MAX_NUM_ITERATIONS;
  my_options.jacobi_scaling = 0;
  my_options.trust_region_strategy_type = ceres::DOGLEG;
  my_options.num_threads = 4;
  my_options.linear_solver_ordering.reset(new ceres::ParameterBlockOrdering());
}
void SWFOptimization::Vector2Double() {  // swf.cpp:140-164 (the feature loop of :177-180 has no effect)
  for (int i = 0; i < rover_count + image_count; i++) {
    para_pose[i][0] = Ps[i].x();
    para_pose[i][1] = Ps[i].y();
    para_pose[i][2] = Ps[i].z();
    Quaterniond q{Rs[i]};
    para_pose[i][3] = q.x();
    para_pose[i][4] = q.y();
    para_pose[i][5] = q.z();
    para_pose[i][6] = q.w();
    para_speed_bias[i][0] = Vs[i].x();
    para_speed_bias[i][1] = Vs[i].y();
    para_speed_bias[i][2] = Vs[i].z();
    para_speed_bias[i][3] = Bas[i].x();
    para_speed_bias[i][4] = Bas[i].y();
    para_speed_bias[i][5] = Bas[i].z();
    para_speed_bias[i][6] = Bgs[i].x();
    para_speed_bias[i][7] = Bgs[i].y();
    para_speed_bias[i][8] = Bgs[i].z();
  }
}
static void not_reached(const char* what) {
  std::fprintf(stderr, "libref_estimator.so: SWFOptimization::%s is not part of the executed path\n", what);
  std::abort();
}
void SWFOptimization::Double2Vector() { not_reached("Double2Vector"); }
void SWFOptimization::InitializePos(Eigen::Matrix3d&) { not_reached("InitializePos"); }
void SWFOptimization::ResetImuGnssFactor(int, MarginalizationInfo*) { not_reached("ResetImuGnssFactor"); }
void SWFOptimization::SlideWindowFrame(int, int, bool) { not_reached("SlideWindowFrame"); }
void SWFOptimization::UpdateVisualGnssIndex() { not_reached("UpdateVisualGnssIndex"); }
MarginalizationInfo* SWFOptimization::MargGNSSFrames(std::set<int>, IMUGNSSBase*) {
  not_reached("MargGNSSFrames");
  return nullptr;
}

// ---- the driver -----------------------------------------------------------------------------------------------------
namespace {
struct Estimator {
  SWFOptimization swf;
  std::vector<mea_t*> epochs;  // kept alive: the ambiguity lists are pointed at from inside
  ~Estimator() {
    for (mea_t* m : epochs) {
      delete m->marg_info_gnss;
      delete m;
    }
  }
};
void set_config(const swgn_gnss_config& c) {
  USE_IMU = c.use_imu;
  USE_RTK = c.use_rtk;
  USE_RTD = c.use_rtd;
  USE_SPP_PHASE = c.use_spp_phase;
  USE_SPP_CORRECTION = c.use_spp_correction;
  USE_DOPPLER = c.use_doppler;
  Phase_ALL_RESET_COUNT = c.phase_all_reset_count;
  for (int s = 0; s < 3; ++s)
    for (int f = 0; f < 2; ++f) lams[s][f] = c.lams[s][f];
}
int family_of(SWFOptimization& s, const double* p, int* sat2f, int* pos) {
  std::list<PBtype>* fam[3] = {s.rtk_phase_bias_variables, s.spp_phase_bias_variables, s.pseudorange_correction_variables};
  for (int f = 0; f < 3; ++f)
    for (int i = 0; i < MAXSATNUM * 2; ++i) {
      int k = 0;
      for (auto it = fam[f][i].begin(); it != fam[f][i].end(); ++it, ++k)
        if (&it->value == p) {
          *sat2f = i;
          *pos = k;
          return f;
        }
    }
  return -1;
}
}  // namespace

extern "C" {
void* ref_est_create(const swgn_gnss_config* cfg) {
  set_config(*cfg);
  return new Estimator();
}
void ref_est_destroy(void* h) { delete (Estimator*)h; }

// SWFOptimization::GnssPreprocess (swf_gnss.cpp:265-587) for one epoch, preceded by the update_azel call of GnssProcess
// (:177-183).  The estimator is put into the state GnssProcess leaves it in: the epoch is the newest of rover_count
// frames, its frame holds (pose, speed-bias), para_gnss_dt[0] and blackvalue carry the running estimates.
// Outputs: the epoch (el, masked measurements), frame (gnss_dt, blackvalue after the initialisation solve) and, for the
// epoch's marg_info_gnss: n, m, keep blocks in the reference's own order -- kind (SWGN_KEEP_*), for ambiguities
// sat * 2 + f and the position in that list, first column --, the linearised factor (n x n row-major, n) and its x0.
int ref_est_gnss_preprocess(void* h, const swgn_gnss_config* cfg, swgn_epoch* epoch, swgn_gnss_frame* frame, int cap_keep, int cap_n,
                            int32_t* n_out, int32_t* n_keep, int32_t* keep_kind, int32_t* keep_sat2f, int32_t* keep_pos, int32_t* keep_idx,
                            double* x0, double* J0, double* r0) {
  Estimator* E = (Estimator*)h;
  SWFOptimization& S = E->swf;
  set_config(*cfg);
  if (epoch->n_obs > MAXOBS) return -1;
  mea_t* rover = new mea_t();
  std::memset((void*)rover, 0, sizeof(mea_t));
  E->epochs.push_back(rover);
  rover->obs_count = epoch->n_obs;
  rover->ros_time = epoch->ros_time;
  rover->br_time_diff = epoch->br_time_diff;
  for (int c = 0; c < 3; ++c) rover->base_xyz[c] = epoch->base_xyz[c];
  for (int i = 0; i < epoch->n_obs; ++i) {
    ObsMea& d = rover->obs_data[i];
    const swgn_obs& o = epoch->obs[i];
    d.sat = o.sat;
    d.sys = o.sys;
    d.SVH = o.svh;
    for (int f = 0; f < NFREQ; ++f) {
      d.RTK_SLIP_COUNT[f] = o.rtk_slip_count[f];
      d.SPP_SLIP_COUNT[f] = o.spp_slip_count[f];
      d.half_flag[f] = o.half_flag[f];
      d.SPP_P[f] = o.spp_p[f];
      d.SPP_L[f] = o.spp_l[f];
      d.SPP_D[f] = o.spp_d[f];
      d.SPP_Lstd[f] = o.spp_lstd[f];
      d.SPP_Pstd[f] = o.spp_pstd[f];
      d.SPP_Dstd[f] = o.spp_dstd[f];
      d.RTK_P[f] = o.rtk_p[f];
      d.RTK_L[f] = o.rtk_l[f];
      d.RTK_Pstd[f] = o.rtk_pstd[f];
      d.RTK_Lstd[f] = o.rtk_lstd[f];
      d.SPP_P0[f] = o.spp_p0[f];
    }
    for (int c = 0; c < 3; ++c) {
      d.satellite_pos[c] = o.sat_pos[c];
      d.satellite_vel[c] = o.sat_vel[c];
    }
    d.el = o.el;
    d.sat_var = o.sat_var;
    d.ion_var = o.ion_var;
    d.trop_var = o.trop_var;
  }
  // estimator state as GnssProcess has it when it calls GnssPreprocess: two frames, the epoch is the newest one
  S.rover_count = 2;
  S.image_count = 0;
  S.g2f[0] = 0;
  S.g2f[1] = 1;
  S.frame_types[0] = S.frame_types[1] = SWFOptimization::GnssFrame;
  S.rovers[1] = rover;
  S.solver_flag = frame->nonlinear ? SWFOptimization::NonLinear : SWFOptimization::Initial;
  S.not_fix_count = frame->not_fix_count;
  S.rover_count_accumulate = frame->epochs_since_start + 1;  // rover_count_accumulate - rover_count + ir with ir = rover_count - 1
  S.blackvalue = frame->blackvalue;
  const Eigen::Quaterniond q(frame->pose[6], frame->pose[3], frame->pose[4], frame->pose[5]);
  for (int i = 0; i < 2; ++i) {
    S.Ps[i] = Eigen::Vector3d(frame->pose[0], frame->pose[1], frame->pose[2]);
    S.Rs[i] = q.toRotationMatrix();
    S.Vs[i] = Eigen::Vector3d(frame->speed_bias[0], frame->speed_bias[1], frame->speed_bias[2]);
    S.Bas[i] = Eigen::Vector3d(frame->speed_bias[3], frame->speed_bias[4], frame->speed_bias[5]);
    S.Bgs[i] = Eigen::Vector3d(frame->speed_bias[6], frame->speed_bias[7], frame->speed_bias[8]);
  }
  for (int s = 0; s < 13; ++s) S.para_gnss_dt[0][s] = frame->gnss_dt[s];
  {  // GnssProcess :177-183
    double globalxyz[3] = {S.Ps[1].x() + rover->base_xyz[0], S.Ps[1].y() + rover->base_xyz[1], S.Ps[1].z() + rover->base_xyz[2]};
    update_azel(globalxyz, rover);
  }
  S.GnssPreprocess(rover);

  // ---- read the results back ----
  for (int i = 0; i < epoch->n_obs; ++i) {
    const ObsMea& d = rover->obs_data[i];
    swgn_obs& o = epoch->obs[i];
    o.el = d.el;
    for (int f = 0; f < NFREQ; ++f) {
      o.rtk_l[f] = d.RTK_L[f];
      o.spp_l[f] = d.SPP_L[f];
      o.spp_p[f] = d.SPP_P[f];
      o.spp_p0[f] = d.SPP_P0[f];
    }
  }
  for (int s = 0; s < 13; ++s) frame->gnss_dt[s] = S.para_gnss_dt[0][s];
  frame->blackvalue = S.blackvalue;
  MarginalizationInfo* M = rover->marg_info_gnss;
  *n_out = M->n;
  *n_keep = (int)M->keep_block_addr.size();
  if (*n_keep > cap_keep || M->n > cap_n) return -2;
  int xo = 0;
  for (int k = 0; k < *n_keep; ++k) {
    const double* p = M->keep_block_addr[k];
    keep_idx[k] = M->keep_block_idx[k] - M->m;
    keep_sat2f[k] = keep_pos[k] = -1;
    if (p == S.para_pose[1]) keep_kind[k] = SWGN_KEEP_POSE;
    else if (p == S.para_speed_bias[1]) keep_kind[k] = SWGN_KEEP_SPEED_BIAS;
    else if (p == &S.blackvalue) keep_kind[k] = SWGN_KEEP_BLACK;
    else {
      const int fam = family_of(S, p, &keep_sat2f[k], &keep_pos[k]);
      if (fam < 0) return -3;
      keep_kind[k] = SWGN_KEEP_AMB_RTK + fam;
    }
    for (int c = 0; c < M->keep_block_size[k]; ++c) x0[xo++] = M->keep_block_data[k][c];
  }
  for (int r = 0; r < M->n; ++r) {
    for (int c = 0; c < M->n; ++c) J0[(size_t)r * M->n + c] = M->linearized_jacobians(r, c);
    r0[r] = M->linearized_residuals(r);
  }
  return 0;
}

// the estimator's ambiguity lists: entries of family fam (0 RTK, 1 SPP, 2 pseudorange correction) in list order
int ref_est_ambiguities(void* h, int fam, int cap, int32_t* sat2f, int32_t* pos, double* value, int32_t* continue_count, int32_t* slip_count,
                        double* last_update_time) {
  SWFOptimization& S = ((Estimator*)h)->swf;
  std::list<PBtype>* lists = fam == 0 ? S.rtk_phase_bias_variables : fam == 1 ? S.spp_phase_bias_variables : S.pseudorange_correction_variables;
  int n = 0;
  for (int i = 0; i < MAXSATNUM * 2; ++i) {
    int k = 0;
    for (auto it = lists[i].begin(); it != lists[i].end(); ++it, ++k, ++n) {
      if (n >= cap) return -1;
      sat2f[n] = i;
      pos[n] = k;
      value[n] = it->value;
      continue_count[n] = it->continue_count;
      slip_count[n] = it->SLIP_COUNT;
      last_update_time[n] = it->last_update_time;
    }
  }
  return n;
}
void ref_est_set_ambiguity(void* h, int fam, int sat2f, int pos, double value) {
  SWFOptimization& S = ((Estimator*)h)->swf;
  std::list<PBtype>* lists = fam == 0 ? S.rtk_phase_bias_variables : fam == 1 ? S.spp_phase_bias_variables : S.pseudorange_correction_variables;
  auto it = lists[sat2f].begin();
  std::advance(it, pos);
  it->value = value;
}

// SWFOptimization::LambdaSearch (swf_lambda.cpp:82-365, with FindReferenceSatellites :8-53 and lambda()) on the last
// n_window epochs this estimator preprocessed (they become rovers[0 .. n_window-1], oldest first).  The caller provides
// what UpdateSchurHessianOnly leaves behind -- A (n x n row-major information of the n ambiguities named by
// (amb_sat2f, amb_pos) in the RTK lists; b is not read by the decision) -- and last_marg_info as a prior over exactly
// those ambiguities (J0 n x n row-major, r0).  FIX_CONTINUE_THRESHOLD = 0: an accepted fix rebuilds the prior at once
// (:249-355).  Outputs: flags = {rtk_fix, fix, last_fix, not_fix_count, gnss_fix_solution_count, prior rebuilt}; when
// rebuilt, the new last_marg_info: keep_row[k] = row of A its k-th keep block is, its first column, J (n x n), r (n).
int ref_est_lambda_search(void* h, int n_window, int n, const int32_t* amb_sat2f, const int32_t* amb_pos, const double* A, const double* J0,
                          const double* r0, int32_t* flags, int32_t* keep_row, int32_t* keep_col, double* Jn, double* rn) {
  Estimator* E = (Estimator*)h;
  SWFOptimization& S = E->swf;
  if (n_window < 1 || n_window > (int)E->epochs.size()) return -1;
  FIX_CONTINUE_THRESHOLD = 0;
  USE_GLOBAL_OPTIMIZATION = false;
  S.rover_count = n_window;
  S.image_count = 0;
  for (int i = 0; i < n_window; ++i) {
    S.rovers[i] = E->epochs[E->epochs.size() - n_window + i];
    S.g2f[i] = i;
  }
  std::vector<double*> addr(n);
  for (int k = 0; k < n; ++k) {
    auto it = S.rtk_phase_bias_variables[amb_sat2f[k]].begin();
    std::advance(it, amb_pos[k]);
    addr[k] = &it->value;
  }
  S.A = Eigen::MatrixXd(n, n);
  S.b = Eigen::VectorXd(n);
  S.parameter_block_addr.clear();
  S.parameter_block_global_size.clear();
  for (int r = 0; r < n; ++r) {
    S.b(r) = 0.0;
    S.parameter_block_addr.push_back(addr[r]);
    S.parameter_block_global_size.push_back(1);
    for (int c = 0; c < n; ++c) S.A(r, c) = A[(size_t)r * n + c];
  }
  static std::vector<std::vector<double>> lin_keepalive;
  lin_keepalive.emplace_back(n, 0.0);
  std::vector<double>& lin = lin_keepalive.back();
  MarginalizationInfo* M = new MarginalizationInfo();
  M->n = n;
  M->m = 0;
  for (int k = 0; k < n; ++k) {
    M->keep_block_size.push_back(1);
    M->keep_block_idx.push_back(k);
    M->keep_block_data.push_back(&lin[k]);
    M->keep_block_addr.push_back(addr[k]);
  }
  M->linearized_jacobians.resize(n, n);
  M->linearized_residuals = Eigen::VectorXd(n);
  for (int r = 0; r < n; ++r) {
    M->linearized_residuals(r) = r0[r];
    for (int c = 0; c < n; ++c) M->linearized_jacobians(r, c) = J0[(size_t)r * n + c];
  }
  S.last_marg_info = M;
  S.rtk_fix = false;
  S.fix = false;
  S.LambdaSearch();
  flags[0] = S.rtk_fix;
  flags[1] = S.fix;
  flags[2] = S.last_fix;
  flags[3] = S.not_fix_count;
  flags[4] = S.gnss_fix_solution_count;
  flags[5] = S.last_marg_info != M;
  if (flags[5]) {
    MarginalizationInfo* N = S.last_marg_info;  // (the old one was deleted by LambdaSearch)
    if (N->n != n || (int)N->keep_block_addr.size() != n) return -2;
    for (int k = 0; k < n; ++k) {
      keep_row[k] = -1;
      for (int q = 0; q < n; ++q)
        if (N->keep_block_addr[k] == addr[q]) keep_row[k] = q;
      keep_col[k] = N->keep_block_idx[k] - N->m;
    }
    for (int r = 0; r < n; ++r) {
      for (int c = 0; c < n; ++c) Jn[(size_t)r * n + c] = N->linearized_jacobians(r, c);
      rn[r] = N->linearized_residuals(r);
    }
  }
  return 0;
}

// SWFOptimization::UpdateSchur / UpdateSchurHessianOnly (swf_gnss.cpp:25-94) applied to what the shim exports after a
// ceres::Solve of a synthetic window built from the reference's factor classes (shim/ceres_shim_refdemo.cpp, linked into
// this library): hessian_only = 0 -> export-mode solve, then UpdateSchur (A, b of the parameter_head blocks);
// hessian_only = 1 -> optimising solve, then UpdateSchurHessianOnly (A = L_nn L_nn').
}
extern "C" int swgn_ceres_refdemo_solve(int which, uint64_t window_id, int variant, int strategy, int host_factors, int device,
                                        double* state_out, double* cost_out, int* steps_out, char* message, int message_len);
extern "C" void swgn_ceres_refdemo_set_hooks(int is_optimize, void (*after_solve)(ceres::Problem*));
namespace {
SWFOptimization* g_us_swf = nullptr;
int g_us_hessian_only = 0;
void update_schur_hook(ceres::Problem* problem) {
  if (g_us_hessian_only) g_us_swf->UpdateSchurHessianOnly(*problem);
  else g_us_swf->UpdateSchur(*problem);
}
}  // namespace
extern "C" {
int ref_est_update_schur(int which, uint64_t window_id, int hessian_only, int cap_n, int32_t* n_out, double* A_out, double* b_out) {
  static SWFOptimization* swf = new SWFOptimization();
  g_us_swf = swf;
  g_us_hessian_only = hessian_only;
  swgn_ceres_refdemo_set_hooks(hessian_only ? 1 : 0, &update_schur_hook);
  char msg[256];
  const int rc = swgn_ceres_refdemo_solve(which, window_id, 0, 0, 0, 0, nullptr, nullptr, nullptr, msg, sizeof(msg));
  swgn_ceres_refdemo_set_hooks(1, nullptr);
  if (rc < 0) return -1;
  const int n = (int)swf->A.rows();
  *n_out = n;
  if (n > cap_n) return -2;
  for (int r = 0; r < n; ++r) {
    if (!hessian_only) b_out[r] = swf->b(r);
    for (int c = 0; c < n; ++c) A_out[(size_t)r * n + c] = swf->A(r, c);
  }
  return 0;
}
}

// ---- SWFOptimization::MyOrdering (swf_gnss.cpp:629-783) executed on a composition-A synthetic window ---------------------
// The window is built by shim/ceres_shim_refdemo.cpp from the reference's factor classes, but its parameter blocks live in
// an ESTIMATOR's storage -- para_pose[f], para_speed_bias[f], para_ex_Pose[0], f_manager.feature entries (landmarks),
// rtk_phase_bias_variables entries (ambiguities), blackvalue2 -- so that MyOrdering recognises them by address; it is then
// called where MyOptimization calls it (swf_image.cpp:212) and the solve runs with the ordering it produced.
#include "../rtk-visual-inertial-navigation_b200/synth/swgn_synth.h"
extern "C" void swgn_ceres_refdemo_set_build_hooks(double* (*block_memory)(int, int),
                                                   void (*before_solve)(ceres::Problem*, ceres::Solver::Options*, const swgn_graph*, double* const*));
namespace {
struct OrderingRun {
  SWFOptimization* swf = nullptr;
  int F = 0, n_lm = 0, first_amb = 0, n_amb = 0, n_blocks = 0;
  std::vector<int32_t> groups;
  std::vector<FeaturePerId*> features;
} g_mo;
double* mo_block_memory(int b, int size) {
  SWFOptimization& S = *g_mo.swf;
  if (b < g_mo.F) return S.para_pose[b];
  if (b < 2 * g_mo.F) return S.para_speed_bias[b - g_mo.F];
  if (b == 2 * g_mo.F) return S.para_ex_Pose[0];
  if (b < 2 * g_mo.F + 1 + g_mo.n_lm) {
    S.f_manager.feature.push_back(FeaturePerId(b - (2 * g_mo.F + 1), 0));
    return S.f_manager.feature.back().ptsInWorld.data();
  }
  if (b >= g_mo.first_amb && b < g_mo.first_amb + g_mo.n_amb) {
    PBtype n;
    n.value = 0;
    n.continue_count = 0;
    const int sat = b - g_mo.first_amb;
    S.rtk_phase_bias_variables[sat * 2].push_back(n);
    return &S.rtk_phase_bias_variables[sat * 2].back().value;
  }
  if (b == g_mo.n_blocks - 1) return &S.blackvalue2;
  (void)size;
  return &S.blackvalue;  // composition A holds no other scalar block
}
void mo_before_solve(ceres::Problem* problem, ceres::Solver::Options* options, const swgn_graph* g, double* const* ptr) {
  SWFOptimization& S = *g_mo.swf;
  S.image_count = g_mo.F;
  S.rover_count = 0;
  MarginalizationInfo* M = new MarginalizationInfo();
  if (g->n_prior > 0)
    for (int k = g->prior_blk_begin[0]; k < g->prior_blk_begin[1]; ++k) M->keep_block_addr.push_back(ptr[g->prior_blocks[k]]);
  S.last_marg_info = M;
  // the hidden GNSS frames are not parameter blocks of the reference's problem (SetLastImuFactor removes them,
  // gnss_imu_factor.cpp:108-113); the window builder added every block of the flat graph
  for (int b = 0; b < g->n_blocks; ++b) {
    std::vector<ceres::ResidualBlockId> rbs;
    problem->GetResidualBlocksForParameterBlock(ptr[b], &rbs);
    if (rbs.empty()) problem->RemoveParameterBlock(ptr[b]);
  }
  S.MyOrdering(*problem, *options);
  g_mo.groups.assign(g->n_blocks, -1);
  for (int b = 0; b < g->n_blocks; ++b) g_mo.groups[b] = options->linear_solver_ordering->GroupId(ptr[b]);
}
}  // namespace
extern "C" int ref_est_my_ordering(int which, uint64_t window_id, int cap_blocks, int32_t* n_blocks, int32_t* groups, double* state_out,
                                   int* steps_out) {
  swgn_synth_config cfg;
  swgn_synth_default_config(which, &cfg);
  if (cfg.composition != 1) return -1;  // MyOrdering knows the blocks of the reference's own graph: composition A
  swgn_synth* W = swgn_synth_create(&cfg, window_id);
  if (!W) return -1;
  const swgn_graph* g = swgn_synth_graph(W);
  int32_t info[8];
  swgn_synth_info(W, info);
  g_mo = OrderingRun();
  g_mo.swf = new SWFOptimization();
  g_mo.n_blocks = g->n_blocks;
  for (int b = 0; b < g->n_blocks; ++b) {
    if (g->block_size[b] == 9) g_mo.F++;
    if (g->block_size[b] == 3) g_mo.n_lm++;
  }
  g_mo.n_amb = info[4];
  g_mo.first_amb = info[5];
  *n_blocks = g->n_blocks;
  if (g->n_blocks > cap_blocks || g_mo.F > FEATURE_WINDOW_SIZE + GNSS_WINDOW_SIZE) {
    swgn_synth_destroy(W);
    return -2;
  }
  swgn_synth_destroy(W);
  NUM_OF_CAM = 1;
  swgn_ceres_refdemo_set_build_hooks(&mo_block_memory, &mo_before_solve);
  char msg[256];
  double cost[4];
  const int rc = swgn_ceres_refdemo_solve(which, window_id, 0, 0, 0, 0, state_out, cost, steps_out, msg, sizeof(msg));
  swgn_ceres_refdemo_set_build_hooks(nullptr, nullptr);
  NUM_OF_CAM = 0;
  if (rc < 0 || (int)g_mo.groups.size() != *n_blocks) return -3;
  for (int b = 0; b < *n_blocks; ++b) groups[b] = g_mo.groups[b];
  return rc;
}

// ---- SWFOptimization::AddAllResidual(NormalMode, ...) (swf_core.cpp:209-412) executed: the reference's OWN assembly of the
// sliding-window problem from estimator state -- blackvalue2, IMU / IMUGNSS factors per image gap, visual features from
// f_manager, the marginalisation prior -- followed by AddParameter2Problem, MyOrdering and ceres::Solve, which is what
// MyOptimization runs without USE_GLOBAL_OPTIMIZATION (swf_image.cpp:241-250).  The estimator state is filled from a
// composition-A synthetic window (its flat graph only serves as the source of the numbers).
extern "C" void swgn_ceres_refdemo_prepare(const swgn_graph* g);
extern "C" void* swgn_ceres_refdemo_integration(const double* record);
namespace {
// An estimator filled from a composition-A synthetic window (see ref_est_add_all_residual)
struct FilledEstimator {
  swgn_synth* W = nullptr;
  const swgn_graph* g = nullptr;
  swgn_options so;
  SWFOptimization* S = nullptr;
  std::vector<double*> ptr;         // block -> estimator storage
  std::vector<double>* hidden = nullptr;
  int n_hidden = 0, F = 0, n_lm = 0, b_lm = 0;
  std::vector<int> image_of_frame;
};
int fill_estimator(int which, uint64_t window_id, int cap_state, FilledEstimator& X) {

  swgn_synth_config cfg;
  swgn_synth_default_config(which, &cfg);
  if (cfg.composition != 1) return -1;
  swgn_synth* W = X.W = swgn_synth_create(&cfg, window_id);
  if (!W) return -1;
  const swgn_graph* g = X.g = swgn_synth_graph(W);
  swgn_synth_options(W, &X.so);
  int32_t info[8];
  swgn_synth_info(W, info);
  if (g->n_state > cap_state) return -2;
  swgn_ceres_refdemo_prepare(g);
  int F = 0, n_lm = 0;
  for (int b = 0; b < g->n_blocks; ++b) {
    if (g->block_size[b] == 9) F++;
    if (g->block_size[b] == 3) n_lm++;
  }
  const int b_ext = 2 * F, b_lm = 2 * F + 1, first_amb = info[5], n_amb = info[4], b_black2 = g->n_blocks - 1;
  SWFOptimization& S = *(X.S = new SWFOptimization());
  USE_IMAGE = true;
  USE_GLOBAL_OPTIMIZATION = false;
  USE_MAG_CORRECT_YAW = false;
  NUM_OF_CAM = 1;
  ESTIMATE_EXTRINSIC = 0;
  // block -> estimator storage
  std::vector<double*>& ptr = X.ptr;
  ptr.assign(g->n_blocks, nullptr);
  for (int f = 0; f < F; ++f) {
    ptr[f] = S.para_pose[f];
    ptr[F + f] = S.para_speed_bias[f];
  }
  ptr[b_ext] = S.para_ex_Pose[0];
  for (int l = 0; l < n_lm; ++l) {
    S.f_manager.feature.push_back(FeaturePerId(l, 0));
    ptr[b_lm + l] = S.f_manager.feature.back().ptsInWorld.data();
  }
  for (int a = 0; a < n_amb; ++a) {
    PBtype n;
    n.value = 0;
    n.continue_count = 0;
    S.rtk_phase_bias_variables[a * 2].push_back(n);
    ptr[first_amb + a] = &S.rtk_phase_bias_variables[a * 2].back().value;
  }
  ptr[b_black2] = &S.blackvalue2;
  for (int b = 0; b < g->n_blocks; ++b) {
    if (!ptr[b]) return -3;  // composition A holds no other block
    std::memcpy(ptr[b], g->state + g->block_offset[b], sizeof(double) * g->block_size[b]);
  }
  // frames: keyframes are the pose blocks some factor touches, the others are the hidden GNSS frames of the chains
  std::vector<char> touched(g->n_blocks, 0);
  for (int k = 0; k < 3 * g->n_proj; ++k) touched[g->proj_blocks[k]] = 1;
  for (int k = 0; k < 4 * g->n_imu; ++k) touched[g->imu_blocks[k]] = 1;
  for (int k = 0; k < g->chain_blk_begin[g->n_chain]; ++k) touched[g->chain_blocks[k]] = 1;
  S.image_count = 0;
  S.rover_count = 0;
  std::vector<int>& image_of_frame = X.image_of_frame;
  image_of_frame.assign(F, -1);
  for (int f = 0; f < F; ++f) {
    if (touched[f]) {
      S.frame_types[f] = SWFOptimization::ImagFrame;
      image_of_frame[f] = S.image_count;
      S.i2f[S.image_count++] = f;
    } else {
      S.frame_types[f] = SWFOptimization::GnssFrame;
      mea_t* m = new mea_t();
      std::memset((void*)m, 0, sizeof(mea_t));
      S.rovers[S.rover_count] = m;
      S.g2f[S.rover_count++] = f;
    }
  }
  for (int i = 0; i < FEATURE_WINDOW_SIZE + 2; ++i) S.imu_gnss_factor[i] = nullptr;
  // plain IMU links: pre_integrations[frame_j]
  for (int i = 0; i < g->n_imu; ++i) {
    const int fj = g->imu_blocks[4 * i + 2];
    if (fj != g->imu_blocks[4 * i] + 1) return -4;
    S.pre_integrations[fj] = (IntegrationBase*)swgn_ceres_refdemo_integration(g->imu_data + (size_t)SWGN_IMU_STRIDE * i);
  }
  // chains: IMUGNSSBase objects, members as AddMargInfo / SetLastImuFactor leave them
  const int n_hidden = g->n_chain > 0 ? g->chain_frame_begin[g->n_chain] : 0;
  std::vector<double>& hidden = *(X.hidden = new std::vector<double>((size_t)16 * n_hidden));
  std::vector<double>& hidden_lin = *new std::vector<double>((size_t)16 * n_hidden);
  std::vector<IMUGNSSBase*> bases;
  {
    size_t fN = 0, cN = 0, imu = 0;
    for (int c = 0; c < g->n_chain; ++c) {
      const int b0 = g->chain_blk_begin[c], k = g->chain_blk_begin[c + 1] - b0 - 4;
      const int f0 = g->chain_frame_begin[c], m = g->chain_frame_begin[c + 1] - f0;
      const int fi = g->chain_blocks[b0], fj = g->chain_blocks[b0 + 2];
      if (fj - fi != m + 1 || image_of_frame[fj] < 0) return -5;
      IMUGNSSBase* B = new IMUGNSSBase(ptr[fi], ptr[F + fi], &S.my_problem);
      bases.push_back(B);
      for (int i = 0; i < m; ++i) {
        const double* fr = g->chain_frame_data + (size_t)SWGN_CHAIN_FRAME_STRIDE * (f0 + i);
        double* h = &hidden[(size_t)16 * (f0 + i)];
        double* hl = &hidden_lin[(size_t)16 * (f0 + i)];
        std::memcpy(h, fr + SWGN_CHAIN_POSE, sizeof(double) * 16);
        std::memcpy(hl, fr + SWGN_CHAIN_POSE_LIN, sizeof(double) * 16);
        B->gnss_poses.push_back(h);
        B->gnss_speed_bias.push_back(h + 7);
        B->gnss_poses_lin.push_back(hl);
        B->gnss_speed_bias_lin.push_back(hl + 7);
        Eigen::Matrix<double, 15, 15, Eigen::RowMajor> H;
        Eigen::Matrix<double, 15, 1, Eigen::ColMajor> rhs;
        Eigen::Matrix<double, 15, Eigen::Dynamic, Eigen::RowMajor> HN(15, k);
        for (int a = 0; a < 15; ++a) {
          rhs(a) = fr[SWGN_CHAIN_RHS + a];
          for (int q = 0; q < 15; ++q) H(a, q) = fr[SWGN_CHAIN_HESSIAN + 15 * a + q];
          for (int q = 0; q < k; ++q) HN(a, q) = g->chain_frame_N[fN + ((size_t)i * 15 + a) * k + q];
        }
        B->pose_hessians.push_back(H);
        B->pose_rhses.push_back(rhs);
        B->pose_phase_biases_hessians.push_back(HN);
        // every frame of the estimator has its pre-integration from the frame before (AddAllResidual reads all of them)
        S.pre_integrations[fi + 1 + i] = (IntegrationBase*)swgn_ceres_refdemo_integration(g->chain_imu_data + (size_t)SWGN_IMU_STRIDE * (imu + i));
        B->imu_factors.push_back(new IMUFactor(S.pre_integrations[fi + 1 + i]));
      }
      S.pre_integrations[fj] = (IntegrationBase*)swgn_ceres_refdemo_integration(g->chain_imu_data + (size_t)SWGN_IMU_STRIDE * (imu + m));
      B->last_imu_factor = new IMUFactor(S.pre_integrations[fj]);
      B->pose1_pose2_hessians.setZero();
      B->phase_biases_hessians.resize(k, k);
      B->phase_biases_rhs.resize(k);
      B->param = {ptr[fi], ptr[F + fi], ptr[fj], ptr[F + fj]};
      for (int a = 0; a < k; ++a) {
        B->phase_biases_rhs(a) = g->chain_N[cN + (size_t)k * k + a];
        for (int q = 0; q < k; ++q) B->phase_biases_hessians(a, q) = g->chain_N[cN + (size_t)a * k + q];
        B->gnss_phase_biases.push_back(ptr[g->chain_blocks[b0 + 4 + a]]);
        B->param.push_back(ptr[g->chain_blocks[b0 + 4 + a]]);
      }
      B->gnss_Index = m;
      B->Init();
      S.imu_gnss_factor[image_of_frame[fj]] = B;  // the gap that ends at image image_of_frame[fj] (swf_core.cpp:224-240)
      fN += (size_t)m * 15 * k;
      cN += (size_t)k * k + k;
      imu += m + 1;
    }
  }
  // visual features: observations of landmark l on consecutive images from its first one
  {
    std::vector<std::vector<std::pair<int, int>>> obs(n_lm);  // (image index, projection factor)
    for (int i = 0; i < g->n_proj; ++i) {
      const int f = g->proj_blocks[3 * i], l = g->proj_blocks[3 * i + 2] - b_lm;
      obs[l].push_back({image_of_frame[f], i});
    }
    int l = 0;
    for (auto& feat : S.f_manager.feature) {
      std::sort(obs[l].begin(), obs[l].end());
      if (!obs[l].empty()) {
        feat.start_frame = obs[l][0].first;
        for (size_t q = 0; q < obs[l].size(); ++q) {
          if (obs[l][q].first != feat.start_frame + (int)q) return -6;  // the reference walks consecutive images
          Eigen::Matrix<double, 7, 1> pt;
          pt.setZero();
          pt(0) = g->proj_uv[2 * obs[l][q].second];
          pt(1) = g->proj_uv[2 * obs[l][q].second + 1];
          pt(2) = 1.0;
          feat.feature_per_frame.push_back(FeaturePerFrame(pt));
        }
      }
      ++l;
    }
  }
  // the marginalisation prior
  if (g->n_prior != 1) return -7;
  {
    MarginalizationInfo* M = new MarginalizationInfo();
    const int n = g->prior_n[0];
    M->n = n;
    M->m = 0;
    std::vector<double>& lin = *new std::vector<double>();
    const double* x0 = g->prior_x0 + g->prior_x0_begin[0];
    int nx = 0;
    for (int k = g->prior_blk_begin[0]; k < g->prior_blk_begin[1]; ++k) nx += g->block_size[g->prior_blocks[k]];
    lin.assign(x0, x0 + nx);
    int xo = 0;
    for (int k = g->prior_blk_begin[0]; k < g->prior_blk_begin[1]; ++k) {
      const int b = g->prior_blocks[k];
      M->keep_block_size.push_back(g->block_size[b]);
      M->keep_block_idx.push_back(g->prior_blk_idx[k]);
      M->keep_block_data.push_back(lin.data() + xo);
      M->keep_block_addr.push_back(ptr[b]);
      xo += g->block_size[b];
    }
    M->linearized_jacobians.resize(n, n);
    M->linearized_residuals = Eigen::VectorXd(n);
    for (int a = 0; a < n; ++a) {
      M->linearized_residuals(a) = g->prior_r0[g->prior_r_begin[0] + a];
      for (int c = 0; c < n; ++c) M->linearized_jacobians(a, c) = g->prior_J[g->prior_J_begin[0] + (size_t)a * n + c];
    }
    S.last_marg_info = M;
  }
  X.n_hidden = n_hidden;
  X.F = F;
  X.n_lm = n_lm;
  X.b_lm = b_lm;
  return 0;
}
void release_estimator(FilledEstimator& X) {
  ceres::internal::parameter_head.clear();
  ceres::internal::is_optimize = true;
  swgn_synth_destroy(X.W);
  USE_IMAGE = false;
  NUM_OF_CAM = 0;
}
}  // namespace

extern "C" int ref_est_add_all_residual(int which, uint64_t window_id, int cap_state, double* state_out, double* chain_frames_out,
                                        int32_t* n_frames_out) {
  FilledEstimator X;
  const int rc = fill_estimator(which, window_id, cap_state, X);
  if (rc) return rc;
  SWFOptimization& S = *X.S;
  const swgn_graph* g = X.g;
  int32_t info[8];
  swgn_synth_info(X.W, info);
  ceres::internal::parameter_head.clear();
  for (int k = 0; k < X.so.n_parameter_head; ++k) ceres::internal::parameter_head.push_back(X.ptr[info[5] + k]);
  ceres::internal::is_optimize = true;
  {
    ceres::Problem problem;
    ceres::Solver::Options options;
    S.AddAllResidual(SWFOptimization::NormalMode, std::set<double*>{}, nullptr, problem, options, true, true, true);
  }
  for (int b = 0; b < g->n_blocks; ++b) std::memcpy(state_out + g->block_offset[b], X.ptr[b], sizeof(double) * g->block_size[b]);
  *n_frames_out = X.n_hidden;
  if (chain_frames_out) std::memcpy(chain_frames_out, X.hidden->data(), sizeof(double) * X.hidden->size());
  release_estimator(X);
  return 0;
}

// SWFOptimization::AddAllResidual(MargeIncludeMode2) (RVI/swf/swf_core.cpp:209-468), the marginalisation of the oldest image
// frame as MargFrames runs it (RVI/swf/swf.cpp:343-364): MargePoint = the frame's pose and speed-bias and every landmark whose
// track starts there.  The reference code collects the factors touching them, appends the blocks to keep to
// ceres::internal::parameter_head, calls MyOrdering, ceres::Solve (export mode), UpdateSchur and
// MarginalizationInfo::setmarginalizeinfo / getParameterBlocks.  Returns the new prior: drop flags per block, the keep blocks
// (graph block index, first column) in the reference's order, linearized_jacobians (n x n row-major), linearized_residuals.
extern "C" int ref_est_marginalize_oldest(int which, uint64_t window_id, int cap_n, uint8_t* drop_out, int32_t* n_out, int32_t* n_keep_out,
                                          int32_t* keep_block_out, int32_t* keep_idx_out, double* J0_out, double* r0_out) {
  FilledEstimator X;
  const int rc = fill_estimator(which, window_id, 1 << 30, X);
  if (rc) return rc;
  SWFOptimization& S = *X.S;
  const swgn_graph* g = X.g;
  std::set<double*> marge{S.para_pose[0], S.para_speed_bias[0]};
  std::memset(drop_out, 0, g->n_blocks);
  drop_out[0] = drop_out[X.F] = 1;
  {
    int l = 0;
    for (auto& feat : S.f_manager.feature) {
      if ((int)feat.feature_per_frame.size() >= FEATURE_CONTINUE && feat.start_frame == 0) {
        marge.insert(feat.ptsInWorld.data());
        drop_out[X.b_lm + l] = 1;
      }
      ++l;
    }
  }
  ceres::internal::parameter_head.clear();
  ceres::internal::is_optimize = true;
  MarginalizationInfo* M = new MarginalizationInfo();
  {
    ceres::Problem problem;
    ceres::Solver::Options options;
    S.AddAllResidual(SWFOptimization::MargeIncludeMode2, marge, M, problem, options, true, true, true);
  }
  int ret = 0;
  if (S.last_marg_info != M) {
    ret = -10;
  } else {
    const int n = (int)M->linearized_jacobians.rows();
    *n_out = n;
    *n_keep_out = (int32_t)M->keep_block_addr.size();
    if (n > cap_n || n != (int)M->linearized_jacobians.cols()) {
      ret = -11;
    } else {
      for (size_t k = 0; k < M->keep_block_addr.size(); ++k) {
        int b = -1;
        for (int q = 0; q < g->n_blocks; ++q)
          if (X.ptr[q] == M->keep_block_addr[k]) b = q;
        if (b < 0 || M->keep_block_size[k] != g->block_size[b]) ret = -12;
        keep_block_out[k] = b;
        keep_idx_out[k] = M->keep_block_idx[k];
      }
      for (int a = 0; a < n; ++a) {
        r0_out[a] = M->linearized_residuals(a);
        for (int c = 0; c < n; ++c) J0_out[(size_t)a * n + c] = M->linearized_jacobians(a, c);
      }
    }
  }
  release_estimator(X);
  return ret;
}
