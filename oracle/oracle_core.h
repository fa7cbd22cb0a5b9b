// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement ("oracle") of the reference's sliding-window Gauss-Newton path:
// the app's factors (RVI/factor/*.cpp), the modified Ceres 2.0.0 pipeline it drives
// (CERES/internal/ceres/*), the read-backs (RVI/swf/swf_gnss.cpp:25-94) and the ambiguity-fix
// decision (RVI/swf/swf_lambda.cpp, RVI/gnss/src/lambda.cpp).  Plain C++17, no Eigen.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or call anything in this directory.  Nothing under rtk-visual-inertial-navigation_b200/
// links or includes it.
//
// Parity pin status (the table in DESIGN.md section 2 lists the test behind every line):
//   * Schur eliminate / reduced solve / back-substitute: pinned on the reference's own known-answer fixture
//     (CERES linear_least_squares_problems.cc:135-178, problems 2-4); dogleg / Levenberg-Marquardt steps on the fixtures
//     of dogleg_strategy_test.cc; loss correction and CauchyLoss on corrector_test.cc / loss_function_test.cc.
//   * lambda(), matinv(), distance(), velecitydistance(), update_azel(): pinned against the reference's own sources
//     compiled into oracle/_ref/libref_gnss.so (oracle/build_ref.sh).
//   * factors a1 / a3 / a4 / a5, IntegrationBase, PoseLocalParameterization, varerr2: pinned on the reference's own classes
//     (RVI/factor/*.cpp compiled unmodified into oracle/_ref against a minimal Eigen stand-in; GNSS factors bit-exact).
//   * MarginalizationInfo::marginalize + MarginalizationFactor: pinned on the reference's own marginalization_factor.cpp
//     compiled into oracle/_ref (tests/test_gnss_epoch.py).
//   * LambdaSearch (decision + prior rebuild) and GnssPreprocess: pinned on the reference's own estimator code -- swf_lambda.cpp,
//     swf_gnss.cpp, swf_core.cpp compiled unmodified into oracle/_ref/libref_estimator.so and executed
//     (tests/test_gnss_epoch.py); UpdateSchur / UpdateSchurHessianOnly executed on the shim's exports (tests/test_ceres_shim.py).
//     MyOrdering executed on composition-A windows whose blocks live in an estimator's storage (tests/test_ceres_shim.py).
//   * IMUGNSSFactor (oracle_chain.cpp): pinned on the reference's own IMUGNSSBase::Evaluate -- gnss_imu_factor.cpp
//     compiled unmodified into oracle/_ref and executed on synthetic chains through the Jacobian / cost-only /
//     Jacobian protocol, hidden-state back-substitution included -- and on the dense Schur complement of the whole
//     chain (tests/test_chain_factor.py).
//
// RVI/   = /root/reference/rtk_visual_inertial_src/rtk_visual_inertial/src/
// CERES/ = ceres-solver-modified/ inside /root/reference/ceres-solver-modified.tar
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../include/swgn.h"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// tiny dense algebra (row-major), standing in for the Eigen calls the reference makes
// ---------------------------------------------------------------------------------------------
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
  double* data() { return a.data(); }
  const double* data() const { return a.data(); }
  static Mat Identity(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
  }
};
Mat matmul(const Mat& A, const Mat& B);
Mat transpose(const Mat& A);

// Eigen::LLT<Matrix, Upper> on selfadjointView<Upper>(): returns false on a non-positive pivot.
// On success U (row-major upper, A = U^T U) overwrites the upper triangle of a (n x n, ld n).
bool llt_upper_inplace(double* a, int n);
// solve (U^T U) x = b given the factor from llt_upper_inplace
void llt_upper_solve(const double* u, int n, double* x);
// InvertPSDMatrix<Dynamic>(assume_full_rank = true, m): LLT-solve-identity
// (CERES/internal/ceres/invert_psd_matrix.h:62-67).  m is n x n row-major (upper used).
bool invert_psd(const double* m, int n, double* inv);
// general inverse by LU with partial pivoting (Eigen's MatrixXd::inverse() for n > 4)
bool inverse_lu(const Mat& A, Mat* inv);
// symmetric eigen-decomposition (cyclic Jacobi), eigenvalues ascending like
// Eigen::SelfAdjointEigenSolver; V columns are eigenvectors.
void eig_sym(const Mat& A, std::vector<double>* w, Mat* V);

// ---------------------------------------------------------------------------------------------
// quaternion helpers following Eigen's formulas and RVI/utility/utility.h:11-49
// storage order of a pose block: (px,py,pz,qx,qy,qz,qw)
// ---------------------------------------------------------------------------------------------
struct Quat {
  double w, x, y, z;
};
inline Quat qmul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
          a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Quat qinv(const Quat& q) {  // Eigen: conjugate / squaredNorm
  double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}
inline Quat qnormalized(const Quat& q) {
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
inline void qrot(const Quat& q, const double v[3], double out[3]) {  // Eigen _transformVector
  double uv[3] = {2.0 * (q.y * v[2] - q.z * v[1]), 2.0 * (q.z * v[0] - q.x * v[2]),
                  2.0 * (q.x * v[1] - q.y * v[0])};
  out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
  out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
  out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
void qtoR(const Quat& q, double R[9]);                    // toRotationMatrix, row-major
inline Quat deltaQ(const double th[3]) {                  // utility.h:11-22, NOT normalised
  return {1.0, th[0] / 2.0, th[1] / 2.0, th[2] / 2.0};
}
void skew(const double v[3], double S[9]);                // utility.h:25-31
void Qleft_br(const Quat& q, double M[9]);                // bottom-right 3x3 of Qleft, :34-40
void Qright_br(const Quat& q, double M[9]);               // bottom-right 3x3 of Qright, :43-49
inline Quat pose_q(const double* p) { return {p[6], p[3], p[4], p[5]}; }

// ---------------------------------------------------------------------------------------------
// cost functions / loss / parameterization (CERES/include/ceres/cost_function.h:116 etc.)
// ---------------------------------------------------------------------------------------------
struct CostFunction {
  std::vector<int> block_sizes;
  int num_residuals = 0;
  virtual ~CostFunction() {}
  // jacobians[i]: row-major num_residuals x block_sizes[i] (GLOBAL size), may be null
  virtual bool Evaluate(double const* const* parameters, double* residuals,
                        double** jacobians) const = 0;
};

struct AppGlobals {
  double Pbg[3];
  double gravity[3];  // Rwgw * G
  double proj_sqrt_info[4];
};

CostFunction* make_projection_factor(const AppGlobals* g, const double uv[2]);
CostFunction* make_imu_factor(const AppGlobals* g, const double* imu_record);
CostFunction* make_gnss_factor(int kind, const double* record);
CostFunction* make_prior_factor(int n, const std::vector<int>& sizes, const std::vector<int>& idx,
                                const double* x0, const double* J0, const double* r0);
CostFunction* make_unit_factor(double istd);
// IMUGNSSFactor (RVI/factor/gnss_imu_factor.cpp:678-835); arrays as in swgn_graph's chain_* fields
CostFunction* make_chain_factor(const AppGlobals* g, int m, int k, const double* frames, const double* frameN,
                                const double* chainN, const double* imu_data);
int chain_factor_frames(const CostFunction* c, double* out16_per_frame);

// IMU pre-integration -> SWGN_IMU_STRIDE record (RVI/factor/integration_base.cpp:5-142), oracle_preint.cpp
bool preintegrate(int n_samples, const double* samples7, const double* bias6, const double* noise4, double* record);

// range model: RVI/gnss/src/common_function.cpp:103-108,126-139,411-421
double dot_rtk(const double* a, const double* b, int n);
double distance_rtk(const double* rr, const double* rs, double* e);
double velocity_distance_rtk(const double* rr, const double* rs, const double* vr, const double* vs,
                             double* e);
// gnss_factor.cpp:98-103 (float sinf)
double varerr2(double el, double dt, double mea_var);

// CauchyLoss(a) CERES/internal/ceres/loss_function.cc:73-80
void cauchy_loss(double a, double s, double rho[3]);

// PoseLocalParameterization::Plus, RVI/factor/pose_local_parameterization.cpp:5-20
void pose_plus(const double* x, const double* delta, double* out);

// ---------------------------------------------------------------------------------------------
// the mini "Ceres": program, evaluator, Schur eliminator, dogleg, trust-region loop
// ---------------------------------------------------------------------------------------------
struct ParamBlock {
  int graph_index = -1;
  int size = 0, local = 0, manifold = 0;
  bool constant = false;
  int group = -1;
  double* user_state = nullptr;  // into Solver::state
  int index = -1;                // position in the reduced program
  int state_offset = 0;          // offset in the reduced state vector
  int delta_offset = 0;          // offset in the tangent vector
};

struct ResidualBlock {
  std::unique_ptr<CostFunction> cost;
  double cauchy_a = 0.0;  // <= 0: no loss
  std::vector<ParamBlock*> blocks;
  bool is_use = true;
  int program_index = 0;  // position in the user's program order
};

struct Cell {
  int block_id;   // column block
  int position;   // offset in values
};
struct RowBlock {
  int size, position;       // residual rows
  std::vector<Cell> cells;  // sorted by block_id
};
struct ColBlock {
  int size, position;
};
struct BlockSparse {
  std::vector<ColBlock> cols;
  std::vector<RowBlock> rows;
  std::vector<double> values;
  int num_rows = 0, num_cols = 0;
};

// SchurEliminator<-1,-1,-1> on a block-sparse matrix; returns S (n_f x n_f row-major, block
// upper triangle), rhs; CERES/internal/ceres/schur_eliminator_impl.h:177-306.
struct Chunk {
  int start, size;
  std::map<int, int> buffer_layout;
};
struct SchurEliminator {
  int num_eliminate_blocks = 0;
  std::vector<Chunk> chunks;
  std::vector<int> lhs_row_layout;
  int uneliminated_row_begins = 0;
  int buffer_size = 1;
  int lhs_num_rows = 0;
  void Init(int num_e, const BlockSparse& bs);
  void Eliminate(const BlockSparse& A, const double* b, const double* D, double* lhs,
                 double* rhs) const;
  void BackSubstitute(const BlockSparse& A, const double* b, const double* D, const double* z,
                      double* y) const;
};

struct Exports {  // ceres::internal::{lhs_out, rhs_out, lhs_out2, hs_row}
  int hs_row = 0;
  std::vector<double> lhs_out, rhs_out, lhs_out2;
  bool have_reduced = false, have_factor = false;
};

struct IterationRecord {
  double cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius;
  int step_is_valid, step_is_successful;
};

struct Solver {
  // inputs
  swgn_options opt;
  AppGlobals globals;
  std::vector<double> state;  // user state (all blocks, graph layout)
  std::vector<ParamBlock> blocks;
  std::vector<std::unique_ptr<ResidualBlock>> residual_blocks;  // program order
  std::string error;

  // reduced program
  std::vector<ParamBlock*> pblocks;      // ordered columns
  std::vector<ResidualBlock*> rblocks;   // ordered rows
  int num_eliminate_blocks = 0;
  int num_parameters = 0, num_effective_parameters = 0, num_residuals = 0;
  double fixed_cost = 0.0;
  BlockSparse jac;
  std::vector<int> residual_layout;
  SchurEliminator eliminator;
  Exports exports;
  std::vector<IterationRecord> iterations;
  int num_linear_solves = 0;

  bool Build(const swgn_graph* g, const swgn_options* o);
  bool Preprocess();
  // ProgramEvaluator::Evaluate; x in reduced-state layout.  Any output may be null.
  bool Evaluate(const double* x, double* cost, double* residuals, double* gradient, bool jacobian);
  void Plus(const double* x, const double* delta, double* out) const;
  // DENSE_SCHUR linear solve of min |J y - r|^2 + |D y|^2
  bool LinearSolve(const double* residuals, const double* D, double* y, bool* exported_only);
  bool Minimize(swgn_summary* summary);
  void StateToUser(const double* x);
};

// ---------------------------------------------------------------------------------------------
// ambiguity resolution
// ---------------------------------------------------------------------------------------------
// lambda(): RVI/gnss/src/lambda.cpp:204-235 (column-major like the reference)
int lambda_rtk(int n, int m, const double* a, const double* Q, double* F, double* s);
// matinv(): RVI/gnss/src/common_function.cpp:165-189,348-386 (LU, column-major, in place)
int matinv_rtk(double* A, int n);
// decision part of LambdaSearch: RVI/swf/swf_lambda.cpp:8-53,101-245
int ambiguity_fix(int n, const double* A, const double* y, int n_epochs, const int* epoch_begin,
                  const int* obs_amb, const int* obs_sysfreq, int last_fix, int* dd_pairs,
                  double* F, swgn_fix_result* res);

}  // namespace oracle
