// TEST INFRASTRUCTURE -- see oracle_core.h.  Restatement of the stateful IMUGNSSFactor
// (RVI/factor/gnss_imu_factor.cpp): the GNSS frames between two consecutive keyframes are hidden
// inside one cost function that eliminates them frame by frame on every Jacobian evaluation.
//   IMUGNSSBase::Evaluate              :678-799
//   JacobianResidualUpdateHessianRhs   :358-379
//   MargPose1 / MoveHessianData        :403-456
//   UpdateSchurComponent               :458-494   (SelfAdjointEigenSolver -> eig_sym)
//   UpdateJacobResidual                :495-530
//   UpdateRhsN / UpdateRhsPose         :532-564
//   UpdateDeltaValues                  :566-611
//   UpdateHiddenState                  :613-646
//   GetInc                             :676-691
// The "middle marginalisation" link (pose1_pose2_hessians, :741-760) is not part of the flat graph
// (include/swgn.h) and is not restated.
// Pinned on the reference's own IMUGNSSBase::Evaluate, compiled from gnss_imu_factor.cpp into oracle/_ref and executed on
// the same chains (tests/test_chain_factor.py::test_oracle_chain_matches_the_reference_imugnss_factor).
#include <cstdlib>

#include "oracle_core.h"

namespace oracle {
namespace {

enum { O_Pose1 = 0, O_Pose2 = 1, O_N = 2, O_Pose0 = 3, O_Size = 4 };

struct ChainFactor : CostFunction {
  const AppGlobals* g;
  int m, k;  // hidden frames, phase biases
  std::vector<std::unique_ptr<CostFunction>> imu;  // m + 1: imu_factors[0..m-1], last_imu_factor
  // constants
  std::vector<double> pose_lin, sb_lin;            // m x 7, m x 9
  std::vector<double> pose_hessians;               // m x 225
  std::vector<double> pose_N;                      // m x 15 x k
  std::vector<double> pose_rhs;                    // m x 15
  std::vector<double> NN, N_rhs;                   // k x k, k
  // mutable state of the reference object
  mutable std::vector<double> pose, sb;            // m x 7, m x 9 (gnss_poses, gnss_speed_bias)
  mutable bool history_flag = false, update_flag = false;
  mutable std::vector<double> H[O_Size * O_Size], rhs5[O_Size], delta5[O_Size];
  mutable std::vector<std::vector<double>> hmn_save[O_Size];
  mutable std::vector<std::vector<double>> rhsmn_save;
  mutable std::vector<double> Nval, Nval_old, INC;
  mutable double Pi_old[7], Bi_old[9], Pj_old[7], Bj_old[9];
  mutable Mat schur_jacobian;
  mutable std::vector<double> schur_residual;
  int hsize[O_Size];

  ChainFactor(const AppGlobals* g_, int m_, int k_, const double* frames, const double* frameN,
              const double* chainN, const double* imu_data)
      : g(g_), m(m_), k(k_) {
    block_sizes = {7, 9, 7, 9};
    for (int i = 0; i < k; ++i) block_sizes.push_back(1);
    num_residuals = 30 + k;
    for (int i = 0; i <= m; ++i) imu.emplace_back(make_imu_factor(g, imu_data + (size_t)SWGN_IMU_STRIDE * i));
    for (int i = 0; i < m; ++i) {
      const double* f = frames + (size_t)SWGN_CHAIN_FRAME_STRIDE * i;
      pose.insert(pose.end(), f + SWGN_CHAIN_POSE, f + SWGN_CHAIN_POSE + 7);
      sb.insert(sb.end(), f + SWGN_CHAIN_SB, f + SWGN_CHAIN_SB + 9);
      pose_lin.insert(pose_lin.end(), f + SWGN_CHAIN_POSE_LIN, f + SWGN_CHAIN_POSE_LIN + 7);
      sb_lin.insert(sb_lin.end(), f + SWGN_CHAIN_SB_LIN, f + SWGN_CHAIN_SB_LIN + 9);
      pose_rhs.insert(pose_rhs.end(), f + SWGN_CHAIN_RHS, f + SWGN_CHAIN_RHS + 15);
      pose_hessians.insert(pose_hessians.end(), f + SWGN_CHAIN_HESSIAN, f + SWGN_CHAIN_HESSIAN + 225);
    }
    pose_N.assign(frameN, frameN + (size_t)m * 15 * k);
    NN.assign(chainN, chainN + (size_t)k * k);
    N_rhs.assign(chainN + (size_t)k * k, chainN + (size_t)k * k + k);
    // Init / InitHessianRhs :33-97
    hsize[O_Pose1] = 15; hsize[O_Pose2] = 15; hsize[O_N] = k; hsize[O_Pose0] = 15;
    for (int i = 0; i < O_Size; ++i) {
      rhs5[i].assign(hsize[i], 0.0);
      delta5[i].assign(hsize[i], 0.0);
      for (int j = i; j < O_Size; ++j) H[i * O_Size + j].assign((size_t)hsize[i] * hsize[j], 0.0);
    }
    rhsmn_save.assign(m, std::vector<double>(15, 0.0));
    for (int i = O_Pose1; i < O_Size; ++i) hmn_save[i].assign(m, std::vector<double>((size_t)15 * hsize[i], 0.0));
    Nval.assign(k, 0.0);
    Nval_old.assign(k, 0.0);
    INC.assign(30 + k, 0.0);
  }

  // x [-] x0 of a (pose, speed-bias) pair with the sign fix of GetInc :676-691
  static void inc15(const double* x, const double* x0, const double* s, const double* s0, double sign, double* dx) {
    for (int i = 0; i < 3; ++i) dx[i] = sign * (x[i] - x0[i]);
    Quat q = qmul(qinv(pose_q(x0)), pose_q(x));
    double f = 2.0 * sign;
    if (!(q.w >= 0)) f = -f;
    dx[3] = f * q.x; dx[4] = f * q.y; dx[5] = f * q.z;
    for (int i = 0; i < 9; ++i) dx[6 + i] = sign * (s[i] - s0[i]);
  }

  // IMUFactor::Evaluate2 (imu_factor.cpp:103-193): 15x15 Jacobians over (pose 6 | speed-bias 9)
  void imu2(int idx, const double* pi, const double* si, const double* pj, const double* sj, double* r,
            double* J1, double* J2) const {
    double jpi[15 * 7], jsi[15 * 9], jpj[15 * 7], jsj[15 * 9];
    double* jac[4] = {jpi, jsi, jpj, jsj};
    const double* par[4] = {pi, si, pj, sj};
    imu[idx]->Evaluate(par, r, jac);
    for (int a = 0; a < 15; ++a) {
      for (int c = 0; c < 6; ++c) { J1[a * 15 + c] = jpi[a * 7 + c]; J2[a * 15 + c] = jpj[a * 7 + c]; }
      for (int c = 0; c < 9; ++c) { J1[a * 15 + 6 + c] = jsi[a * 9 + c]; J2[a * 15 + 6 + c] = jsj[a * 9 + c]; }
    }
  }
  // JacobianResidualUpdateHessianRhs :358-379 for two 15-dim blocks
  void add_jtj(int b0, int b1, const double* J0, const double* J1, const double* r) const {
    const int idx[2] = {b0, b1};
    const double* J[2] = {J0, J1};
    for (int i = 0; i < 2; ++i) {
      for (int c = 0; c < 15; ++c) {
        double s = 0.0;
        for (int a = 0; a < 15; ++a) s += J[i][a * 15 + c] * r[a];
        rhs5[idx[i]][c] += s;
      }
      for (int j = 0; j < 2; ++j) {
        if (idx[j] < idx[i]) continue;
        double* h = H[idx[i] * O_Size + idx[j]].data();
        for (int c = 0; c < 15; ++c)
          for (int d = 0; d < 15; ++d) {
            double s = 0.0;
            for (int a = 0; a < 15; ++a) s += J[i][a * 15 + c] * J[j][a * 15 + d];
            h[c * 15 + d] += s;
          }
      }
    }
  }
  void marg_pose1() const {  // MargPose1 :403-435
    double inv[225];
    invert_psd(H[O_Pose1 * O_Size + O_Pose1].data(), 15, inv);
    std::memcpy(H[O_Pose1 * O_Size + O_Pose1].data(), inv, sizeof(inv));
    for (int i = O_Pose1 + 1; i < O_Size; ++i) {
      const int sn = hsize[i];
      std::vector<double> AnmAmm((size_t)sn * 15, 0.0);
      const double* h1i = H[O_Pose1 * O_Size + i].data();  // 15 x sn
      for (int a = 0; a < sn; ++a)
        for (int c = 0; c < 15; ++c) {
          double s = 0.0;
          for (int t = 0; t < 15; ++t) s += h1i[t * sn + a] * inv[t * 15 + c];
          AnmAmm[(size_t)a * 15 + c] = s;
        }
      for (int a = 0; a < sn; ++a) {
        double s = 0.0;
        for (int c = 0; c < 15; ++c) s += AnmAmm[(size_t)a * 15 + c] * rhs5[O_Pose1][c];
        rhs5[i][a] -= s;
      }
      for (int j = i; j < O_Size; ++j) {
        const int sv = hsize[j];
        const double* h1j = H[O_Pose1 * O_Size + j].data();  // 15 x sv
        double* hij = H[i * O_Size + j].data();
        for (int a = 0; a < sn; ++a)
          for (int b = 0; b < sv; ++b) {
            double s = 0.0;
            for (int c = 0; c < 15; ++c) s += AnmAmm[(size_t)a * 15 + c] * h1j[c * sv + b];
            hij[(size_t)a * sv + b] -= s;
          }
      }
    }
  }
  void move_hessian(int index) const {  // MoveHessianData :437-456
    rhsmn_save[index] = rhs5[O_Pose1];
    for (int i = O_Pose1; i < O_Size; ++i) hmn_save[i][index] = H[O_Pose1 * O_Size + i];
    H[O_Pose1 * O_Size + O_Pose1] = H[O_Pose2 * O_Size + O_Pose2];
    std::fill(H[O_Pose2 * O_Size + O_Pose2].begin(), H[O_Pose2 * O_Size + O_Pose2].end(), 0.0);
    rhs5[O_Pose1] = rhs5[O_Pose2];
    std::fill(rhs5[O_Pose2].begin(), rhs5[O_Pose2].end(), 0.0);
    for (int i = O_Pose2 + 1; i < O_Size; ++i) {
      H[O_Pose1 * O_Size + i] = H[O_Pose2 * O_Size + i];
      std::fill(H[O_Pose2 * O_Size + i].begin(), H[O_Pose2 * O_Size + i].end(), 0.0);
    }
    std::fill(H[O_Pose1 * O_Size + O_Pose2].begin(), H[O_Pose1 * O_Size + O_Pose2].end(), 0.0);
  }
  void update_schur_component() const {  // :458-494
    const int n = 30 + k;
    const int mapindex[3] = {O_Pose0, O_Pose1, O_N};
    const int hidx[3] = {0, 15, 30};
    Mat Hd(n, n);
    std::vector<double> rd(n);
    for (int i = 0; i < 3; ++i) {
      const int i2 = mapindex[i];
      for (int a = 0; a < hsize[i2]; ++a) rd[hidx[i] + a] = rhs5[i2][a];
      for (int j = i; j < 3; ++j) {
        const int j2 = mapindex[j];
        for (int a = 0; a < hsize[i2]; ++a)
          for (int b = 0; b < hsize[j2]; ++b)
            Hd(hidx[i] + a, hidx[j] + b) = (j2 >= i2) ? H[i2 * O_Size + j2][(size_t)a * hsize[j2] + b]
                                                      : H[j2 * O_Size + i2][(size_t)b * hsize[i2] + a];
      }
    }
    if (const char* e = std::getenv("ORACLE_CHAIN_PERTURB")) {
      // sensitivity knob for tests: relative perturbation of the eliminated Hessian at the level of
      // rounding noise, to measure how much the eps-thresholded factorisation amplifies it
      const double rel = std::atof(e);
      uint64_t sd = 88172645463325252ull;
      for (int a = 0; a < n; ++a)
        for (int b = a; b < n; ++b) {
          sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17;
          Hd(a, b) *= 1.0 + rel * ((double)(sd >> 11) / 9007199254740992.0 - 0.5);
        }
    }
    std::vector<double> w;
    Mat V;
    eig_sym(Hd, &w, &V);  // uses the upper triangle (selfadjointView<Upper>)
    const double eps = 1e-8;
    schur_jacobian = Mat(n, n);
    schur_residual.assign(n, 0.0);
    for (int i = 0; i < n; ++i) {
      const double S = w[i] > eps ? w[i] : 0.0, Sinv = w[i] > eps ? 1.0 / w[i] : 0.0;
      const double ss = std::sqrt(S), si = std::sqrt(Sinv);
      double dotr = 0.0;
      for (int c = 0; c < n; ++c) {
        schur_jacobian(i, c) = ss * V(c, i);
        dotr += V(c, i) * rd[c];
      }
      schur_residual[i] = si * dotr;
    }
  }

  bool Evaluate(double const* const* p, double* residuals, double** jacobians) const override {
    const int n = 30 + k;
    for (int i = 0; i < k; ++i) Nval[i] = p[4 + i][0];
    const double *Pi = p[0], *Bi = p[1], *Pj = p[2], *Bj = p[3];
    auto save_last = [&]() {
      std::memcpy(Pi_old, Pi, sizeof(Pi_old));
      std::memcpy(Bi_old, Bi, sizeof(Bi_old));
      std::memcpy(Pj_old, Pj, sizeof(Pj_old));
      std::memcpy(Bj_old, Bj, sizeof(Bj_old));
      Nval_old = Nval;
    };
    if (!history_flag) save_last();
    // UpdateDeltaValues :566-611: INC = old [-] new
    inc15(Pj, Pj_old, Bj, Bj_old, -1.0, delta5[O_Pose2].data());
    for (int i = 0; i < k; ++i) delta5[O_N][i] = Nval_old[i] - Nval[i];
    inc15(Pi, Pi_old, Bi, Bi_old, -1.0, delta5[O_Pose0].data());
    for (int i = 0; i < 15; ++i) { INC[i] = delta5[O_Pose0][i]; INC[15 + i] = delta5[O_Pose2][i]; }
    for (int i = 0; i < k; ++i) INC[30 + i] = delta5[O_N][i];
    update_flag = jacobians != nullptr;
    if (history_flag && update_flag) {  // UpdateHiddenState :613-646
      for (int i = m - 1; i >= 0; --i) {
        for (int j = O_Pose2; j < O_Size; ++j)
          for (int a = 0; a < 15; ++a) {
            double s = 0.0;
            for (int b = 0; b < hsize[j]; ++b) s += hmn_save[j][i][(size_t)a * hsize[j] + b] * delta5[j][b];
            rhsmn_save[i][a] -= s;
          }
        for (int a = 0; a < 15; ++a) {
          double s = 0.0;
          for (int b = 0; b < 15; ++b) s += hmn_save[O_Pose1][i][a * 15 + b] * rhsmn_save[i][b];
          delta5[O_Pose2][a] = s;
        }
        double* P = pose.data() + 7 * i;
        double* B = sb.data() + 9 * i;
        const double* d = delta5[O_Pose2].data();
        for (int a = 0; a < 3; ++a) P[a] -= d[a];
        const double th[3] = {-d[3], -d[4], -d[5]};
        Quat q = qnormalized(qmul(pose_q(P), deltaQ(th)));
        P[3] = q.x; P[4] = q.y; P[5] = q.z; P[6] = q.w;
        for (int a = 0; a < 9; ++a) B[a] -= d[6 + a];
      }
    }
    if (!history_flag || update_flag) {
      history_flag = true;
      save_last();
      for (int i = 0; i < O_Size; ++i) {  // ResetMem
        std::fill(rhs5[i].begin(), rhs5[i].end(), 0.0);
        for (int j = i; j < O_Size; ++j) std::fill(H[i * O_Size + j].begin(), H[i * O_Size + j].end(), 0.0);
      }
      H[O_N * O_Size + O_N] = NN;
      rhs5[O_N] = N_rhs;
      for (int a = 0; a < k; ++a) {  // UpdateRhsN
        double s = 0.0;
        for (int b = 0; b < k; ++b) s += NN[(size_t)a * k + b] * Nval[b];
        rhs5[O_N][a] += s;
      }
      double r[15], J1[225], J2[225];
      imu2(0, Pi, Bi, pose.data(), sb.data(), r, J1, J2);
      add_jtj(O_Pose0, O_Pose1, J1, J2, r);
      for (int i = 0; i < m; ++i) {
        const double* hp = pose.data() + 7 * i;
        const double* hs = sb.data() + 9 * i;
        if (i != m - 1) imu2(i + 1, hp, hs, hp + 7, hs + 9, r, J1, J2);
        else imu2(m, hp, hs, Pj, Bj, r, J1, J2);
        add_jtj(O_Pose1, O_Pose2, J1, J2, r);
        // UpdateRhsPose(i) :540-564
        double dx[15];
        inc15(hp, pose_lin.data() + 7 * i, hs, sb_lin.data() + 9 * i, 1.0, dx);
        const double* ph = pose_hessians.data() + (size_t)225 * i;
        const double* pn = pose_N.data() + (size_t)15 * k * i;
        for (int a = 0; a < 15; ++a) {
          double s = 0.0;
          for (int b = 0; b < 15; ++b) s += ph[a * 15 + b] * dx[b];
          rhs5[O_Pose1][a] += s;
        }
        for (int a = 0; a < 15; ++a) {
          double s = 0.0;
          for (int b = 0; b < k; ++b) s += pn[(size_t)a * k + b] * Nval[b];
          rhs5[O_Pose1][a] += s;
        }
        for (int b = 0; b < k; ++b) {
          double s = 0.0;
          for (int a = 0; a < 15; ++a) s += pn[(size_t)a * k + b] * dx[a];
          rhs5[O_N][b] += s;
        }
        for (int a = 0; a < 225; ++a) H[O_Pose1 * O_Size + O_Pose1][a] += ph[a];
        for (int a = 0; a < 15 * k; ++a) H[O_Pose1 * O_Size + O_N][a] += pn[a];
        for (int a = 0; a < 15; ++a) rhs5[O_Pose1][a] += pose_rhs[(size_t)15 * i + a];
        marg_pose1();
        move_hessian(i);
      }
      update_schur_component();
    }
    // UpdateJacobResidual :495-530
    if (residuals) {
      for (int a = 0; a < n; ++a) {
        double s = schur_residual[a];
        if (!update_flag) {
          double t = 0.0;
          for (int c = 0; c < n; ++c) t += schur_jacobian(a, c) * INC[c];
          s -= t;
        }
        residuals[a] = s;
      }
    }
    if (jacobians) {
      const int col0[4] = {0, 6, 15, 21}, gs[4] = {7, 9, 7, 9}, ls[4] = {6, 9, 6, 9};
      for (int b = 0; b < 4; ++b) {
        if (!jacobians[b]) continue;
        for (int a = 0; a < n; ++a)
          for (int c = 0; c < gs[b]; ++c) jacobians[b][a * gs[b] + c] = c < ls[b] ? schur_jacobian(a, col0[b] + c) : 0.0;
      }
      for (int i = 0; i < k; ++i)
        if (jacobians[4 + i])
          for (int a = 0; a < n; ++a) jacobians[4 + i][a] = schur_jacobian(a, 30 + i);
    }
    return true;
  }
};

}  // namespace

CostFunction* make_chain_factor(const AppGlobals* g, int m, int k, const double* frames, const double* frameN,
                                const double* chainN, const double* imu_data) {
  return new ChainFactor(g, m, k, frames, frameN, chainN, imu_data);
}
// current hidden states (16 doubles per frame) of a cost function made by make_chain_factor
int chain_factor_frames(const CostFunction* c, double* out) {
  const ChainFactor* f = dynamic_cast<const ChainFactor*>(c);
  if (!f) return 0;
  if (out)
    for (int i = 0; i < f->m; ++i) {
      std::memcpy(out + 16 * i, f->pose.data() + 7 * i, sizeof(double) * 7);
      std::memcpy(out + 16 * i + 7, f->sb.data() + 9 * i, sizeof(double) * 9);
    }
  return f->m;
}

}  // namespace oracle
