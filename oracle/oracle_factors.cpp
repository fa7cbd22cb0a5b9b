// TEST INFRASTRUCTURE -- see oracle_core.h.  Restatement of the reference's factor classes:
// same Evaluate() contract (global-size row-major Jacobians, 7th pose column zero).
#include "oracle_core.h"

namespace oracle {
namespace {

inline void m33_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] +
                     A[i * 3 + 2] * B[2 * 3 + j];
}
inline void m33_vec(const double* A, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
inline void m33_T(const double* A, double* T) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}

// ---------------------------------------------------------------------------------------------
// projection_factor <2;7,7,3>   RVI/factor/projection_factor.cpp:13-65
// ---------------------------------------------------------------------------------------------
struct ProjectionFactor : CostFunction {
  const AppGlobals* g;
  double uv[2];
  ProjectionFactor(const AppGlobals* g_, const double* uv_) : g(g_) {
    uv[0] = uv_[0];
    uv[1] = uv_[1];
    block_sizes = {7, 7, 3};
    num_residuals = 2;
  }
  bool Evaluate(double const* const* p, double* residuals, double** jacobians) const override {
    const double* Pj = p[0];
    Quat Qj = pose_q(p[0]);
    const double* tic = p[1];
    Quat qic = pose_q(p[1]);
    const double* X = p[2];
    double d[3] = {X[0] - Pj[0], X[1] - Pj[1], X[2] - Pj[2]};
    double pts_imu[3];
    qrot(qinv(Qj), d, pts_imu);                                  // :21
    double t[3] = {pts_imu[0] + g->Pbg[0] - tic[0], pts_imu[1] + g->Pbg[1] - tic[1],
                   pts_imu[2] + g->Pbg[2] - tic[2]};
    double pc[3];
    qrot(qinv(qic), t, pc);                                      // :22
    double dep = pc[2];
    double r0 = pc[0] / dep - uv[0], r1 = pc[1] / dep - uv[1];   // :27
    const double* W = g->proj_sqrt_info;
    residuals[0] = W[0] * r0 + W[1] * r1;                        // :28
    residuals[1] = W[2] * r0 + W[3] * r1;
    if (!jacobians) return true;
    double Rj[9], ric[9];
    qtoR(Qj, Rj);
    qtoR(qic, ric);
    double red0[6] = {1. / dep, 0, -pc[0] / (dep * dep), 0, 1. / dep, -pc[1] / (dep * dep)};
    double red[6];
    for (int j = 0; j < 3; ++j) {                                // :38 reduce = sqrt_info * reduce
      red[j] = W[0] * red0[j] + W[1] * red0[3 + j];
      red[3 + j] = W[2] * red0[j] + W[3] * red0[3 + j];
    }
    double ricT[9], RjT[9];
    m33_T(ric, ricT);
    m33_T(Rj, RjT);
    if (jacobians[0]) {                                          // :40-49
      double negRjT[9], A[9], S[9], B[9];
      for (int i = 0; i < 9; ++i) negRjT[i] = -RjT[i];
      m33_mul(ricT, negRjT, A);
      skew(pts_imu, S);
      m33_mul(ricT, S, B);
      double* J = jacobians[0];
      for (int r = 0; r < 2; ++r) {
        for (int c = 0; c < 3; ++c) {
          J[r * 7 + c] = red[r * 3] * A[c] + red[r * 3 + 1] * A[3 + c] + red[r * 3 + 2] * A[6 + c];
          J[r * 7 + 3 + c] =
              red[r * 3] * B[c] + red[r * 3 + 1] * B[3 + c] + red[r * 3 + 2] * B[6 + c];
        }
        J[r * 7 + 6] = 0.0;
      }
    }
    if (jacobians[1]) {                                          // :50-57
      double S[9];
      skew(pc, S);
      double* J = jacobians[1];
      for (int r = 0; r < 2; ++r) {
        for (int c = 0; c < 3; ++c) {
          J[r * 7 + c] = -(red[r * 3] * ricT[c] + red[r * 3 + 1] * ricT[3 + c] +
                           red[r * 3 + 2] * ricT[6 + c]);
          J[r * 7 + 3 + c] =
              red[r * 3] * S[c] + red[r * 3 + 1] * S[3 + c] + red[r * 3 + 2] * S[6 + c];
        }
        J[r * 7 + 6] = 0.0;
      }
    }
    if (jacobians[2]) {                                          // :58-61
      double A[9], T[6];
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c)
          T[r * 3 + c] =
              red[r * 3] * ricT[c] + red[r * 3 + 1] * ricT[3 + c] + red[r * 3 + 2] * ricT[6 + c];
      (void)A;
      double* J = jacobians[2];
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c)
          J[r * 3 + c] =
              T[r * 3] * RjT[c] + T[r * 3 + 1] * RjT[3 + c] + T[r * 3 + 2] * RjT[6 + c];
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// IMUFactor <15;7,9,7,9>   RVI/factor/imu_factor.cpp:5-101, integration_base.cpp:144-174
// ---------------------------------------------------------------------------------------------
struct ImuFactor : CostFunction {
  const AppGlobals* g;
  double dp[3], dv[3], ba0[3], bg0[3], gyri[3], gyrj[3], sum_dt;
  Quat dq;
  double jac[225], sqrt_info[225];
  ImuFactor(const AppGlobals* g_, const double* rec) : g(g_) {
    block_sizes = {7, 9, 7, 9};
    num_residuals = 15;
    for (int i = 0; i < 3; ++i) {
      dp[i] = rec[SWGN_IMU_DELTA_P + i];
      dv[i] = rec[SWGN_IMU_DELTA_V + i];
      ba0[i] = rec[SWGN_IMU_LIN_BA + i];
      bg0[i] = rec[SWGN_IMU_LIN_BG + i];
      gyri[i] = rec[SWGN_IMU_GYRI + i];
      gyrj[i] = rec[SWGN_IMU_GYRJ + i];
    }
    dq = {rec[SWGN_IMU_DELTA_Q + 3], rec[SWGN_IMU_DELTA_Q], rec[SWGN_IMU_DELTA_Q + 1],
          rec[SWGN_IMU_DELTA_Q + 2]};
    sum_dt = rec[SWGN_IMU_SUM_DT];
    std::memcpy(jac, rec + SWGN_IMU_JACOBIAN, sizeof(jac));
    std::memcpy(sqrt_info, rec + SWGN_IMU_SQRT_INFO, sizeof(sqrt_info));
  }
  void block33(int r0, int c0, double* out) const {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) out[i * 3 + j] = jac[(r0 + i) * 15 + c0 + j];
  }
  static void set33(double* J, int ld, int r0, int c0, const double* B, double scale = 1.0) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J[(r0 + i) * ld + c0 + j] = scale * B[i * 3 + j];
  }
  void premul(const double* raw, int cols, double* out) const {  // out = sqrt_info * raw
    for (int i = 0; i < 15; ++i)
      for (int j = 0; j < cols; ++j) {
        double s = 0.0;
        for (int k = 0; k < 15; ++k) s += sqrt_info[i * 15 + k] * raw[k * cols + j];
        out[i * cols + j] = s;
      }
  }
  bool Evaluate(double const* const* p, double* residuals, double** jacobians) const override {
    enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };
    const double* Pi = p[0];
    Quat Qi = pose_q(p[0]);
    const double *Vi = p[1], *Bai = p[1] + 3, *Bgi = p[1] + 6;
    const double* Pj = p[2];
    Quat Qj = pose_q(p[2]);
    const double *Vj = p[3], *Baj = p[3] + 3, *Bgj = p[3] + 6;
    const double* Pbg = g->Pbg;
    const double* G = g->gravity;  // newG = Rwgw * G
    double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
    block33(O_P, O_BA, dp_dba);
    block33(O_P, O_BG, dp_dbg);
    block33(O_R, O_BG, dq_dbg);
    block33(O_V, O_BA, dv_dba);
    block33(O_V, O_BG, dv_dbg);
    double dba[3], dbg[3];
    for (int i = 0; i < 3; ++i) {
      dba[i] = Bai[i] - ba0[i];
      dbg[i] = Bgi[i] - bg0[i];
    }
    // integration_base.cpp:164-166
    double th[3];
    m33_vec(dq_dbg, dbg, th);
    Quat cdq = qmul(dq, deltaQ(th));
    double t1[3], t2[3], cdv[3], cdp[3];
    m33_vec(dv_dba, dba, t1);
    m33_vec(dv_dbg, dbg, t2);
    for (int i = 0; i < 3; ++i) cdv[i] = dv[i] + t1[i] + t2[i];
    m33_vec(dp_dba, dba, t1);
    m33_vec(dp_dbg, dbg, t2);
    for (int i = 0; i < 3; ++i) cdp[i] = dp[i] + t1[i] + t2[i];
    Quat Qi_inv = qinv(Qi);
    double QjPbg[3];
    qrot(Qj, Pbg, QjPbg);
    double wi[3] = {gyri[0] - Bgi[0], gyri[1] - Bgi[1], gyri[2] - Bgi[2]};
    double wj[3] = {gyrj[0] - Bgj[0], gyrj[1] - Bgj[1], gyrj[2] - Bgj[2]};
    double Swi[9], Swj[9], wiPbg[3], wjPbg[3];
    skew(wi, Swi);
    skew(wj, Swj);
    m33_vec(Swi, Pbg, wiPbg);
    m33_vec(Swj, Pbg, wjPbg);
    double a[3], ra[3];
    for (int i = 0; i < 3; ++i)                                  // :168
      a[i] = 0.5 * G[i] * sum_dt * sum_dt + ((Pj[i] - Pi[i]) - QjPbg[i]) - Vi[i] * sum_dt;
    qrot(Qi_inv, a, ra);
    double raw_r[15];
    for (int i = 0; i < 3; ++i) raw_r[O_P + i] = ra[i] - cdp[i] + Pbg[i] + wiPbg[i] * sum_dt;
    Quat qr = qmul(qinv(cdq), qmul(Qi_inv, Qj));                 // :169
    raw_r[O_R + 0] = 2 * qr.x;
    raw_r[O_R + 1] = 2 * qr.y;
    raw_r[O_R + 2] = 2 * qr.z;
    double QjwjPbg[3], b[3], rb[3];
    qrot(Qj, wjPbg, QjwjPbg);
    for (int i = 0; i < 3; ++i) b[i] = G[i] * sum_dt + (Vj[i] - QjwjPbg[i]) - Vi[i];  // :170
    qrot(Qi_inv, b, rb);
    for (int i = 0; i < 3; ++i) {
      raw_r[O_V + i] = rb[i] - cdv[i] + wiPbg[i];
      raw_r[O_BA + i] = Baj[i] - Bai[i];
      raw_r[O_BG + i] = Bgj[i] - Bgi[i];
    }
    for (int i = 0; i < 15; ++i) {                               // imu_factor.cpp:27
      double s = 0.0;
      for (int k = 0; k < 15; ++k) s += sqrt_info[i * 15 + k] * raw_r[k];
      residuals[i] = s;
    }
    if (!jacobians) return true;
    double Ri_inv[9], Rj[9];
    qtoR(Qi_inv, Ri_inv);
    qtoR(Qj, Rj);
    double negRi[9];
    for (int i = 0; i < 9; ++i) negRi[i] = -Ri_inv[i];
    double SPbg[9];
    skew(Pbg, SPbg);
    if (jacobians[0]) {                                          // :47-60
      double raw[15 * 7] = {0};
      set33(raw, 7, O_P, O_P, negRi);
      double S[9];
      skew(ra, S);
      set33(raw, 7, O_P, O_R, S);
      double th2[3];
      double dbg2[3] = {Bgi[0] - bg0[0], Bgi[1] - bg0[1], Bgi[2] - bg0[2]};
      m33_vec(dq_dbg, dbg2, th2);
      Quat cq = qmul(dq, deltaQ(th2));
      double L[9], R[9], LR[9];
      // -(Qleft(Qj^-1 Qi) * Qright(cq)).bottomRightCorner<3,3>(): the 4x4 product's lower-right
      // block = vecL * (-vecR)^T + L_br * R_br with the first row/col terms.
      Quat ql = qmul(qinv(Qj), Qi);
      Qleft_br(ql, L);
      Qright_br(cq, R);
      m33_mul(L, R, LR);
      double vl[3] = {ql.x, ql.y, ql.z}, vr[3] = {cq.x, cq.y, cq.z};
      double blk[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) blk[i * 3 + j] = -(vl[i] * (-vr[j]) + LR[i * 3 + j]);
      set33(raw, 7, O_R, O_R, blk);
      skew(rb, S);
      set33(raw, 7, O_V, O_R, S);
      double out[15 * 7];
      premul(raw, 7, out);
      std::memcpy(jacobians[0], out, sizeof(out));
    }
    if (jacobians[1]) {                                          // :61-75
      double raw[15 * 9] = {0};
      set33(raw, 9, O_P, 0, negRi, sum_dt);
      double M[9];
      for (int i = 0; i < 9; ++i) M[i] = -dp_dba[i];
      set33(raw, 9, O_P, 3, M);
      for (int i = 0; i < 9; ++i) M[i] = -dp_dbg[i] + SPbg[i] * sum_dt;
      set33(raw, 9, O_P, 6, M);
      Quat q3 = qmul(qmul(qinv(Qj), Qi), dq);
      double L[9], LB[9];
      Qleft_br(q3, L);
      m33_mul(L, dq_dbg, LB);
      for (int i = 0; i < 9; ++i) M[i] = -LB[i];
      set33(raw, 9, O_R, 6, M);
      set33(raw, 9, O_V, 0, negRi);
      for (int i = 0; i < 9; ++i) M[i] = -dv_dba[i];
      set33(raw, 9, O_V, 3, M);
      for (int i = 0; i < 9; ++i) M[i] = -dv_dbg[i] + SPbg[i];
      set33(raw, 9, O_V, 6, M);
      double negI[9] = {-1, 0, 0, 0, -1, 0, 0, 0, -1};
      set33(raw, 9, O_BA, 3, negI);
      set33(raw, 9, O_BG, 6, negI);
      double out[15 * 9];
      premul(raw, 9, out);
      std::memcpy(jacobians[1], out, sizeof(out));
    }
    if (jacobians[2]) {                                          // :76-86
      double raw[15 * 7] = {0};
      set33(raw, 7, O_P, O_P, Ri_inv);
      double RiRj[9], M[9];
      m33_mul(Ri_inv, Rj, RiRj);
      m33_mul(RiRj, SPbg, M);
      set33(raw, 7, O_P, O_R, M);
      double th2[3];
      double dbg2[3] = {Bgi[0] - bg0[0], Bgi[1] - bg0[1], Bgi[2] - bg0[2]};
      m33_vec(dq_dbg, dbg2, th2);
      Quat cq = qmul(dq, deltaQ(th2));
      Quat q3 = qmul(qmul(qinv(cq), qinv(Qi)), Qj);
      double L[9];
      Qleft_br(q3, L);
      set33(raw, 7, O_R, O_R, L);
      double S2[9];
      skew(wjPbg, S2);
      m33_mul(RiRj, S2, M);
      set33(raw, 7, O_V, O_R, M);
      double out[15 * 7];
      premul(raw, 7, out);
      std::memcpy(jacobians[2], out, sizeof(out));
    }
    if (jacobians[3]) {                                          // :87-96
      double raw[15 * 9] = {0};
      set33(raw, 9, O_V, 0, Ri_inv);
      double RiRj[9], M[9], negS[9];
      m33_mul(Ri_inv, Rj, RiRj);
      for (int i = 0; i < 9; ++i) negS[i] = -SPbg[i];
      m33_mul(RiRj, negS, M);
      set33(raw, 9, O_V, 6, M);
      double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      set33(raw, 9, O_BA, 3, I);
      set33(raw, 9, O_BG, 6, I);
      double out[15 * 9];
      premul(raw, 9, out);
      std::memcpy(jacobians[3], out, sizeof(out));
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// GNSS factors   RVI/factor/gnss_factor.cpp:9-212
// ---------------------------------------------------------------------------------------------
struct GnssFactor : CostFunction {
  int kind;
  double sat[3], vel[3], base[3], meas, lam, w;
  GnssFactor(int kind_, const double* rec) : kind(kind_) {
    for (int i = 0; i < 3; ++i) {
      sat[i] = rec[SWGN_GNSS_SAT_POS + i];
      vel[i] = rec[SWGN_GNSS_SAT_VEL + i];
      base[i] = rec[SWGN_GNSS_BASE_POS + i];
    }
    meas = rec[SWGN_GNSS_MEAS];
    lam = rec[SWGN_GNSS_LAM];
    w = rec[SWGN_GNSS_WEIGHT];
    num_residuals = 1;
    switch (kind) {
      case SWGN_GNSS_SPP_PSEUDORANGE: block_sizes = {7, 1}; break;
      case SWGN_GNSS_SPP_CARRIER: block_sizes = {7, 1, 1}; break;
      case SWGN_GNSS_RTK_CARRIER: block_sizes = {7, 1, 1}; break;
      case SWGN_GNSS_RTK_PSEUDORANGE: block_sizes = {7, 1}; break;
      case SWGN_GNSS_DOPPLER: block_sizes = {9, 1, 7}; break;
      case SWGN_GNSS_FIXED_INTEGER: block_sizes = {1, 1}; break;
    }
  }
  bool Evaluate(double const* const* p, double* residuals, double** J) const override {
    if (kind == SWGN_GNSS_FIXED_INTEGER) {                       // :85-96
      residuals[0] = w * ((p[1][0] - p[0][0]) - meas);
      if (J) {
        if (J[0]) J[0][0] = -w;
        if (J[1]) J[1][0] = w;
      }
      return true;
    }
    if (kind == SWGN_GNSS_DOPPLER) {                             // :174-212
      const double* vxyz = p[0];
      const double* dt = p[1];
      const double* xyz = p[2];
      double xg[3] = {xyz[0] + base[0], xyz[1] + base[1], xyz[2] + base[2]}, e[3];
      double rate = velocity_distance_rtk(xg, sat, vxyz, vel, e);
      residuals[0] = w * (rate + *dt + meas);
      if (J) {
        if (J[0]) {
          std::memset(J[0], 0, sizeof(double) * 9);
          J[0][0] = w * e[0];
          J[0][1] = w * e[1];
          J[0][2] = w * e[2];
        }
        if (J[1]) J[1][0] = w * 1;
        if (J[2]) {
          std::memset(J[2], 0, sizeof(double) * 7);
          double e2[3] = {xg[0] - sat[0], xg[1] - sat[1], xg[2] - sat[2]};
          double r = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
          for (int i = 0; i < 3; ++i) e2[i] /= r;
          double ev[3] = {vxyz[0] - vel[0], vxyz[1] - vel[1], vxyz[2] - vel[2]};
          // J2 = w * ev^T (I - e2 e2^T) / r
          for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += (w * ev[k]) * ((k == c ? 1.0 : 0.0) - e2[k] * e2[c]);
            J[2][c] = s / r;
          }
        }
      }
      return true;
    }
    const double* xyz = p[0];
    double xg[3] = {xyz[0] + base[0], xyz[1] + base[1], xyz[2] + base[2]}, e[3];
    double r1 = distance_rtk(xg, sat, e);
    int clk_slot = -1, n_slot = -1;
    switch (kind) {
      case SWGN_GNSS_SPP_PSEUDORANGE:                            // :9-39
        residuals[0] = w * (r1 + p[1][0] - meas);
        clk_slot = 1;
        break;
      case SWGN_GNSS_SPP_CARRIER:                                // :45-80
        residuals[0] = w * (r1 + p[1][0] - p[2][0] * lam - meas);
        clk_slot = 1;
        n_slot = 2;
        break;
      case SWGN_GNSS_RTK_CARRIER:                                // :105-138
        residuals[0] = w * (r1 - p[1][0] * lam - meas + p[2][0]);
        n_slot = 1;
        clk_slot = 2;
        break;
      case SWGN_GNSS_RTK_PSEUDORANGE:                            // :140-168
        residuals[0] = w * (r1 - meas + p[1][0]);
        clk_slot = 1;
        break;
    }
    if (J) {
      if (J[0]) {
        std::memset(J[0], 0, sizeof(double) * 7);
        J[0][0] = w * e[0];
        J[0][1] = w * e[1];
        J[0][2] = w * e[2];
      }
      if (clk_slot >= 0 && J[clk_slot]) J[clk_slot][0] = w;
      if (n_slot >= 0 && J[n_slot]) J[n_slot][0] = -w * lam;
    }
    return true;
  }
};

// ---------------------------------------------------------------------------------------------
// MarginalizationFactor   RVI/factor/marginalization_factor.cpp:410-446
// ---------------------------------------------------------------------------------------------
struct PriorFactor : CostFunction {
  int n;
  std::vector<int> idx;
  std::vector<int> x0_off;
  std::vector<double> x0, J0, r0;
  PriorFactor(int n_, const std::vector<int>& sizes, const std::vector<int>& idx_,
              const double* x0_, const double* J0_, const double* r0_)
      : n(n_), idx(idx_) {
    block_sizes = sizes;
    num_residuals = n;
    int tot = 0;
    for (int s : sizes) {
      x0_off.push_back(tot);
      tot += s;
    }
    x0.assign(x0_, x0_ + tot);
    J0.assign(J0_, J0_ + (size_t)n * n);
    r0.assign(r0_, r0_ + n);
  }
  bool Evaluate(double const* const* p, double* residuals, double** jacobians) const override {
    std::vector<double> dx(n, 0.0);
    for (size_t i = 0; i < block_sizes.size(); ++i) {
      int size = block_sizes[i];
      const double* x = p[i];
      const double* y = x0.data() + x0_off[i];
      if (size != 7) {
        for (int k = 0; k < size; ++k) dx[idx[i] + k] = x[k] - y[k];
      } else {
        for (int k = 0; k < 3; ++k) dx[idx[i] + k] = x[k] - y[k];
        Quat q = qmul(qinv(pose_q(y)), pose_q(x));
        double sgn = (q.w >= 0) ? 1.0 : -1.0;                    // :425-428
        dx[idx[i] + 3] = 2.0 * sgn * q.x;
        dx[idx[i] + 4] = 2.0 * sgn * q.y;
        dx[idx[i] + 5] = 2.0 * sgn * q.z;
      }
    }
    for (int r = 0; r < n; ++r) {
      double s = 0.0;
      for (int c = 0; c < n; ++c) s += J0[(size_t)r * n + c] * dx[c];
      residuals[r] = r0[r] + s;
    }
    if (jacobians) {
      for (size_t i = 0; i < block_sizes.size(); ++i) {
        if (!jacobians[i]) continue;
        int size = block_sizes[i], local = (size == 7) ? 6 : size;
        double* J = jacobians[i];
        for (int r = 0; r < n; ++r) {
          for (int c = 0; c < size; ++c) J[(size_t)r * size + c] = 0.0;
          for (int c = 0; c < local; ++c) J[(size_t)r * size + c] = J0[(size_t)r * n + idx[i] + c];
        }
      }
    }
    return true;
  }
};

// InitialBlackFactor <1;1>   RVI/factor/initial_factor.cpp:90-96
struct UnitFactor : CostFunction {
  double istd;
  explicit UnitFactor(double s) : istd(s) {
    block_sizes = {1};
    num_residuals = 1;
  }
  bool Evaluate(double const* const* p, double* residuals, double** jacobians) const override {
    residuals[0] = p[0][0] * istd;
    if (jacobians && jacobians[0]) jacobians[0][0] = 1 * istd;
    return true;
  }
};

}  // namespace

double dot_rtk(const double* a, const double* b, int n) {  // common_function.cpp:103-108
  double c = 0.0;
  while (--n >= 0) c += a[n] * b[n];
  return c;
}

double distance_rtk(const double* rr, const double* rs, double* e) {  // :126-134
  const double OMGE = 7.2921151467E-5, clight = 299792458.0;
  for (int i = 0; i < 3; ++i) e[i] = rr[i] - rs[i];
  double r = std::sqrt(dot_rtk(e, e, 3));
  for (int i = 0; i < 3; ++i) e[i] /= r;
  return r + OMGE * (rs[0] * rr[1] - rs[1] * rr[0]) / clight;
}

double velocity_distance_rtk(const double* rr, const double* rs, const double* vr,
                             const double* vs, double* e) {  // :411-421
  const double OMGE = 7.2921151467E-5, clight = 299792458.0;
  double ev[3];
  for (int i = 0; i < 3; ++i) e[i] = rr[i] - rs[i];
  double r = std::sqrt(dot_rtk(e, e, 3));
  for (int i = 0; i < 3; ++i) e[i] /= r;
  for (int i = 0; i < 3; ++i) ev[i] = vr[i] - vs[i];
  return dot_rtk(ev, e, 3) +
         OMGE / clight * (vs[1] * rr[0] + rs[1] * vr[0] - vs[0] * rr[1] - rs[0] * vr[1]);
}

double varerr2(double el, double dt, double mea_var) {  // gnss_factor.cpp:98-103
  const double clight = 299792458.0;
  double b = clight * 5e-12 * dt;
  double sinel = sinf(el);
  return (mea_var / sinel / sinel) + b * b;
}

CostFunction* make_projection_factor(const AppGlobals* g, const double uv[2]) {
  return new ProjectionFactor(g, uv);
}
CostFunction* make_imu_factor(const AppGlobals* g, const double* rec) {
  return new ImuFactor(g, rec);
}
CostFunction* make_gnss_factor(int kind, const double* rec) { return new GnssFactor(kind, rec); }
CostFunction* make_prior_factor(int n, const std::vector<int>& sizes, const std::vector<int>& idx,
                                const double* x0, const double* J0, const double* r0) {
  return new PriorFactor(n, sizes, idx, x0, J0, r0);
}
CostFunction* make_unit_factor(double istd) { return new UnitFactor(istd); }

}  // namespace oracle
