// TEST INFRASTRUCTURE -- see oracle_core.h.  Restatement of the IMU pre-integration that produces the
// constants of IMUFactor:
//   IntegrationBase::IntegrationBase      RVI/factor/integration_base.cpp:5-23   (noise, identity Jacobian)
//   push_back / propagate                 :25-28, 115-142   (delta_q normalised after every step)
//   midPointIntegration                   :32-101           (F 15x15, V 15x18, jacobian = F jacobian,
//                                                            covariance = F cov F' + V noise V')
//   get_sqrtinfo                          :105-113          (LLT(covariance.inverse()).matrixL().transpose())
// Eigen semantics kept: quaternion * vector is _transformVector (valid for the un-normalised
// result_delta_q), toRotationMatrix is evaluated on the un-normalised quaternion as written, the
// 15x15 fixed-size inverse() is a partial-pivoting LU, LLT reads the lower triangle.
// PARITY UNPINNED by the reference (no test or fixture for IntegrationBase); cross-checked against
// the independent implementation inside the synthetic generator (tests/test_preintegration.py).
#include "oracle_core.h"

namespace oracle {

static void put33(Mat& M, int r, int c, const double* B, double s) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M(r + i, c + j) = s * B[i * 3 + j];
}

// samples: 7 per sample (dt, acc[3], gyr[3]); sample 0 is (acc_0, gyr_0) of the constructor.
// bias: linearized_ba[3], linearized_bg[3]; noise: ACC_N, GYR_N, ACC_W, GYR_W.
// record: SWGN_IMU_STRIDE doubles.  Returns false when the covariance cannot be inverted / factored.
bool preintegrate(int n_samples, const double* samples, const double* bias, const double* noise4, double* record) {
  const double* ba = bias;
  const double* bg = bias + 3;
  double acc0[3] = {samples[1], samples[2], samples[3]}, gyr0[3] = {samples[4], samples[5], samples[6]};
  double gyri[3] = {gyr0[0], gyr0[1], gyr0[2]}, gyrj[3] = {gyr0[0], gyr0[1], gyr0[2]};
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, sum_dt = 0.0;
  Quat dq = {1, 0, 0, 0};
  Mat jac = Mat::Identity(15), cov(15, 15), N(18, 18);
  const double nd[6] = {noise4[0] * noise4[0], noise4[1] * noise4[1], noise4[0] * noise4[0],
                        noise4[1] * noise4[1], noise4[2] * noise4[2], noise4[3] * noise4[3]};
  for (int b = 0; b < 6; ++b)
    for (int i = 0; i < 3; ++i) N(3 * b + i, 3 * b + i) = nd[b];
  for (int s = 1; s < n_samples; ++s) {
    const double dt = samples[7 * s];
    const double* acc1 = samples + 7 * s + 1;
    const double* gyr1 = samples + 7 * s + 4;
    for (int i = 0; i < 3; ++i) gyrj[i] = gyr1[i];
    double a0[3], a1[3], w[3], un_acc0[3], un_acc1[3];
    for (int i = 0; i < 3; ++i) {
      a0[i] = acc0[i] - ba[i];
      a1[i] = acc1[i] - ba[i];
      w[i] = 0.5 * (gyr0[i] + gyr1[i]) - bg[i];
    }
    qrot(dq, a0, un_acc0);
    const Quat rq = qmul(dq, Quat{1, w[0] * dt / 2, w[1] * dt / 2, w[2] * dt / 2});
    qrot(rq, a1, un_acc1);
    double un_acc[3], ndp[3], ndv[3];
    for (int i = 0; i < 3; ++i) {
      un_acc[i] = 0.5 * (un_acc0[i] + un_acc1[i]);
      ndp[i] = dp[i] + dv[i] * dt + 0.5 * un_acc[i] * dt * dt;
      ndv[i] = dv[i] + un_acc[i] * dt;
    }
    // Jacobian and covariance :48-97
    double Rq[9], Rn[9], Rw[9], Ra0[9], Ra1[9], I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    qtoR(dq, Rq);
    qtoR(rq, Rn);
    skew(w, Rw);
    skew(a0, Ra0);
    skew(a1, Ra1);
    double ImRw[9], RqRa0[9], RnRa1[9], RnRa1I[9], RqpRn[9];
    for (int i = 0; i < 9; ++i) ImRw[i] = I3[i] - Rw[i] * dt;
    auto mul33 = [](const double* A, const double* B, double* C) {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    };
    mul33(Rq, Ra0, RqRa0);
    mul33(Rn, Ra1, RnRa1);
    mul33(RnRa1, ImRw, RnRa1I);
    for (int i = 0; i < 9; ++i) RqpRn[i] = Rq[i] + Rn[i];
    Mat F(15, 15), V(15, 18);
    double blk[9];
    put33(F, 0, 0, I3, 1.0);
    for (int i = 0; i < 9; ++i) blk[i] = -0.25 * RqRa0[i] * dt * dt + -0.25 * RnRa1I[i] * dt * dt;
    put33(F, 0, 3, blk, 1.0);
    put33(F, 0, 6, I3, dt);
    put33(F, 0, 9, RqpRn, -0.25 * dt * dt);
    put33(F, 0, 12, RnRa1, -0.25 * dt * dt * -dt);
    put33(F, 3, 3, ImRw, 1.0);
    put33(F, 3, 12, I3, -1.0 * dt);
    for (int i = 0; i < 9; ++i) blk[i] = -0.5 * RqRa0[i] * dt + -0.5 * RnRa1I[i] * dt;
    put33(F, 6, 3, blk, 1.0);
    put33(F, 6, 6, I3, 1.0);
    put33(F, 6, 9, RqpRn, -0.5 * dt);
    put33(F, 6, 12, RnRa1, -0.5 * dt * -dt);
    put33(F, 9, 9, I3, 1.0);
    put33(F, 12, 12, I3, 1.0);
    put33(V, 0, 0, Rq, 0.25 * dt * dt);
    put33(V, 0, 3, RnRa1, 0.25 * -1.0 * dt * dt * 0.5 * dt);
    put33(V, 0, 6, Rn, 0.25 * dt * dt);
    put33(V, 0, 9, RnRa1, 0.25 * -1.0 * dt * dt * 0.5 * dt);
    put33(V, 3, 3, I3, 0.5 * dt);
    put33(V, 3, 9, I3, 0.5 * dt);
    put33(V, 6, 0, Rq, 0.5 * dt);
    put33(V, 6, 3, RnRa1, 0.5 * -1.0 * dt * 0.5 * dt);
    put33(V, 6, 6, Rn, 0.5 * dt);
    put33(V, 6, 9, RnRa1, 0.5 * -1.0 * dt * 0.5 * dt);
    put33(V, 9, 12, I3, dt);
    put33(V, 12, 15, I3, dt);
    jac = matmul(F, jac);
    Mat FCF = matmul(matmul(F, cov), transpose(F));
    Mat VNV = matmul(matmul(V, N), transpose(V));
    for (int i = 0; i < 225; ++i) cov.a[i] = FCF.a[i] + VNV.a[i];
    for (int i = 0; i < 3; ++i) {
      dp[i] = ndp[i];
      dv[i] = ndv[i];
      acc0[i] = acc1[i];
      gyr0[i] = gyr1[i];
    }
    dq = qnormalized(rq);
    sum_dt += dt;
  }
  std::fill(record, record + SWGN_IMU_STRIDE, 0.0);
  for (int i = 0; i < 3; ++i) {
    record[SWGN_IMU_DELTA_P + i] = dp[i];
    record[SWGN_IMU_DELTA_V + i] = dv[i];
    record[SWGN_IMU_LIN_BA + i] = ba[i];
    record[SWGN_IMU_LIN_BG + i] = bg[i];
    record[SWGN_IMU_GYRI + i] = gyri[i];
    record[SWGN_IMU_GYRJ + i] = gyrj[i];
  }
  record[SWGN_IMU_DELTA_Q] = dq.x;
  record[SWGN_IMU_DELTA_Q + 1] = dq.y;
  record[SWGN_IMU_DELTA_Q + 2] = dq.z;
  record[SWGN_IMU_DELTA_Q + 3] = dq.w;
  record[SWGN_IMU_SUM_DT] = sum_dt;
  for (int i = 0; i < 225; ++i) record[SWGN_IMU_JACOBIAN + i] = jac.a[i];
  // get_sqrtinfo :105-113
  Mat inv;
  if (!inverse_lu(cov, &inv)) return false;
  // Eigen::LLT (lower): L L' = A from the lower triangle, column by column
  Mat L(15, 15);
  for (int j = 0; j < 15; ++j) {
    double x = inv(j, j);
    for (int p = 0; p < j; ++p) x -= L(j, p) * L(j, p);
    if (!(x > 0.0)) return false;
    x = std::sqrt(x);
    L(j, j) = x;
    for (int i = j + 1; i < 15; ++i) {
      double s = inv(i, j);
      for (int p = 0; p < j; ++p) s -= L(i, p) * L(j, p);
      L(i, j) = s / x;
    }
  }
  for (int i = 0; i < 15; ++i)
    for (int j = 0; j < 15; ++j) record[SWGN_IMU_SQRT_INFO + i * 15 + j] = L(j, i);
  // the covariance itself is not part of the factor record; expose it after the record for tests
  return true;
}

}  // namespace oracle
