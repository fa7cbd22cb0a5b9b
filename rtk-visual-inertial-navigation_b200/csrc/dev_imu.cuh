// IMU pre-integration factor, one warp per evaluation: the residual of IMUFactor::Evaluate
// (RVI/factor/imu_factor.cpp:5-101, integration_base.cpp:144-174) and its raw 15 x 30 Jacobian
// (columns: pose_i 6 | speed-bias_i 9 | pose_j 6 | speed-bias_j 9, tangent space) before the
// sqrt_info pre-multiplication.  Shared by k_eval (IMUFactor) and k_chain (IMUGNSSFactor, whose
// IMUFactor::Evaluate2, imu_factor.cpp:103-193, is the same arithmetic in a 15x15|15x15 layout).
#pragma once
#include "dev_common.cuh"
#include "../../include/swgn.h"

namespace swgn {

__device__ __forceinline__ void put33(double* raw, int r0, int c0, const double* B, double s) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) raw[(r0 + i) * 30 + c0 + j] = s * B[i * 3 + j];
}
__device__ __forceinline__ void qleft_br(const Quat& q, double* M) {  // w I + [v]x
  const double vv[3] = {q.x, q.y, q.z};
  skew3(vv, M);
  M[0] += q.w; M[4] += q.w; M[8] += q.w;
}
__device__ __forceinline__ void qright_br(const Quat& q, double* M) {  // w I - [v]x
  const double vv[3] = {q.x, q.y, q.z};
  double S[9];
  skew3(vv, S);
#pragma unroll
  for (int i = 0; i < 9; ++i) M[i] = -S[i];
  M[0] += q.w; M[4] += q.w; M[8] += q.w;
}


// rec: device IMU record (IMU_DEV_STRIDE).  Every lane computes the raw residual; lanes 0..14
// return row `lane` of sqrt_info * raw_r (other lanes 0).  With want_jac the warp also fills
// raw[15*30] (shared memory) and synchronises.
__device__ __forceinline__ double imu_residual_raw(const double* rec, const double* Pbg, const double* G, const double* pi,
                                                   const double* si, const double* pj, const double* sj, double* raw,
                                                   bool want_jac, int lane) {
  enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };
  const Quat Qi = pose_q(pi), Qj = pose_q(pj);
  const Quat dq = {rec[SWGN_IMU_DELTA_Q + 3], rec[SWGN_IMU_DELTA_Q], rec[SWGN_IMU_DELTA_Q + 1], rec[SWGN_IMU_DELTA_Q + 2]};
  const double dt = rec[SWGN_IMU_SUM_DT];
  const double* dp_dba = rec + IMU_DEV_BLOCKS;
  const double* dp_dbg = dp_dba + 9;
  const double* dq_dbg = dp_dba + 18;
  const double* dv_dba = dp_dba + 27;
  const double* dv_dbg = dp_dba + 36;
  double dba[3], dbg[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dba[k] = si[3 + k] - rec[SWGN_IMU_LIN_BA + k];
    dbg[k] = si[6 + k] - rec[SWGN_IMU_LIN_BG + k];
  }
  double th[3], t1[3], t2[3], cdv[3], cdp[3];
  m33_vec(dq_dbg, dbg, th);
  const Quat cdq = qmul(dq, Quat{1.0, th[0] / 2.0, th[1] / 2.0, th[2] / 2.0});
  m33_vec(dv_dba, dba, t1);
  m33_vec(dv_dbg, dbg, t2);
#pragma unroll
  for (int k = 0; k < 3; ++k) cdv[k] = rec[SWGN_IMU_DELTA_V + k] + t1[k] + t2[k];
  m33_vec(dp_dba, dba, t1);
  m33_vec(dp_dbg, dbg, t2);
#pragma unroll
  for (int k = 0; k < 3; ++k) cdp[k] = rec[SWGN_IMU_DELTA_P + k] + t1[k] + t2[k];
  const Quat Qi_inv = qinv(Qi);
  double QjPbg[3];
  qrot(Qj, Pbg, QjPbg);
  const double wi[3] = {rec[SWGN_IMU_GYRI] - si[6], rec[SWGN_IMU_GYRI + 1] - si[7], rec[SWGN_IMU_GYRI + 2] - si[8]};
  const double wj[3] = {rec[SWGN_IMU_GYRJ] - sj[6], rec[SWGN_IMU_GYRJ + 1] - sj[7], rec[SWGN_IMU_GYRJ + 2] - sj[8]};
  double Sw[9], wiPbg[3], wjPbg[3];
  skew3(wi, Sw);
  m33_vec(Sw, Pbg, wiPbg);
  skew3(wj, Sw);
  m33_vec(Sw, Pbg, wjPbg);
  double a[3], ra[3], bb[3], rb[3], QjwjPbg[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) a[k] = 0.5 * G[k] * dt * dt + ((pj[k] - pi[k]) - QjPbg[k]) - si[k] * dt;
  qrot(Qi_inv, a, ra);
  qrot(Qj, wjPbg, QjwjPbg);
#pragma unroll
  for (int k = 0; k < 3; ++k) bb[k] = G[k] * dt + (sj[k] - QjwjPbg[k]) - si[k];
  qrot(Qi_inv, bb, rb);
  double raw_r[15];
  const Quat qr = qmul(qinv(cdq), qmul(Qi_inv, Qj));
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    raw_r[O_P + k] = ra[k] - cdp[k] + Pbg[k] + wiPbg[k] * dt;
    raw_r[O_V + k] = rb[k] - cdv[k] + wiPbg[k];
    raw_r[O_BA + k] = sj[3 + k] - si[3 + k];
    raw_r[O_BG + k] = sj[6 + k] - si[6 + k];
  }
  raw_r[O_R] = 2 * qr.x;
  raw_r[O_R + 1] = 2 * qr.y;
  raw_r[O_R + 2] = 2 * qr.z;
  const double* sqrt_info = rec + IMU_DEV_SQRT;
  double rk = 0.0;
  if (lane < 15) {
#pragma unroll
    for (int m = 0; m < 15; ++m) rk += sqrt_info[lane * 15 + m] * raw_r[m];
  }
  if (!want_jac) return rk;
  // raw Jacobian 15 x 30 (columns: pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9) in the warp's scratch
  for (int k = lane; k < 15 * 30; k += 32) raw[k] = 0.0;
  __syncwarp();
  if (lane < 4) {
    double Ri_inv[9], M[9];
    qtoR(Qi_inv, Ri_inv);
    double SPbg[9];
    skew3(Pbg, SPbg);
    if (lane == 0) {  // d/d pose_i   imu_factor.cpp:47-60
      put33(raw, O_P, 0, Ri_inv, -1.0);
      skew3(ra, M);
      put33(raw, O_P, 3, M, 1.0);
      const Quat ql = qmul(qinv(Qj), Qi);
      double L[9], Rr[9], LR[9];
      qleft_br(ql, L);
      qright_br(cdq, Rr);
      m33_mul(L, Rr, LR);
      const double vl[3] = {ql.x, ql.y, ql.z}, vr[3] = {cdq.x, cdq.y, cdq.z};
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) M[r * 3 + cc] = -(vl[r] * (-vr[cc]) + LR[r * 3 + cc]);
      put33(raw, O_R, 3, M, 1.0);
      skew3(rb, M);
      put33(raw, O_V, 3, M, 1.0);
    } else if (lane == 1) {  // d/d speed-bias_i   :61-75
      put33(raw, O_P, 6, Ri_inv, -dt);
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = -dp_dba[k];
      put33(raw, O_P, 9, M, 1.0);
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = -dp_dbg[k] + SPbg[k] * dt;
      put33(raw, O_P, 12, M, 1.0);
      const Quat q3 = qmul(qmul(qinv(Qj), Qi), dq);
      double L[9], LB[9];
      qleft_br(q3, L);
      m33_mul(L, dq_dbg, LB);
      put33(raw, O_R, 12, LB, -1.0);
      put33(raw, O_V, 6, Ri_inv, -1.0);
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = -dv_dba[k];
      put33(raw, O_V, 9, M, 1.0);
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = -dv_dbg[k] + SPbg[k];
      put33(raw, O_V, 12, M, 1.0);
      const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      put33(raw, O_BA, 9, I3, -1.0);
      put33(raw, O_BG, 12, I3, -1.0);
    } else if (lane == 2) {  // d/d pose_j   :76-86
      double Rj[9], RiRj[9];
      qtoR(Qj, Rj);
      put33(raw, O_P, 15, Ri_inv, 1.0);
      m33_mul(Ri_inv, Rj, RiRj);
      m33_mul(RiRj, SPbg, M);
      put33(raw, O_P, 18, M, 1.0);
      const Quat q3 = qmul(qmul(qinv(cdq), qinv(Qi)), Qj);
      double L[9];
      qleft_br(q3, L);
      put33(raw, O_R, 18, L, 1.0);
      double S2[9];
      skew3(wjPbg, S2);
      m33_mul(RiRj, S2, M);
      put33(raw, O_V, 18, M, 1.0);
    } else {  // d/d speed-bias_j   :87-96
      double Rj[9], RiRj[9];
      qtoR(Qj, Rj);
      put33(raw, O_V, 21, Ri_inv, 1.0);
      m33_mul(Ri_inv, Rj, RiRj);
      m33_mul(RiRj, SPbg, M);
      put33(raw, O_V, 27, M, -1.0);
      const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      put33(raw, O_BA, 24, I3, 1.0);
      put33(raw, O_BG, 27, I3, 1.0);
    }
  }
  __syncwarp();
  return rk;
}

}  // namespace swgn
