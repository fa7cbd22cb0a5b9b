// IMU pre-integration on the device: produces the constants of IMUFactor that the
// reference computes on the host at 400 Hz -- IntegrationBase::push_back / propagate / midPointIntegration /
// get_sqrtinfo, RVI/factor/integration_base.cpp:5-142.  One warp per factor, its 15x15 Jacobian and
// covariance in shared memory; every sample is one midpoint step: delta_p/q/v in registers (all lanes
// redundantly), the 13 non-zero 3x3 blocks of F and the 12 of V written by one lane each, then
// jacobian = F jacobian and covariance = F cov F' + V noise V' as lane-parallel 15-term dot products.
// sqrt_info = LLT(covariance^-1).L' is finished by lane 0 with the same partial-pivoting LU and
// column Cholesky the CPU path uses.  Compiled with -fmad=false like the other evaluation kernels.
#include "dev_common.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

constexpr int kWarpsPerCta = 4;
struct PreintSmem {
  double J[225], C[225], F[225], T[225], V[270], inv[225], L[225];
};

__device__ __forceinline__ void put33s(double* M, int ld, int r, int c, const double* B, double s) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[(r + i) * ld + c + j] = s * B[i * 3 + j];
}

}  // namespace

__global__ void __launch_bounds__(32 * kWarpsPerCta) k_preintegrate(int n_factors, const int32_t* begin, const double* samples,
                                                                    const double* bias, double an, double gn, double aw, double gw,
                                                                    double* records, int32_t* status) {
  extern __shared__ __align__(16) unsigned char raw_smem[];
  PreintSmem* all = reinterpret_cast<PreintSmem*>(raw_smem);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = blockIdx.x * kWarpsPerCta + wid;
  if (f >= n_factors) return;
  PreintSmem& S = all[wid];
  const int s0 = begin[f], n_samples = begin[f + 1] - s0;
  const double* sm = samples + (size_t)7 * s0;
  const double ba[3] = {bias[6 * f], bias[6 * f + 1], bias[6 * f + 2]};
  const double bg[3] = {bias[6 * f + 3], bias[6 * f + 4], bias[6 * f + 5]};
  double acc0[3] = {sm[1], sm[2], sm[3]}, gyr0[3] = {sm[4], sm[5], sm[6]};
  const double gyri[3] = {gyr0[0], gyr0[1], gyr0[2]};
  double gyrj[3] = {gyr0[0], gyr0[1], gyr0[2]};
  double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0}, sum_dt = 0.0;
  Quat dq = {1, 0, 0, 0};
  const double nd[6] = {an * an, gn * gn, an * an, gn * gn, aw * aw, gw * gw};
  for (int o = lane; o < 225; o += 32) {
    S.J[o] = (o / 15 == o % 15) ? 1.0 : 0.0;
    S.C[o] = 0.0;
    S.F[o] = 0.0;
  }
  for (int o = lane; o < 270; o += 32) S.V[o] = 0.0;
  __syncwarp();
  for (int s = 1; s < n_samples; ++s) {
    const double dt = sm[7 * s];
    const double acc1[3] = {sm[7 * s + 1], sm[7 * s + 2], sm[7 * s + 3]};
    const double gyr1[3] = {sm[7 * s + 4], sm[7 * s + 5], sm[7 * s + 6]};
    double a0[3], a1[3], w[3], un_acc0[3], un_acc1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gyrj[i] = gyr1[i];
      a0[i] = acc0[i] - ba[i];
      a1[i] = acc1[i] - ba[i];
      w[i] = 0.5 * (gyr0[i] + gyr1[i]) - bg[i];
    }
    qrot(dq, a0, un_acc0);
    const Quat rq = qmul(dq, Quat{1, w[0] * dt / 2, w[1] * dt / 2, w[2] * dt / 2});
    qrot(rq, a1, un_acc1);
    double ndp[3], ndv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double un_acc = 0.5 * (un_acc0[i] + un_acc1[i]);
      ndp[i] = dp[i] + dv[i] * dt + 0.5 * un_acc * dt * dt;
      ndv[i] = dv[i] + un_acc * dt;
    }
    // F and V blocks (integration_base.cpp:48-92), one lane per non-zero 3x3 block
    {
      double Rq[9], Rn[9], Rw[9], Ra0[9], Ra1[9], ImRw[9], RqRa0[9], RnRa1[9], RnRa1I[9], RqpRn[9], blk[9];
      const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      qtoR(dq, Rq);
      qtoR(rq, Rn);
      skew3(w, Rw);
      skew3(a0, Ra0);
      skew3(a1, Ra1);
#pragma unroll
      for (int i = 0; i < 9; ++i) ImRw[i] = I3[i] - Rw[i] * dt;
      m33_mul(Rq, Ra0, RqRa0);
      m33_mul(Rn, Ra1, RnRa1);
      m33_mul(RnRa1, ImRw, RnRa1I);
#pragma unroll
      for (int i = 0; i < 9; ++i) RqpRn[i] = Rq[i] + Rn[i];
      switch (lane) {
        case 0: put33s(S.F, 15, 0, 0, I3, 1.0); break;
        case 1:
#pragma unroll
          for (int i = 0; i < 9; ++i) blk[i] = -0.25 * RqRa0[i] * dt * dt + -0.25 * RnRa1I[i] * dt * dt;
          put33s(S.F, 15, 0, 3, blk, 1.0);
          break;
        case 2: put33s(S.F, 15, 0, 6, I3, dt); break;
        case 3: put33s(S.F, 15, 0, 9, RqpRn, -0.25 * dt * dt); break;
        case 4: put33s(S.F, 15, 0, 12, RnRa1, -0.25 * dt * dt * -dt); break;
        case 5: put33s(S.F, 15, 3, 3, ImRw, 1.0); break;
        case 6: put33s(S.F, 15, 3, 12, I3, -1.0 * dt); break;
        case 7:
#pragma unroll
          for (int i = 0; i < 9; ++i) blk[i] = -0.5 * RqRa0[i] * dt + -0.5 * RnRa1I[i] * dt;
          put33s(S.F, 15, 6, 3, blk, 1.0);
          break;
        case 8: put33s(S.F, 15, 6, 6, I3, 1.0); break;
        case 9: put33s(S.F, 15, 6, 9, RqpRn, -0.5 * dt); break;
        case 10: put33s(S.F, 15, 6, 12, RnRa1, -0.5 * dt * -dt); break;
        case 11: put33s(S.F, 15, 9, 9, I3, 1.0); break;
        case 12: put33s(S.F, 15, 12, 12, I3, 1.0); break;
        case 13: put33s(S.V, 18, 0, 0, Rq, 0.25 * dt * dt); break;
        case 14: put33s(S.V, 18, 0, 3, RnRa1, 0.25 * -1.0 * dt * dt * 0.5 * dt); break;
        case 15: put33s(S.V, 18, 0, 6, Rn, 0.25 * dt * dt); break;
        case 16: put33s(S.V, 18, 0, 9, RnRa1, 0.25 * -1.0 * dt * dt * 0.5 * dt); break;
        case 17: put33s(S.V, 18, 3, 3, I3, 0.5 * dt); break;
        case 18: put33s(S.V, 18, 3, 9, I3, 0.5 * dt); break;
        case 19: put33s(S.V, 18, 6, 0, Rq, 0.5 * dt); break;
        case 20: put33s(S.V, 18, 6, 3, RnRa1, 0.5 * -1.0 * dt * 0.5 * dt); break;
        case 21: put33s(S.V, 18, 6, 6, Rn, 0.5 * dt); break;
        case 22: put33s(S.V, 18, 6, 9, RnRa1, 0.5 * -1.0 * dt * 0.5 * dt); break;
        case 23: put33s(S.V, 18, 9, 12, I3, dt); break;
        case 24: put33s(S.V, 18, 12, 15, I3, dt); break;
        default: break;
      }
    }
    __syncwarp();
    // jacobian = F * jacobian
    for (int o = lane; o < 225; o += 32) {
      const int i = o / 15, j = o - i * 15;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 15; ++k) acc += S.F[i * 15 + k] * S.J[k * 15 + j];
      S.T[o] = acc;
    }
    __syncwarp();
    for (int o = lane; o < 225; o += 32) S.J[o] = S.T[o];
    __syncwarp();
    // covariance = F cov F' + V noise V'
    for (int o = lane; o < 225; o += 32) {
      const int i = o / 15, j = o - i * 15;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 15; ++k) acc += S.F[i * 15 + k] * S.C[k * 15 + j];
      S.T[o] = acc;
    }
    __syncwarp();
    for (int o = lane; o < 225; o += 32) {
      const int i = o / 15, j = o - i * 15;
      double fcf = 0.0, vnv = 0.0;
#pragma unroll
      for (int k = 0; k < 15; ++k) fcf += S.T[i * 15 + k] * S.F[j * 15 + k];
#pragma unroll
      for (int k = 0; k < 18; ++k) vnv += (S.V[i * 18 + k] * nd[k / 3]) * S.V[j * 18 + k];
      S.C[o] = fcf + vnv;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      dp[i] = ndp[i];
      dv[i] = ndv[i];
      acc0[i] = acc1[i];
      gyr0[i] = gyr1[i];
    }
    dq = qnormalized(rq);
    sum_dt += dt;
  }
  double* rec = records + (size_t)SWGN_IMU_STRIDE * f;
  for (int o = lane; o < SWGN_IMU_STRIDE; o += 32) rec[o] = 0.0;
  __syncwarp();
  if (lane == 0) {
    for (int i = 0; i < 3; ++i) {
      rec[SWGN_IMU_DELTA_P + i] = dp[i];
      rec[SWGN_IMU_DELTA_V + i] = dv[i];
      rec[SWGN_IMU_LIN_BA + i] = ba[i];
      rec[SWGN_IMU_LIN_BG + i] = bg[i];
      rec[SWGN_IMU_GYRI + i] = gyri[i];
      rec[SWGN_IMU_GYRJ + i] = gyrj[i];
    }
    rec[SWGN_IMU_DELTA_Q] = dq.x;
    rec[SWGN_IMU_DELTA_Q + 1] = dq.y;
    rec[SWGN_IMU_DELTA_Q + 2] = dq.z;
    rec[SWGN_IMU_DELTA_Q + 3] = dq.w;
    rec[SWGN_IMU_SUM_DT] = sum_dt;
  }
  for (int o = lane; o < 225; o += 32) rec[SWGN_IMU_JACOBIAN + o] = S.J[o];
  // get_sqrtinfo: partial-pivoting LU inverse, then LLT (lower) of the inverse, sqrt_info = L'
  int ok = 1;
  if (lane == 0) {
    double* lu = S.T;
    int piv[15];
    for (int o = 0; o < 225; ++o) lu[o] = S.C[o];
    for (int i = 0; i < 15; ++i) piv[i] = i;
    for (int k = 0; k < 15 && ok; ++k) {
      int p = k;
      double best = fabs(lu[k * 15 + k]);
      for (int i = k + 1; i < 15; ++i)
        if (fabs(lu[i * 15 + k]) > best) {
          best = fabs(lu[i * 15 + k]);
          p = i;
        }
      if (best == 0.0 || !finite_d(best)) { ok = 0; break; }
      if (p != k) {
        for (int j = 0; j < 15; ++j) {
          const double t = lu[k * 15 + j];
          lu[k * 15 + j] = lu[p * 15 + j];
          lu[p * 15 + j] = t;
        }
        const int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
      }
      for (int i = k + 1; i < 15; ++i) {
        lu[i * 15 + k] /= lu[k * 15 + k];
        const double fct = lu[i * 15 + k];
        for (int j = k + 1; j < 15; ++j) lu[i * 15 + j] -= fct * lu[k * 15 + j];
      }
    }
    if (ok) {
      double x[15];
      for (int c = 0; c < 15; ++c) {
        for (int i = 0; i < 15; ++i) x[i] = (piv[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i < 15; ++i)
          for (int j = 0; j < i; ++j) x[i] -= lu[i * 15 + j] * x[j];
        for (int i = 14; i >= 0; --i) {
          for (int j = i + 1; j < 15; ++j) x[i] -= lu[i * 15 + j] * x[j];
          x[i] /= lu[i * 15 + i];
        }
        for (int i = 0; i < 15; ++i) S.inv[i * 15 + c] = x[i];
      }
      for (int o = 0; o < 225; ++o) S.L[o] = 0.0;
      for (int j = 0; j < 15 && ok; ++j) {
        double x0 = S.inv[j * 15 + j];
        for (int p = 0; p < j; ++p) x0 -= S.L[j * 15 + p] * S.L[j * 15 + p];
        if (!(x0 > 0.0)) { ok = 0; break; }
        x0 = sqrt(x0);
        S.L[j * 15 + j] = x0;
        for (int i = j + 1; i < 15; ++i) {
          double sacc = S.inv[i * 15 + j];
          for (int p = 0; p < j; ++p) sacc -= S.L[i * 15 + p] * S.L[j * 15 + p];
          S.L[i * 15 + j] = sacc / x0;
        }
      }
    }
    status[f] = ok ? 0 : 1;
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  __syncwarp();
  if (ok)
    for (int o = lane; o < 225; o += 32) rec[SWGN_IMU_SQRT_INFO + o] = S.L[(o % 15) * 15 + o / 15];
}

cudaError_t launch_preintegrate(int n_factors, const int32_t* begin, const double* samples, const double* bias, const double* noise4,
                                double* records, int32_t* status, cudaStream_t s) {
  const size_t dyn = sizeof(PreintSmem) * kWarpsPerCta;
  cudaError_t e = cudaFuncSetAttribute(k_preintegrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  const int grid = (n_factors + kWarpsPerCta - 1) / kWarpsPerCta;
  k_preintegrate<<<grid, 32 * kWarpsPerCta, dyn, s>>>(n_factors, begin, samples, bias, noise4[0], noise4[1], noise4[2], noise4[3], records, status);
  return cudaGetLastError();
}

}  // namespace swgn
