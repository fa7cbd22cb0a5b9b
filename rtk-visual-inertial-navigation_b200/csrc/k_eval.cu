// K1/K2: fused residual + Jacobian evaluation of every factor type of one window in one launch,
// followed (same launch) by the gradient J^T r, the dogleg diagonal sqrt(clamp(|J_col|^2)) and the
// cost.  Replaces ProgramEvaluator::Evaluate (CERES/internal/ceres/program_evaluator.h:139-286)
// -> ResidualBlock::Evaluate (residual_block.cc:69-199) -> the app's CostFunction::Evaluate:
//   projection_factor   RVI/factor/projection_factor.cpp:13-65   (+ CauchyLoss, corrector.cc:42-156)
//   IMUFactor           RVI/factor/imu_factor.cpp:5-101, integration_base.cpp:144-174
//   GNSS factors        RVI/factor/gnss_factor.cpp:9-212, RVI/gnss/src/common_function.cpp:126-134,411-421
//   MarginalizationFactor RVI/factor/marginalization_factor.cpp:410-446
//   InitialBlackFactor  RVI/factor/initial_factor.cpp:90-96
// Jacobians are written directly in tangent space (PoseLocalParameterization::ComputeJacobian is
// [I6; 0], pose_local_parameterization.cpp:21-27, so the 7th global column is simply dropped).
#include <float.h>

#include "dev_common.cuh"
#include "dev_imu.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

#ifndef SWGN_EVAL_CTAS
#define SWGN_EVAL_CTAS 2  // CTAs per SM the register budget is tuned for (2: 128 registers; 3: 80 registers + 512 B of spills measured 3 % slower)
#endif
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kImuScratch = 15 * 30 + 16;

struct Globals {
  double Pbg[3], G[3], W[4], cauchy_a;
};

// ---------------------------------------------------------------------------------------------
// CauchyLoss + Corrector: rho'' < 0 always, so the corrector takes its alpha = 0 branch
// (corrector.cc:82-86): residual and Jacobian are scaled by sqrt(rho').
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cauchy(double a, double s, double* half_rho0, double* sqrt_rho1) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  *half_rho0 = 0.5 * (b * log(sum));
  *sqrt_rho1 = sqrt(fmax(DBL_MIN, inv));
}

// one projection factor per thread
__device__ __forceinline__ void eval_proj(const Win& v, const Globals& gl, int i, const double* x, double* J,
                                          double* R, bool full, bool with_fixed, double& cost, double& fixed,
                                          int& bad) {
  const int32_t* t = v.I(I_PROJ) + 8 * i;
  const int res_off = t[6];
  if (res_off < 0 && !with_fixed) return;
  const double* pj = x + t[0];
  const double* ex = x + t[1];
  const double* X = x + t[2];
  const double* uv = v.C(C_PROJ_UV) + 2 * i;
  const Quat Qj = pose_q(pj), qic = pose_q(ex);
  const double d[3] = {X[0] - pj[0], X[1] - pj[1], X[2] - pj[2]};
  double pts_imu[3], tt[3], pc[3];
  qrot(qinv(Qj), d, pts_imu);
  tt[0] = pts_imu[0] + gl.Pbg[0] - ex[0];
  tt[1] = pts_imu[1] + gl.Pbg[1] - ex[1];
  tt[2] = pts_imu[2] + gl.Pbg[2] - ex[2];
  qrot(qinv(qic), tt, pc);
  const double dep = pc[2];
  const double e0 = pc[0] / dep - uv[0], e1 = pc[1] / dep - uv[1];
  double r0 = gl.W[0] * e0 + gl.W[1] * e1, r1 = gl.W[2] * e0 + gl.W[3] * e1;
  if (!finite_d(r0) || !finite_d(r1)) bad = 1;
  const double sq = r0 * r0 + r1 * r1;
  double c = 0.5 * sq, scale = 1.0;
  if (gl.cauchy_a > 0.0) cauchy(gl.cauchy_a, sq, &c, &scale);
  if (res_off < 0) {
    fixed += c;
    return;
  }
  cost += c;
  if (!full) return;
  R[res_off] = r0 * scale;
  R[res_off + 1] = r1 * scale;
  double Rj[9], ric[9];
  qtoR(Qj, Rj);
  qtoR(qic, ric);
  const double id = 1.0 / dep, id2 = 1.0 / (dep * dep);
  const double red0[6] = {id, 0.0, -pc[0] * id2, 0.0, id, -pc[1] * id2};
  double red[6];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    red[j] = (gl.W[0] * red0[j] + gl.W[1] * red0[3 + j]) * scale;
    red[3 + j] = (gl.W[2] * red0[j] + gl.W[3] * red0[3 + j]) * scale;
  }
  // T = red * ric^T (2x3)
  double T[6];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
      T[r * 3 + cc] = red[r * 3] * ric[cc * 3] + red[r * 3 + 1] * ric[cc * 3 + 1] + red[r * 3 + 2] * ric[cc * 3 + 2];
  // JX = T * Rj^T
  double JX[6];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
      JX[r * 3 + cc] = T[r * 3] * Rj[cc * 3] + T[r * 3 + 1] * Rj[cc * 3 + 1] + T[r * 3 + 2] * Rj[cc * 3 + 2];
  if (t[3] >= 0) {  // pose: [-T Rj^T , T [pts_imu]x]
    double* Jp = J + t[3];
    double S[9];
    skew3(pts_imu, S);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const double jv = -JX[r * 3 + cc];
        const double jr = T[r * 3] * S[cc] + T[r * 3 + 1] * S[3 + cc] + T[r * 3 + 2] * S[6 + cc];
        if (!finite_d(jv) || !finite_d(jr)) bad = 1;
        Jp[r * 6 + cc] = jv;
        Jp[r * 6 + 3 + cc] = jr;
      }
  }
  if (t[4] >= 0) {  // camera extrinsic: [-T , red [pc]x]
    double* Je = J + t[4];
    double S[9];
    skew3(pc, S);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        Je[r * 6 + cc] = -T[r * 3 + cc];
        Je[r * 6 + 3 + cc] = red[r * 3] * S[cc] + red[r * 3 + 1] * S[3 + cc] + red[r * 3 + 2] * S[6 + cc];
      }
  }
  if (t[5] >= 0) {
    double* Jl = J + t[5];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (!finite_d(JX[k])) bad = 1;
      Jl[k] = JX[k];
    }
  }
}

// ---- GNSS scalar factors, one per thread -----------------------------------------------------
__device__ __forceinline__ double dot3_rtk(const double* a, const double* b) {  // high index first
  double c = 0.0;
  c += a[2] * b[2];
  c += a[1] * b[1];
  c += a[0] * b[0];
  return c;
}
#define SWGN_OMGE 7.2921151467E-5
#define SWGN_CLIGHT 299792458.0

__device__ __forceinline__ void eval_gnss(const Win& v, int i, const double* x, double* J, double* R, bool full,
                                          bool with_fixed, double& cost, double& fixed, int& bad) {
  const int32_t* t = v.I(I_GNSS) + 8 * i;
  const int res_off = t[7];
  if (res_off < 0 && !with_fixed) return;
  const int kind = t[0];
  const double* rec = v.C(C_GNSS) + (size_t)GNSS_DEV_STRIDE * i;
  const double *sat = rec, *vel = rec + 3, *base = rec + 6;
  const double meas = rec[9], lam = rec[10], w = rec[11];
  const double* p0 = x + t[1];
  const double* p1 = x + t[2];
  const double* p2 = x + t[3];
  double r;
  double j0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, j1 = 0.0, j2[6] = {0, 0, 0, 0, 0, 0};
  int n0 = 6, n2 = 1;  // tangent widths of parameter 0 and 2
  if (kind == SWGN_GNSS_FIXED_INTEGER) {
    r = w * ((p1[0] - p0[0]) - meas);
    j0[0] = -w;
    j1 = w;
    n0 = 1;
  } else if (kind == SWGN_GNSS_DOPPLER) {
    const double xg[3] = {p2[0] + base[0], p2[1] + base[1], p2[2] + base[2]};
    double e[3] = {xg[0] - sat[0], xg[1] - sat[1], xg[2] - sat[2]};
    const double rr = sqrt(dot3_rtk(e, e));
    e[0] /= rr; e[1] /= rr; e[2] /= rr;
    const double ev[3] = {p0[0] - vel[0], p0[1] - vel[1], p0[2] - vel[2]};
    const double rate = dot3_rtk(ev, e) +
                        SWGN_OMGE / SWGN_CLIGHT * (vel[1] * xg[0] + sat[1] * p0[0] - vel[0] * xg[1] - sat[0] * p0[1]);
    r = w * (rate + p1[0] + meas);
    n0 = 9;
    n2 = 6;
    j0[0] = w * e[0]; j0[1] = w * e[1]; j0[2] = w * e[2];
    j1 = w;
    // d/dp: w ev^T (I - e e^T) / |.|   (gnss_factor.cpp:199-207)
    const double d2[3] = {xg[0] - sat[0], xg[1] - sat[1], xg[2] - sat[2]};
    const double r2 = sqrt(d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]);
    const double e2[3] = {d2[0] / r2, d2[1] / r2, d2[2] / r2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += (w * ev[k]) * ((k == c ? 1.0 : 0.0) - e2[k] * e2[c]);
      j2[c] = s / r2;
    }
  } else {
    const double xg[3] = {p0[0] + base[0], p0[1] + base[1], p0[2] + base[2]};
    double e[3] = {xg[0] - sat[0], xg[1] - sat[1], xg[2] - sat[2]};
    const double rr = sqrt(dot3_rtk(e, e));
    e[0] /= rr; e[1] /= rr; e[2] /= rr;
    const double r1 = rr + SWGN_OMGE * (sat[0] * xg[1] - sat[1] * xg[0]) / SWGN_CLIGHT;
    j0[0] = w * e[0]; j0[1] = w * e[1]; j0[2] = w * e[2];
    if (kind == SWGN_GNSS_SPP_PSEUDORANGE) {
      r = w * (r1 + p1[0] - meas);
      j1 = w;
    } else if (kind == SWGN_GNSS_SPP_CARRIER) {
      r = w * (r1 + p1[0] - p2[0] * lam - meas);
      j1 = w;
      j2[0] = -w * lam;
    } else if (kind == SWGN_GNSS_RTK_CARRIER) {
      r = w * (r1 - p1[0] * lam - meas + p2[0]);
      j1 = -w * lam;
      j2[0] = w;
    } else {  // SWGN_GNSS_RTK_PSEUDORANGE
      r = w * (r1 - meas + p1[0]);
      j1 = w;
    }
  }
  if (!finite_d(r)) bad = 1;
  const double c = 0.5 * r * r;
  if (res_off < 0) {
    fixed += c;
    return;
  }
  cost += c;
  if (!full) return;
  R[res_off] = r;
  if (t[4] >= 0)
    for (int k = 0; k < n0; ++k) J[t[4] + k] = j0[k];
  if (t[5] >= 0) J[t[5]] = j1;
  if (t[6] >= 0)
    for (int k = 0; k < n2; ++k) J[t[6] + k] = j2[k];
}

// ---- IMU factor, one warp per factor ----------------------------------------------------------
__device__ void eval_imu(const Win& v, const Globals& gl, int i, const double* x, double* J, double* R, bool full,
                         bool with_fixed, double* raw /* per-warp scratch */, double& cost, double& fixed, int& bad) {
  const int lane = threadIdx.x & 31;
  const int32_t* t = v.I(I_IMU) + 12 * i;
  const int res_off = t[8];
  if (res_off < 0 && !with_fixed) return;
  const double* rec = v.C(C_IMU) + (size_t)IMU_DEV_STRIDE * i;
  const bool want_jac = full && res_off >= 0;
  const double rk = imu_residual_raw(rec, gl.Pbg, gl.G, x + t[0], x + t[1], x + t[2], x + t[3], raw, want_jac, lane);
  if (lane < 15 && !finite_d(rk)) bad = 1;
  const double c = 0.5 * warp_sum(lane < 15 ? rk * rk : 0.0);
  if (res_off < 0) {
    if (lane == 0) fixed += c;
    return;
  }
  if (lane == 0) cost += c;
  if (!full) return;
  if (lane < 15) R[res_off + lane] = rk;
  const double* sqrt_info = rec + IMU_DEV_SQRT;
  // J = sqrt_info * raw, scattered to the four cells
  for (int o = lane; o < 15 * 30; o += 32) {
    const int k = o / 30, cc = o - k * 30;
    int p, lc, ls;
    if (cc < 6) { p = 0; lc = cc; ls = 6; }
    else if (cc < 15) { p = 1; lc = cc - 6; ls = 9; }
    else if (cc < 21) { p = 2; lc = cc - 15; ls = 6; }
    else { p = 3; lc = cc - 21; ls = 9; }
    const int jo = t[4 + p];
    if (jo < 0) continue;
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < 15; ++m) s += sqrt_info[k * 15 + m] * raw[m * 30 + cc];
    if (!finite_d(s)) bad = 1;
    J[jo + k * ls + lc] = s;
  }
  __syncwarp();
}

}  // namespace

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, SWGN_EVAL_CTAS) k_eval(DeviceBatch b, int mode, int only_window) {
  __shared__ WinDesc sd;
  __shared__ double red[33];
  __shared__ double imu_scratch[kWarps][kImuScratch];
  extern __shared__ double dyn[];  // prior dx: max_prior_n doubles
  const int w = only_window >= 0 ? only_window : blockIdx.x;
  TRState* st = b.state + w;
  bool run;
  if (mode == EVAL_INIT || mode == EVAL_FORCE) run = true;
  else if (mode == EVAL_CANDIDATE) run = st->active && st->step_valid && !st->need_solve;
  else run = st->active && st->accepted;
  if (!run) return;
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  const bool full = mode != EVAL_CANDIDATE;
  const bool with_fixed = mode == EVAL_INIT || mode == EVAL_FORCE;
  const double* x = v.W(mode == EVAL_CANDIDATE ? W_XCAND : W_X);
  double* J = v.W(W_JAC);
  double* R = v.W(W_RES);
  Globals gl;
  {
    const double* g = v.C(C_GLOBALS);
#pragma unroll
    for (int k = 0; k < 3; ++k) { gl.Pbg[k] = g[k]; gl.G[k] = g[3 + k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) gl.W[k] = g[6 + k];
    gl.cauchy_a = g[10];
  }
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double cost = 0.0, fixed = 0.0;
  int bad = 0;

  for (int i = tid; i < d.n_proj; i += kThreads) eval_proj(v, gl, i, x, J, R, full, with_fixed, cost, fixed, bad);
  for (int i = tid; i < d.n_gnss; i += kThreads) eval_gnss(v, i, x, J, R, full, with_fixed, cost, fixed, bad);
  {
    const int32_t* ut = v.I(I_UNIT);
    const double* us = v.C(C_UNIT);
    for (int i = tid; i < d.n_unit; i += kThreads) {  // InitialBlackFactor
      const int res_off = ut[4 * i + 2];
      if (res_off < 0 && !with_fixed) continue;
      const double r = x[ut[4 * i]] * us[i];
      if (!finite_d(r)) bad = 1;
      if (res_off < 0) { fixed += 0.5 * r * r; continue; }
      cost += 0.5 * r * r;
      if (full) {
        R[res_off] = r;
        if (ut[4 * i + 1] >= 0) J[ut[4 * i + 1]] = us[i];
      }
    }
  }
  for (int i = wid; i < d.n_imu; i += kWarps) eval_imu(v, gl, i, x, J, R, full, with_fixed, imu_scratch[wid], cost, fixed, bad);

  // MarginalizationFactor: r = r0 + J0 (x [-] x0); J cells are column slices of J0
  for (int pr = 0; pr < d.n_prior; ++pr) {
    const int32_t* t = v.I(I_PRIOR) + 8 * pr;
    const int n = t[0], nblk = t[1], res_off = t[2], blk0 = t[3];
    if (res_off < 0 && !with_fixed) continue;
    const double* J0 = v.C(C_PRIOR_J) + t[4];
    const double* r0 = v.C(C_PRIOR_R0) + t[5];
    const int32_t* bt = v.I(I_PRIOR_BLK) + 6 * blk0;
    __syncthreads();
    for (int k = tid; k < n; k += kThreads) dyn[k] = 0.0;
    __syncthreads();
    for (int kb = tid; kb < nblk; kb += kThreads) {
      const double* xb = x + bt[6 * kb];
      const int gs = bt[6 * kb + 1], idx = bt[6 * kb + 2];
      const double* y = v.C(C_PRIOR_X0) + bt[6 * kb + 4];
      if (gs == 7 && bt[6 * kb + 5] == 6) {
        dyn[idx] = xb[0] - y[0];
        dyn[idx + 1] = xb[1] - y[1];
        dyn[idx + 2] = xb[2] - y[2];
        const Quat q = qmul(qinv(pose_q(y)), pose_q(xb));
        const double sgn = (q.w >= 0) ? 1.0 : -1.0;
        dyn[idx + 3] = 2.0 * sgn * q.x;
        dyn[idx + 4] = 2.0 * sgn * q.y;
        dyn[idx + 5] = 2.0 * sgn * q.z;
      } else {
        for (int k = 0; k < gs; ++k) dyn[idx + k] = xb[k] - y[k];
      }
    }
    __syncthreads();
    for (int r = tid; r < n; r += kThreads) {
      double s = 0.0;
      for (int c = 0; c < n; ++c) s += J0[(size_t)r * n + c] * dyn[c];
      s = r0[r] + s;
      if (!finite_d(s)) bad = 1;
      if (res_off < 0) { fixed += 0.5 * s * s; continue; }
      cost += 0.5 * s * s;
      if (full) R[res_off + r] = s;
    }
    if (full && res_off >= 0) {
      for (int kb = 0; kb < nblk; ++kb) {
        const int jo = bt[6 * kb + 3];
        if (jo < 0) continue;
        const int idx = bt[6 * kb + 2], ls = bt[6 * kb + 5];
        for (int o = tid; o < n * ls; o += kThreads) {
          const int r = o / ls, c = o - r * ls;
          J[jo + o] = J0[(size_t)r * n + idx + c];
        }
      }
    }
  }

  {  // host-evaluated factors: residuals and global-size Jacobians were computed by the application's own cost functions at
     // this evaluation point and uploaded (swgn_graph.host_eval); the tangent Jacobian is the first `lsize` columns
    const int32_t* ht = v.I(I_HOST);
    const int32_t* hb = v.I(I_HOST_BLK);
    const int32_t* row_nres = v.I(I_ROW_NRES);
    const int32_t* rs_row = v.I(I_RS_ROW);
    const double* HB = v.W(W_HOSTBUF);
    for (int i = 0; i < d.n_host; ++i) {
      const int res_off = ht[4 * i], blk0 = ht[4 * i + 1], nblk = ht[4 * i + 2];
      if (res_off < 0 && !with_fixed) continue;
      const double* rec = HB + ht[4 * i + 3];
      // number of residuals: active factors read it from their row block; inactive ones from the record layout
      int nres;
      if (res_off >= 0) {
        nres = row_nres[rs_row[res_off]];
      } else {
        int gsum = 0;
        for (int kb = 0; kb < nblk; ++kb) gsum += hb[4 * (blk0 + kb) + 2];
        const int next = (i + 1 < d.n_host) ? ht[4 * (i + 1) + 3] : d.n_hostbuf;
        nres = (next - ht[4 * i + 3]) / (1 + gsum);
      }
      for (int r = tid; r < nres; r += kThreads) {
        const double rv = rec[r];
        if (!finite_d(rv)) bad = 1;
        if (res_off < 0) { fixed += 0.5 * rv * rv; continue; }
        cost += 0.5 * rv * rv;
        if (full) R[res_off + r] = rv;
      }
      if (full && res_off >= 0) {
        const double* Jg = rec + nres;
        for (int kb = 0; kb < nblk; ++kb) {
          const int jo = hb[4 * (blk0 + kb) + 1], gs = hb[4 * (blk0 + kb) + 2], ls = hb[4 * (blk0 + kb) + 3];
          if (jo >= 0)
            for (int o = tid; o < nres * ls; o += kThreads) {
              const int r = o / ls, c = o - r * ls;
              const double jv = Jg[r * gs + c];
              if (!finite_d(jv)) bad = 1;
              J[jo + o] = jv;
            }
          Jg += nres * gs;
        }
      }
    }
  }

  {  // IMUGNSSFactor chains: evaluated by k_chain (launched just before this kernel in the same mode)
    const int32_t* ct = v.I(I_CHAIN);
    for (int i = tid; i < d.n_chain; i += kThreads) {
      const int res_off = ct[8 * i + 2];
      if (res_off < 0 && !with_fixed) continue;
      const ChainLayout L(ct[8 * i], ct[8 * i + 1]);
      const double* fl = v.W(W_CHAIN) + ct[8 * i + 5] + L.w_flags;
      if (fl[3] != 0.0 || !finite_d(fl[2])) bad = 1;
      if (res_off < 0) fixed += fl[2];
      else cost += fl[2];
    }
  }

  const double total = block_sum(cost, red);
  const double fixed_total = with_fixed ? block_sum(fixed, red) : 0.0;
  const int any_bad = block_any(bad);  // also orders the J / R stores before the CSC pass

  double gmax = 0.0, xn2 = 0.0;
  if (full) {
    // gradient g = J^T r and dogleg diagonal, one thread per tangent scalar (CSC traversal);
    // BlockSparseMatrix::LeftMultiply / SquaredColumnNorm (block_sparse_matrix.cc:114,136)
    const int32_t* tcol = v.I(I_TCOL);
    const int32_t* col_pos = v.I(I_COL_POS);
    const int32_t* col_size = v.I(I_COL_SIZE);
    const int32_t* csc_ptr = v.I(I_CSC_PTR);
    const int32_t* csc_row = v.I(I_CSC_ROW);
    const int32_t* csc_val = v.I(I_CSC_VAL);
    const int32_t* row_res = v.I(I_ROW_RES);
    const int32_t* row_nres = v.I(I_ROW_NRES);
    double* G = v.W(W_G);
    double* DG = v.W(W_DIAG);
    double* SC = v.W(W_SCALE);
    // dogleg / LM diagonal sqrt(clamp(|column|^2)) of the (Jacobi-scaled) Jacobian; the scaling 1 / (1 + |column|) is taken
    // from the Jacobian of iteration zero and kept (trust_region_minimizer.cc:261-276); J itself stays unscaled in memory
    const bool jscale = b.params.jacobi_scaling != 0, new_scale = mode == EVAL_INIT;
    auto diag_of = [&](double ns, int idx) {
      if (jscale) {
        double sc;
        if (new_scale) {
          sc = 1.0 / (1.0 + sqrt(ns));
          SC[idx] = sc;
        } else {
          sc = SC[idx];
        }
        ns = ns * sc * sc;
      }
      return sqrt(fmin(fmax(ns, b.params.min_lm_diagonal), b.params.max_lm_diagonal));
    };
    constexpr int kHeavy = 24;  // columns with more row blocks than this are reduced by a whole warp
    const int32_t* csc_nres = v.I(I_CSC_NRES);
    const int32_t* csc_res = v.I(I_CSC_RES);
    // light columns (landmarks, clocks, ambiguities): one thread per column BLOCK, all its tangent
    // components at once, so the index loads are shared by the cs components
    for (int col = tid; col < d.n_cols; col += kThreads) {
      const int e0 = csc_ptr[col], e1 = csc_ptr[col + 1];
      if (e1 - e0 > kHeavy) continue;
      const int cs = col_size[col];
      for (int k0 = 0; k0 < cs; k0 += 9) {
        double g[9], nrm[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = nrm[k] = 0.0;
        // the index triple of the next entry is fetched while the current entry's values are in flight
        int nres_n = 0, val_n = 0, res_n = 0;
        if (e0 < e1) {
          nres_n = csc_nres[e0];
          val_n = csc_val[e0];
          res_n = csc_res[e0];
        }
        for (int e = e0; e < e1; ++e) {
          const int nres = nres_n;
          const double* jv = J + val_n + k0;
          const double* rv = R + res_n;
          if (e + 1 < e1) {
            nres_n = csc_nres[e + 1];
            val_n = csc_val[e + 1];
            res_n = csc_res[e + 1];
          }
          for (int rr = 0; rr < nres; ++rr) {
            const double r = rv[rr];
#pragma unroll
            for (int k = 0; k < 9; ++k)
              if (k0 + k < cs) {
                const double a = jv[rr * cs + k];
                g[k] += a * r;
                nrm[k] += a * a;
              }
          }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k)
          if (k0 + k < cs) {
            G[col_pos[col] + k0 + k] = g[k];
            DG[col_pos[col] + k0 + k] = diag_of(nrm[k], col_pos[col] + k0 + k);
          }
      }
    }
    // heavy columns (poses seen by ~100 observations): lanes stride the row blocks, every lane
    // accumulates all cs columns of its rows, then a deterministic shuffle reduction
    for (int col = wid; col < d.n_cols; col += kWarps) {
      const int e0 = csc_ptr[col], e1 = csc_ptr[col + 1];
      if (e1 - e0 <= kHeavy) continue;
      const int cs = col_size[col];
      for (int k0 = 0; k0 < cs; k0 += 9) {  // up to 9 tangent columns per pass (pose 6, speed-bias 9)
        double g[9], nrm[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) g[k] = nrm[k] = 0.0;
        for (int e = e0 + lane; e < e1; e += 32) {
          const int nres = csc_nres[e];
          const double* jv = J + csc_val[e] + k0;
          const double* rv = R + csc_res[e];
          for (int rr = 0; rr < nres; ++rr) {
            const double r = rv[rr];
#pragma unroll
            for (int k = 0; k < 9; ++k)
              if (k0 + k < cs) {
                const double a = jv[rr * cs + k];
                g[k] += a * r;
                nrm[k] += a * a;
              }
          }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const double gs = warp_sum(g[k]), ns = warp_sum(nrm[k]);
          if (lane == 0 && k0 + k < cs) {
            G[col_pos[col] + k0 + k] = gs;
            DG[col_pos[col] + k0 + k] = diag_of(ns, col_pos[col] + k0 + k);
          }
        }
      }
    }
    __syncthreads();
    // max |x - Plus(x, -g)| (trust_region_minimizer.cc:266-287) and |x| over the reduced program
    const int32_t* col_state = v.I(I_COL_STATE);
    const int32_t* col_gsize = v.I(I_COL_GSIZE);
    for (int c = tid; c < d.n_cols; c += kThreads) {
      const double* xb = x + col_state[c];
      const int gs = col_gsize[c], ls = col_size[c];
      const double* g = G + col_pos[c];
      if (gs == 7 && ls == 6) {
        double ng[6], out[7];
#pragma unroll
        for (int k = 0; k < 6; ++k) ng[k] = -g[k];
        block_plus(xb, ng, out, 7, 6);
#pragma unroll
        for (int k = 0; k < 7; ++k) { gmax = fmax(gmax, fabs(xb[k] - out[k])); xn2 += xb[k] * xb[k]; }
      } else {
        for (int k = 0; k < gs; ++k) { gmax = fmax(gmax, fabs(xb[k] - (xb[k] + (-g[k])))); xn2 += xb[k] * xb[k]; }
      }
    }
    gmax = block_max(gmax, red);
    xn2 = block_sum(xn2, red);
  }

  if (mode == EVAL_INIT) {  // keep the uploaded state: best point so far / restore point
    const double* xs = v.W(W_X);
    double* xb = v.W(W_XBEST);
    double* x0 = v.W(W_X0);
    double* xc = v.W(W_XCAND);  // constant blocks are never rewritten by k_step
    for (int k = tid; k < d.n_state; k += kThreads) { xb[k] = xs[k]; x0[k] = xs[k]; xc[k] = xs[k]; }
  }
  if (tid == 0) {
    const SolverParams& P = b.params;
    if (mode == EVAL_INIT) {  // IterationZero, trust_region_minimizer.cc:195-230
      TRState s;
      memset(&s, 0, sizeof(s));
      s.x_cost = total;
      s.fixed_cost = fixed_total;
      s.initial_cost = total + fixed_total;
      s.iter_cost = s.initial_cost;
      s.final_cost = s.initial_cost;
      s.minimum_cost = DBL_MAX;
      s.candidate_cost = 0.0;
      s.radius = P.initial_radius;
      s.decrease_factor = 2.0;
      s.mu = P.min_mu;
      s.x_norm = sqrt(xn2);
      s.gradient_max_norm = gmax;
      s.se_min = s.se_cur = s.se_ref = s.se_cand = total;
      s.last_successful = 1;
      s.active = 1;
      s.termination = SWGN_NO_CONVERGENCE;
      if (any_bad) {  // "Residual and Jacobian evaluation failed."
        s.active = 0;
        s.termination = SWGN_FAILURE;
      }
      *st = s;
    } else if (mode == EVAL_CANDIDATE) {
      st->candidate_cost = any_bad ? DBL_MAX : total;
    } else if (mode == EVAL_ACCEPTED) {
      st->x_cost = total;
      st->gradient_max_norm = gmax;
      st->x_norm = sqrt(xn2);
      st->last_successful = 1;
      st->iter_cost = total + st->fixed_cost;
      if (any_bad) {
        st->active = 0;
        st->termination = SWGN_FAILURE;
      }
    } else {  // EVAL_FORCE (staged test entry point): only the cost is reported
      st->x_cost = total;
      st->fixed_cost = fixed_total;
      st->gradient_max_norm = gmax;
    }
  }
}

void launch_eval(const DeviceBatch& b, int mode, int only_window, cudaStream_t s) {
  const int grid = only_window >= 0 ? 1 : b.n_windows;
  // two spare doubles: the compiler reads the prior's dx vector with 16-byte shared loads
  const size_t dyn = sizeof(double) * (((size_t)(b.max_prior_n > 0 ? b.max_prior_n : 1) + 3) & ~(size_t)1);
  launch_chain(b, mode, only_window, s);
  k_eval<<<grid, kThreads, dyn, s>>>(b, mode, only_window);
}

}  // namespace swgn
