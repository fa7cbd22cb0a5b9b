// K1c: the stateful IMUGNSSFactor (RVI/factor/gnss_imu_factor.cpp:678-835) evaluated on the device,
// one CTA per chain factor, launched right before k_eval in the same mode.  The m GNSS frames hidden
// between keyframes i and j are eliminated frame by frame (Schur complement of one 15-dim frame at a
// time into a block system over [next frame 15 | phase biases k | keyframe i 15]); the resulting
// (30+k)^2 information matrix over (kf_i, kf_j, N) is factored H = V S V' by a parallel-ordered
// Jacobi eigen-solver in shared memory, J = sqrt(S) V', r = S^-1/2 V' rhs.  Cost-only evaluations
// (trust-region candidates) use the linearised residual r - J * INC; at the next Jacobian
// evaluation the hidden states follow the accepted step by back-substitution through the saved
// elimination blocks.  All of the factor's mutable state lives in the window's W_CHAIN area.
//   IMUGNSSBase::Evaluate :678-799, JacobianResidualUpdateHessianRhs :358-379, MargPose1 :403-435,
//   MoveHessianData :437-456, UpdateSchurComponent :458-494, UpdateJacobResidual :495-530,
//   UpdateRhsN/UpdateRhsPose :532-564, UpdateDeltaValues :566-611, UpdateHiddenState :613-646,
//   GetInc :676-691, IMUFactor::Evaluate2 imu_factor.cpp:103-193.
// The eigenvectors of a symmetric matrix are defined up to sign and order, so the rows of J are
// not comparable with another eigen-solver's; J'J, J'r and |r| are.
#include <float.h>

#include "dev_common.cuh"
#include "dev_eig.cuh"
#include "dev_imu.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

constexpr int NT = 256;
enum { B_P1 = 0, B_P2 = 1, B_N = 2, B_P0 = 3 };
constexpr double kEigEps = 1e-8;  // gnss_imu_factor.cpp:9

// shared-memory layout (doubles) for one chain with k phase biases.  Block sizes are (15, 15, k, 15)
// for (P1, P2, N, P0); the offsets are closed forms so that nothing here needs an indexed local array.
struct Sm {
  int k, rhs0, delta0;
  int inv, anm, raw, j12, r15, dx, nval, A, V, rd, cs, red, total;
  __host__ __device__ explicit Sm(int k_) : k(k_) {
    const int n = 30 + k;
    int o = 1350 + 45 * k + k * k;  // the ten upper blocks of H
    rhs0 = o; o += 45 + k;
    delta0 = o; o += 45 + k;
    inv = o; o += 225;
    anm = o; o += (k > 15 ? k : 15) * 15;
    raw = o; o += 450;
    j12 = o; o += 450;
    r15 = o; o += 16;
    dx = o; o += 16;
    nval = o; o += k + 1;
    A = o; o += n * (n | 1);
    V = o; o += n * (n | 1);
    rd = o; o += n;
    cs = o; o += 4 * ((n + 1) / 2) + 4;
    red = o; o += 34;
    total = o;
  }
  __host__ __device__ int size(int b) const { return b == B_N ? k : 15; }
  __host__ __device__ int H(int i, int j) const {  // offset of block (i, j), i <= j
    switch (i * 4 + j) {
      case 0: return 0;
      case 1: return 225;
      case 2: return 450;
      case 3: return 450 + 15 * k;
      case 5: return 675 + 15 * k;
      case 6: return 900 + 15 * k;
      case 7: return 900 + 30 * k;
      case 10: return 1125 + 30 * k;
      case 11: return 1125 + 30 * k + k * k;
      default: return 1125 + 45 * k + k * k;  // (3, 3)
    }
  }
  __host__ __device__ int vec(int base, int b) const { return base + (b == 0 ? 0 : (b == 1 ? 15 : (b == 2 ? 30 : 30 + k))); }
  __host__ __device__ int rhs(int b) const { return vec(rhs0, b); }
  __host__ __device__ int delta(int b) const { return vec(delta0, b); }
};

// x [-] x0 for (pose, speed-bias), sign-fixed like GetInc :676-691; sign = -1 gives x0 [-] x (:566-611)
__device__ void inc15(const double* x, const double* x0, const double* s, const double* s0, double sign, double* dx) {
  for (int i = 0; i < 3; ++i) dx[i] = sign * (x[i] - x0[i]);
  const Quat q = qmul(qinv(pose_q(x0)), pose_q(x));
  double f = 2.0 * sign;
  if (!(q.w >= 0)) f = -f;
  dx[3] = f * q.x; dx[4] = f * q.y; dx[5] = f * q.z;
  for (int i = 0; i < 9; ++i) dx[6 + i] = sign * (s[i] - s0[i]);
}

}  // namespace

__global__ void __launch_bounds__(NT, 2) k_chain(DeviceBatch b, int mode, int only_window) {
  __shared__ WinDesc sd;
  extern __shared__ double sm[];
  const int w = only_window >= 0 ? only_window : blockIdx.x;
  const int c = blockIdx.y;
  const TRState* st = b.state + w;
  bool run;
  if (mode == EVAL_INIT || mode == EVAL_FORCE) run = true;
  else if (mode == EVAL_CANDIDATE) run = st->active && st->step_valid && !st->need_solve;
  else run = st->active && st->accepted;
  if (!run) return;
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  if (c >= d.n_chain) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int32_t* rec = v.I(I_CHAIN) + 8 * c;
  const int m = rec[0], k = rec[1], res_off = rec[2], n = 30 + k;
  const bool with_fixed = mode == EVAL_INIT || mode == EVAL_FORCE;
  if (res_off < 0 && !with_fixed) return;
  const bool update_flag = mode != EVAL_CANDIDATE && res_off >= 0;  // jacobians requested
  const ChainLayout L(m, k);
  const Sm S(k);
  const double* C = v.C(C_CHAIN) + rec[4];
  double* Wk = v.W(W_CHAIN) + rec[5];
  const int32_t* blk = v.I(I_CHAIN_BLK) + 2 * rec[3];
  const double* x = v.W(mode == EVAL_CANDIDATE ? W_XCAND : W_X);
  const double *Pi = x + blk[0], *Bi = x + blk[2], *Pj = x + blk[4], *Bj = x + blk[6];
  double Pbg[3], G[3];
  {
    const double* g = v.C(C_GLOBALS);
    for (int q = 0; q < 3; ++q) { Pbg[q] = g[q]; G[q] = g[3 + q]; }
  }
  double* frames = Wk + L.w_frames;
  double* flags = Wk + L.w_flags;
  double* old = Wk + L.w_old;  // Pi 7 | Bi 9 | Pj 7 | Bj 9 | N k
  double* INC = Wk + L.w_inc;

  // new inputs (swgn_batch_create / swgn_batch_update_inputs): reload the hidden states, forget history
  if (flags[1] != (double)b.chain_epoch) {
    __syncthreads();
    for (int q = tid; q < m * 16; q += NT) frames[q] = C[L.c_frame + (q / 16) * CHAIN_FRAME_STRIDE + (q % 16)];
    if (tid == 0) { flags[0] = 0.0; flags[1] = (double)b.chain_epoch; }
    __syncthreads();
  }
  const bool history = flags[0] != 0.0;
  for (int q = tid; q < k; q += NT) sm[S.nval + q] = x[blk[2 * (4 + q)]];
  __syncthreads();
  auto save_last = [&]() {  // SaveLastStates :669-675
    for (int q = tid; q < 32 + k; q += NT) {
      double val;
      if (q < 7) val = Pi[q];
      else if (q < 16) val = Bi[q - 7];
      else if (q < 23) val = Pj[q - 16];
      else if (q < 32) val = Bj[q - 23];
      else val = sm[S.nval + q - 32];
      old[q] = val;
    }
    __syncthreads();
  };
  if (!history) save_last();
  // UpdateDeltaValues :566-611
  if (tid == 0) inc15(Pj, old + 16, Bj, old + 23, -1.0, sm + S.delta(B_P2));
  if (tid == 32) inc15(Pi, old, Bi, old + 7, -1.0, sm + S.delta(B_P0));
  for (int q = tid; q < k; q += NT) sm[S.delta(B_N) + q] = old[32 + q] - sm[S.nval + q];
  __syncthreads();
  for (int q = tid; q < n; q += NT)
    INC[q] = q < 15 ? sm[S.delta(B_P0) + q] : (q < 30 ? sm[S.delta(B_P2) + q - 15] : sm[S.delta(B_N) + q - 30]);
  __syncthreads();

  if (history && update_flag) {  // UpdateHiddenState :613-646
    for (int i = m - 1; i >= 0; --i) {
      double* sv = Wk + L.w_save + i * L.save_stride;  // Amm_inv 225 | H12 225 | H1N 15k | H10 225 | rhs 15
      double* rs = sv + 675 + 15 * k;
      if (tid < 15) {
        double acc = rs[tid];
        const double* h12 = sv + 225 + tid * 15;
        for (int q = 0; q < 15; ++q) acc -= h12[q] * sm[S.delta(B_P2) + q];
        const double* h1n = sv + 450 + tid * k;
        for (int q = 0; q < k; ++q) acc -= h1n[q] * sm[S.delta(B_N) + q];
        const double* h10 = sv + 450 + 15 * k + tid * 15;
        for (int q = 0; q < 15; ++q) acc -= h10[q] * sm[S.delta(B_P0) + q];
        rs[tid] = acc;
      }
      __syncthreads();
      if (tid < 15) {
        double acc = 0.0;
        for (int q = 0; q < 15; ++q) acc += sv[tid * 15 + q] * rs[q];
        sm[S.delta(B_P2) + tid] = acc;
      }
      __syncthreads();
      if (tid == 0) {
        double* P = frames + 16 * i;
        const double* dd = sm + S.delta(B_P2);
        for (int q = 0; q < 3; ++q) P[q] -= dd[q];
        const Quat dq = {1.0, -dd[3] / 2.0, -dd[4] / 2.0, -dd[5] / 2.0};
        const Quat r = qnormalized(qmul(pose_q(P), dq));
        P[3] = r.x; P[4] = r.y; P[5] = r.z; P[6] = r.w;
        for (int q = 0; q < 9; ++q) P[7 + q] -= dd[6 + q];
      }
      __syncthreads();
    }
  }

  if (!history || update_flag) {
    if (tid == 0) flags[0] = 1.0;
    save_last();
    for (int q = tid; q < S.delta0; q += NT) sm[q] = 0.0;  // ResetMem: all H blocks and rhs
    __syncthreads();
    {  // CopyHessian2Hessian / CopyRhs2Rhs / UpdateRhsN
      const double* NN = C + L.c_NN;
      for (int q = tid; q < k * k; q += NT) sm[S.H(B_N, B_N) + q] = NN[q];
      for (int a = tid; a < k; a += NT) {
        double acc = 0.0;
        for (int q = 0; q < k; ++q) acc += NN[a * k + q] * sm[S.nval + q];
        sm[S.rhs(B_N) + a] = C[L.c_Nrhs + a] + acc;
      }
    }
    __syncthreads();
    // IMU link idx between (pa, sa) and (pb, sb): r15, J12 (15 x 30 = [J1 | J2]) -> H / rhs of blocks (b0, b1)
    auto imu_link = [&](int idx, const double* pa, const double* sa, const double* pb, const double* sb, int b0, int b1) {
      const double* irec = C + L.c_imu + IMU_DEV_STRIDE * idx;
      if (tid < 32) {
        const double rk = imu_residual_raw(irec, Pbg, G, pa, sa, pb, sb, sm + S.raw, true, lane);
        if (lane < 15) sm[S.r15 + lane] = rk;
      }
      __syncthreads();
      const double* sq = irec + IMU_DEV_SQRT;
      for (int o = tid; o < 450; o += NT) {
        const int a = o / 30, cc = o - a * 30;
        double acc = 0.0;
        for (int q = 0; q < 15; ++q) acc += sq[a * 15 + q] * sm[S.raw + q * 30 + cc];
        sm[S.j12 + o] = acc;
      }
      __syncthreads();
      // JacobianResidualUpdateHessianRhs :358-379
      for (int o = tid; o < 30; o += NT) {
        const int i = o / 15, cc = o - i * 15;
        double acc = 0.0;
        for (int a = 0; a < 15; ++a) acc += sm[S.j12 + a * 30 + i * 15 + cc] * sm[S.r15 + a];
        sm[S.rhs(i == 0 ? b0 : b1) + cc] += acc;
      }
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
          const int bi = i == 0 ? b0 : b1, bj = j == 0 ? b0 : b1;
          if (bj < bi) continue;
          double* h = sm + S.H(bi, bj);
          for (int o = tid; o < 225; o += NT) {
            const int cc = o / 15, dd = o - cc * 15;
            double acc = 0.0;
            for (int a = 0; a < 15; ++a) acc += sm[S.j12 + a * 30 + i * 15 + cc] * sm[S.j12 + a * 30 + j * 15 + dd];
            h[o] += acc;
          }
        }
      __syncthreads();
    };
    imu_link(0, Pi, Bi, frames, frames + 7, B_P0, B_P1);
    for (int i = 0; i < m; ++i) {
      const double* hp = frames + 16 * i;
      if (i != m - 1) imu_link(i + 1, hp, hp + 7, hp + 16, hp + 23, B_P1, B_P2);
      else imu_link(m, hp, hp + 7, Pj, Bj, B_P1, B_P2);
      // UpdateRhsPose(i) :540-564 and the frame's GNSS information
      const double* fr = C + L.c_frame + i * CHAIN_FRAME_STRIDE;
      const double* ph = fr + SWGN_CHAIN_HESSIAN;
      const double* pn = C + L.c_frameN + i * L.pn_stride;
      if (tid == 0) inc15(hp, fr + SWGN_CHAIN_POSE_LIN, hp + 7, fr + SWGN_CHAIN_SB_LIN, 1.0, sm + S.dx);
      __syncthreads();
      if (tid < 15) {
        double acc = 0.0;
        for (int q = 0; q < 15; ++q) acc += ph[tid * 15 + q] * sm[S.dx + q];
        double r1 = sm[S.rhs(B_P1) + tid] + acc;
        acc = 0.0;
        for (int q = 0; q < k; ++q) acc += pn[tid * k + q] * sm[S.nval + q];
        r1 += acc;
        sm[S.rhs(B_P1) + tid] = r1 + fr[SWGN_CHAIN_RHS + tid];
      }
      for (int q = tid; q < k; q += NT) {
        double acc = 0.0;
        for (int a = 0; a < 15; ++a) acc += pn[a * k + q] * sm[S.dx + a];
        sm[S.rhs(B_N) + q] += acc;
      }
      for (int q = tid; q < 225; q += NT) sm[S.H(B_P1, B_P1) + q] += ph[q];
      for (int q = tid; q < 15 * k; q += NT) sm[S.H(B_P1, B_N) + q] += pn[q];
      __syncthreads();
      // MargPose1 :403-435.  InvertPSDMatrix<15>: LLT of the upper triangle, solve for the identity
      // (invert_psd_matrix.h:62-67)
      {
        double* U = sm + S.H(B_P1, B_P1);
        double* inv = sm + S.inv;
        if (tid < 32) {  // row kk of U by the lanes j > kk, same arithmetic as the scalar LLT
          bool ok = true;
          for (int kk = 0; kk < 15; ++kk) {
            double xx = U[kk * 15 + kk];
            for (int p = 0; p < kk; ++p) xx -= U[p * 15 + kk] * U[p * 15 + kk];
            if (!(xx > 0.0)) { ok = false; break; }  // uniform across the warp
            xx = sqrt(xx);
            __syncwarp();
            if (lane == kk) U[kk * 15 + kk] = xx;
            if (lane > kk && lane < 15) {
              double s = U[kk * 15 + lane];
              for (int p = 0; p < kk; ++p) s -= U[p * 15 + kk] * U[p * 15 + lane];
              U[kk * 15 + lane] = s / xx;
            }
            __syncwarp();
          }
          if (lane == 0) sm[S.r15 + 15] = ok ? 1.0 : 0.0;
        }
        __syncthreads();
        const bool ok = sm[S.r15 + 15] != 0.0;
        if (tid < 15) {
          double col[15];
#pragma unroll
          for (int i2 = 0; i2 < 15; ++i2) col[i2] = (i2 == tid) ? 1.0 : 0.0;
#pragma unroll
          for (int i2 = 0; i2 < 15; ++i2) {
            double s = col[i2];
#pragma unroll
            for (int p = 0; p < 15; ++p)
              if (p < i2) s -= U[p * 15 + i2] * col[p];
            col[i2] = s / U[i2 * 15 + i2];
          }
#pragma unroll
          for (int i2 = 14; i2 >= 0; --i2) {
            double s = col[i2];
#pragma unroll
            for (int p = 0; p < 15; ++p)
              if (p > i2) s -= U[i2 * 15 + p] * col[p];
            col[i2] = s / U[i2 * 15 + i2];
          }
#pragma unroll
          for (int i2 = 0; i2 < 15; ++i2) inv[i2 * 15 + tid] = ok ? col[i2] : nan("");
        }
        __syncthreads();
        for (int q = tid; q < 225; q += NT) U[q] = inv[q];
        __syncthreads();
        for (int bi = B_P1 + 1; bi < 4; ++bi) {
          const int sn = S.size(bi);
          const double* h1i = sm + S.H(B_P1, bi);  // 15 x sn
          double* anm = sm + S.anm;                 // sn x 15 = H1i' * inv
          for (int o = tid; o < sn * 15; o += NT) {
            const int a = o / 15, cc = o - a * 15;
            double acc = 0.0;
            for (int t = 0; t < 15; ++t) acc += h1i[t * sn + a] * inv[t * 15 + cc];
            anm[o] = acc;
          }
          __syncthreads();
          for (int a = tid; a < sn; a += NT) {
            double acc = 0.0;
            for (int cc = 0; cc < 15; ++cc) acc += anm[a * 15 + cc] * sm[S.rhs(B_P1) + cc];
            sm[S.rhs(bi) + a] -= acc;
          }
          for (int bj = bi; bj < 4; ++bj) {
            const int svn = S.size(bj);
            const double* h1j = sm + S.H(B_P1, bj);
            double* hij = sm + S.H(bi, bj);
            for (int o = tid; o < sn * svn; o += NT) {
              const int a = o / svn, bb = o - a * svn;
              double acc = 0.0;
              for (int cc = 0; cc < 15; ++cc) acc += anm[a * 15 + cc] * h1j[cc * svn + bb];
              hij[o] -= acc;
            }
          }
          __syncthreads();
        }
      }
      // MoveHessianData(i) :437-456
      {
        double* sv = Wk + L.w_save + i * L.save_stride;
        for (int q = tid; q < 225; q += NT) {
          sv[q] = sm[S.H(B_P1, B_P1) + q];
          sv[225 + q] = sm[S.H(B_P1, B_P2) + q];
          sv[450 + 15 * k + q] = sm[S.H(B_P1, B_P0) + q];
        }
        for (int q = tid; q < 15 * k; q += NT) sv[450 + q] = sm[S.H(B_P1, B_N) + q];
        if (tid < 15) sv[675 + 15 * k + tid] = sm[S.rhs(B_P1) + tid];
        __syncthreads();
        for (int q = tid; q < 225; q += NT) {
          sm[S.H(B_P1, B_P1) + q] = sm[S.H(B_P2, B_P2) + q];
          sm[S.H(B_P2, B_P2) + q] = 0.0;
          sm[S.H(B_P1, B_P0) + q] = sm[S.H(B_P2, B_P0) + q];
          sm[S.H(B_P2, B_P0) + q] = 0.0;
          sm[S.H(B_P1, B_P2) + q] = 0.0;
        }
        for (int q = tid; q < 15 * k; q += NT) {
          sm[S.H(B_P1, B_N) + q] = sm[S.H(B_P2, B_N) + q];
          sm[S.H(B_P2, B_N) + q] = 0.0;
        }
        if (tid < 15) {
          sm[S.rhs(B_P1) + tid] = sm[S.rhs(B_P2) + tid];
          sm[S.rhs(B_P2) + tid] = 0.0;
        }
        __syncthreads();
      }
    }
    // UpdateSchurComponent :458-494: dense H over (kf_i | kf_j | N) = blocks (P0 | P1 | N), upper
    // triangle mirrored (selfadjointView<Upper>)
    double* A = sm + S.A;
    double* V = sm + S.V;
    const int ld = n | 1;
    auto assemble = [&]() {
      for (int o = tid; o < n * n; o += NT) {
        int ra = o / n, cb = o - ra * n;
        if (cb < ra) { const int t = ra; ra = cb; cb = t; }
        const int i = ra < 15 ? 0 : (ra < 30 ? 1 : 2), j = cb < 15 ? 0 : (cb < 30 ? 1 : 2);
        const int a = ra - 15 * i, bb = cb - 15 * j;
        const int i2 = i == 0 ? B_P0 : (i == 1 ? B_P1 : B_N), j2 = j == 0 ? B_P0 : (j == 1 ? B_P1 : B_N);
        A[(o / n) * ld + (o % n)] = (j2 >= i2) ? sm[S.H(i2, j2) + a * S.size(j2) + bb] : sm[S.H(j2, i2) + bb * S.size(i2) + a];
        V[(o / n) * ld + (o % n)] = (o / n == o % n) ? 1.0 : 0.0;
      }
      for (int q = tid; q < n; q += NT) sm[S.rd + q] = q < 15 ? sm[S.rhs(B_P0) + q] : (q < 30 ? sm[S.rhs(B_P1) + q - 15] : sm[S.rhs(B_N) + q - 30]);
      __syncthreads();
    };
    assemble();
    // Fast path: when H is safely positive definite no eigenvalue is dropped, and ANY square root gives the
    // solver the same J'J = H, J'r = rhs and |r|^2 = rhs' H^-1 rhs as the reference's sqrt(S) V'.  Try the
    // Cholesky factor H = L L' (J = L', r = L^-1 rhs) and accept it when 1 / trace(H^-1) = 1 / |L^-1|_F^2,
    // a lower bound of the smallest eigenvalue, clears the reference's 1e-8 threshold with margin;
    // otherwise fall through to the eigen-decomposition, which is what drops eigenvalues <= 1e-8.
    bool use_chol = false;
    {
      bool ok = true;
      // right-looking, one barrier per column: the trailing update works on the UNSCALED columns
      // (A[i][c] -= A[i][j] A[c][j] / piv_j, lower triangle; one warp per row, lanes over the columns) and the
      // columns are scaled to L in one pass at the end
      const int wid = tid >> 5, nwarp = NT / 32;
      for (int j = 0; j < n; ++j) {
        const double piv = A[j * ld + j];
        if (!(piv > 0.0) || !finite_d(piv)) { ok = false; break; }  // uniform: every thread reads the same value
        const double pinv = 1.0 / piv;
        for (int i = j + 1 + wid; i < n; i += nwarp) {
          const double lij = A[i * ld + j] * pinv;
          for (int c = j + 1 + lane; c <= i; c += 32) A[i * ld + c] -= lij * A[c * ld + j];
        }
        __syncthreads();
      }
      if (ok) {
        for (int o = tid; o < n * n; o += NT) {
          const int i = o / n, j = o - i * n;
          if (j < i) A[i * ld + j] *= 1.0 / sqrt(A[j * ld + j]);
        }
        __syncthreads();
        for (int j = tid; j < n; j += NT) A[j * ld + j] = sqrt(A[j * ld + j]);
        __syncthreads();
      }
      if (ok) {
        // X = L^-1 (lower), one thread per column, into V
        __syncthreads();
        double fro = 0.0;
        for (int c = tid; c < n; c += NT) {
          for (int i = 0; i < n; ++i) {
            double acc = (i == c) ? 1.0 : 0.0;
            if (i < c) { V[i * ld + c] = 0.0; continue; }
            for (int q = c; q < i; ++q) acc -= A[i * ld + q] * V[q * ld + c];
            acc /= A[i * ld + i];
            V[i * ld + c] = acc;
            fro += acc * acc;
          }
        }
        fro = block_sum(fro, sm + S.red);
        use_chol = finite_d(fro) && fro > 0.0 && 1.0 / fro > 16.0 * kEigEps;
      }
      __syncthreads();
      if (use_chol) {
        double* Jd = Wk + L.w_J;
        double* rdst = Wk + L.w_r;
        for (int o = tid; o < n * n; o += NT) {
          const int i = o / n, cc = o - i * n;
          Jd[o] = cc >= i ? A[cc * ld + i] : 0.0;  // J = L'
        }
        for (int i = tid; i < n; i += NT) {
          double acc = 0.0;
          for (int cc = 0; cc <= i; ++cc) acc += V[i * ld + cc] * sm[S.rd + cc];
          rdst[i] = acc;
        }
        __syncthreads();
      } else {
        assemble();  // A and V were used as scratch: rebuild H and V = I from the blocks
      }
    }
    if (!use_chol) {
    jacobi_eig<NT>(A, V, n, ld, sm + S.cs, sm + S.red);
    // schur_jacobian = sqrt(S) V', schur_residual = S^-1/2 V' rhs, eigenvalues <= eps dropped
    {
      double* Jd = Wk + L.w_J;
      double* rdst = Wk + L.w_r;
      for (int o = tid; o < n * n; o += NT) {
        const int i = o / n, cc = o - i * n;
        const double lam = A[i * ld + i];
        Jd[o] = (lam > kEigEps ? sqrt(lam) : 0.0) * V[cc * ld + i];
      }
      for (int i = tid; i < n; i += NT) {
        const double lam = A[i * ld + i];
        double acc = 0.0;
        for (int cc = 0; cc < n; ++cc) acc += V[cc * ld + i] * sm[S.rd + cc];
        rdst[i] = (lam > kEigEps ? sqrt(1.0 / lam) : 0.0) * acc;
      }
      __syncthreads();
    }
    }  // !use_chol
  }

  // UpdateJacobResidual :495-530
  {
    const double* Jd = Wk + L.w_J;
    const double* rsrc = Wk + L.w_r;
    double* R = v.W(W_RES);
    double* Jw = v.W(W_JAC);
    double cost = 0.0;
    int bad = 0;
    for (int a = tid; a < n; a += NT) {
      double s = rsrc[a];
      if (!update_flag) {
        double t = 0.0;
        for (int cc = 0; cc < n; ++cc) t += Jd[a * n + cc] * INC[cc];
        s -= t;
      }
      if (!finite_d(s)) bad = 1;
      cost += 0.5 * s * s;
      if (update_flag) R[res_off + a] = s;
    }
    if (update_flag) {
      for (int p = 0; p < 4 + k; ++p) {
        const int jo = blk[2 * p + 1];
        if (jo < 0) continue;
        const int lsz = p >= 4 ? 1 : ((p & 1) ? 9 : 6);
        const int c0 = p >= 4 ? 30 + (p - 4) : (p == 0 ? 0 : (p == 1 ? 6 : (p == 2 ? 15 : 21)));
        for (int o = tid; o < n * lsz; o += NT) {
          const int a = o / lsz, cc = o - a * lsz;
          const double val = Jd[a * n + c0 + cc];
          if (!finite_d(val)) bad = 1;
          Jw[jo + o] = val;
        }
      }
    }
    cost = block_sum(cost, sm + S.red);
    bad = block_any(bad);
    if (tid == 0) {
      flags[2] = cost;
      flags[3] = bad ? 1.0 : 0.0;
    }
  }
}

static size_t chain_smem(int max_k) { return sizeof(double) * (size_t)Sm(max_k).total; }

void launch_chain(const DeviceBatch& b, int mode, int only_window, cudaStream_t s) {
  if (b.max_chain <= 0) return;
  dim3 grid(only_window >= 0 ? 1 : b.n_windows, b.max_chain);
  k_chain<<<grid, NT, chain_smem(b.max_chain_k), s>>>(b, mode, only_window);
}

cudaError_t configure_chain(const DeviceBatch& b) {
  if (b.max_chain <= 0) return cudaSuccess;
  const size_t dyn = chain_smem(b.max_chain_k);
  if (dyn > 227 * 1024) return cudaErrorInvalidValue;
  if (dyn > 48 * 1024) return cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  return cudaSuccess;
}

}  // namespace swgn
