// K6: the per-window trust-region state machine -- TrustRegionMinimizer::Minimize
// (CERES/internal/ceres/trust_region_minimizer.cc:67-134) with the TRADITIONAL_DOGLEG strategy
// (dogleg_strategy.cc:79-253, 515-638, kMinMu = 1e-12 as modified by the reference, :51) or the
// LEVENBERG_MARQUARDT strategy (levenberg_marquardt_strategy.cc:66-165; with jacobi_scaling the solve runs on the
// unscaled Jacobian with the damping diagonal divided by the scaling, which is the same linear system in the
// unscaled variable) and TrustRegionStepEvaluator (trust_region_step_evaluator.cc:52-112).  The control flow is data
// dependent per window (accept / reject, mu retries, early convergence), so it lives on the device
// as masks in TRState; the host only issues "ticks" of
//   k_begin -> k_schur -> k_chol -> k_backsub -> k_step -> k_eval(candidate) -> k_end -> k_eval(accepted)
// until no window is active.  One tick is one trust-region iteration, or one mu-retry of its
// linear solve.
#include <float.h>

#include "dev_common.cuh"
#include "../../include/swgn.h"

namespace swgn {
namespace {

constexpr int kThreads = 256;

// y = J v for the block-sparse Jacobian, one thread per residual scalar
// (BlockSparseMatrix::RightMultiply, block_sparse_matrix.cc:92)
__device__ __forceinline__ double row_dot(const Win& v, int rs, const double* vec) {
  const int row = v.I(I_RS_ROW)[rs];
  const int rr = rs - v.I(I_ROW_RES)[row];
  const int32_t* row_cell = v.I(I_ROW_CELL);
  const int32_t* cell_col = v.I(I_CELL_COL);
  const int32_t* cell_val = v.I(I_CELL_VAL);
  const int32_t* col_size = v.I(I_COL_SIZE);
  const int32_t* col_pos = v.I(I_COL_POS);
  const double* J = v.W(W_JAC);
  double s = 0.0;
  for (int c = row_cell[row]; c < row_cell[row + 1]; ++c) {
    const int col = cell_col[c], cs = col_size[col];
    const double* jv = J + cell_val[c] + rr * cs;
    const double* x = vec + col_pos[col];
    for (int k = 0; k < cs; ++k) s += jv[k] * x[k];
  }
  return s;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// k_begin: close the previous iteration (FinalizeIterationAndCheckIfMinimizerCanContinue,
// trust_region_minimizer.cc:303-365), open the next one and prepare the dogleg subproblem
// (DoglegStrategy::ComputeStep :108-131: diagonal, scaled gradient, Cauchy point alpha).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_begin(DeviceBatch b, int tick) {
  __shared__ WinDesc sd;
  __shared__ double red[33];
  __shared__ int s_flag[4];
  const int w = blockIdx.x;
  TRState* st = b.state + w;
  if (!st->active) return;
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  const SolverParams& P = b.params;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int copy_best = 0, prep = 0;
    if (!st->need_solve) {
      if (st->last_successful) {
        st->num_successful += 1;
        if (st->x_cost < st->minimum_cost) {
          st->minimum_cost = st->x_cost;
          copy_best = 1;
        }
      } else {
        st->num_unsuccessful += 1;
      }
      st->final_cost = fmin(st->final_cost, st->iter_cost);
      int term = -1;
      if (st->iteration >= P.max_num_iterations) term = SWGN_NO_CONVERGENCE;
      else if (st->last_successful && st->gradient_max_norm <= P.gradient_tolerance) term = SWGN_CONVERGENCE;
      else if (st->radius <= P.min_radius) term = SWGN_CONVERGENCE;
      if (term >= 0) {
        st->active = 0;
        st->termination = term;
      } else {
        st->iteration += 1;
        st->step_valid = 0;
        st->accepted = 0;
        if (P.strategy == SWGN_LEVENBERG_MARQUARDT) {  // every iteration solves with D = sqrt(diagonal / radius)
          st->need_solve = 1;
          st->solve_ok = 0;
        } else if (!st->reuse) {
          st->reuse = 1;
          st->need_solve = 1;
          st->solve_ok = 0;
          prep = 1;
        }
      }
    }
    if (P.strategy == SWGN_DOGLEG && st->active && st->need_solve && !(st->mu < 1.0)) {
      // the mu < max_mu loop of ComputeGaussNewtonStep ran out: linear solver failure (:542-597)
      st->need_solve = 0;
      st->solve_ok = 0;
    }
    s_flag[0] = copy_best;
    s_flag[1] = prep;
    s_flag[2] = st->active && st->need_solve;
    s_flag[3] = st->active;
  }
  __syncthreads();
  if (s_flag[0]) {
    const double* x = v.W(W_X);
    double* xb = v.W(W_XBEST);
    for (int k = tid; k < d.n_state; k += kThreads) xb[k] = x[k];
  }
  const double* DG = v.W(W_DIAG);
  if (s_flag[1]) {
    // ghat = (J^T r) / d  (:174-180).  The Cauchy point alpha = |ghat|^2 / |J (ghat / d)|^2 (:181-192) costs a pass over
    // the Jacobian and is only read when the Gauss-Newton step leaves the trust region: k_step computes it then
    // (same operands, same order of operations, so the value is the one ComputeCauchyPoint would have stored here)
    const double* G = v.W(W_G);
    double* GH = v.W(W_GHAT);
    double gs = 0.0;
    for (int k = tid; k < d.n_t; k += kThreads) {
      const double gh = G[k] / DG[k];
      GH[k] = gh;
      gs += gh * gh;
    }
    gs = block_sum(gs, red);
    if (tid == 0) {
      st->ghat_sq = gs;
      st->ghat_norm = sqrt(gs);
      st->alpha_valid = 0;
    }
  }
  if (s_flag[2]) {
    double* LM = v.W(W_LMD);
    if (P.strategy == SWGN_LEVENBERG_MARQUARDT) {
      const double ir = 1.0 / sqrt(st->radius);  // lm_diagonal = sqrt(diagonal / radius), :86
      const double* SC = v.W(W_SCALE);
      for (int k = tid; k < d.n_t; k += kThreads) LM[k] = P.jacobi_scaling ? DG[k] * ir / SC[k] : DG[k] * ir;
    } else {
      const double sm = sqrt(st->mu);
      for (int k = tid; k < d.n_t; k += kThreads) LM[k] = DG[k] * sm;
    }
  }
  if (s_flag[3] && tid == 0) atomicAdd(b.counters + (tick & 1), 1);
}

// ---------------------------------------------------------------------------------------------
// k_step: traditional dogleg step (dogleg_strategy.cc:199-253), model cost change
// (trust_region_minimizer.cc:414-431), invalid-step handling (:453-486) and the candidate point
// x [+] step (:761-779).
// ---------------------------------------------------------------------------------------------
// 6 CTAs/SM (40 registers, 0.3 KB of spills in the scalar dogleg code of thread 0): the row pass is bound by the
// latency of its index chain, measured 213.4 / 214.1 / 215.0 / 214.6 k it/s at 4 / 5 / 6 / 8 CTAs per SM
#ifndef SWGN_STEP_CTAS
#define SWGN_STEP_CTAS 6
#endif
__global__ void __launch_bounds__(kThreads, SWGN_STEP_CTAS) k_step(DeviceBatch b) {
  __shared__ WinDesc sd;
  __shared__ double red[33];
  __shared__ double s_c[4];
  __shared__ int s_case;
  const int w = blockIdx.x;
  TRState* st = b.state + w;
  if (!st->active || st->need_solve) return;
  const Win v = load_window(b, w, &sd);
  const WinDesc& d = sd;
  const SolverParams& P = b.params;
  const int tid = threadIdx.x;
  int valid = 0;
  if (st->solve_ok && P.strategy == SWGN_LEVENBERG_MARQUARDT) {
    // step = -(linear solution), already in the unscaled space; model cost change (trust_region_minimizer.cc:414-431)
    const double* Y = v.W(W_Y);
    double* STEP = v.W(W_STEP);
    for (int k = tid; k < d.n_t; k += kThreads) STEP[k] = -Y[k];
    __syncthreads();
    const double* R = v.W(W_RES);
    double* MR = v.W(W_MRES);
    double mc = 0.0;
    for (int rs = tid; rs < d.n_res; rs += kThreads) {
      const double m = row_dot(v, rs, STEP);
      MR[rs] = m;
      mc += m * (R[rs] + m / 2.0);
    }
    mc = block_sum(mc, red);
    valid = (-mc > 0.0);
    if (tid == 0) st->model_cost_change = -mc;
  } else if (st->solve_ok) {
    const double* GN = v.W(W_GN);
    const double* GH = v.W(W_GHAT);
    const double* DG = v.W(W_DIAG);
    double* STEP = v.W(W_STEP);
    double gnn2 = 0.0, gdot = 0.0;
    for (int k = tid; k < d.n_t; k += kThreads) {
      gnn2 += GN[k] * GN[k];
      gdot += GH[k] * GN[k];
    }
    gnn2 = block_sum(gnn2, red);
    gdot = block_sum(gdot, red);
    if (sqrt(gnn2) > st->radius && !st->alpha_valid) {  // ComputeCauchyPoint, deferred from k_begin (uniform per window)
      double* tmp = STEP;
      double* MRc = v.W(W_MRES);
      for (int k = tid; k < d.n_t; k += kThreads) tmp[k] = GH[k] / DG[k];
      __syncthreads();
      double js = 0.0;
      for (int rs = tid; rs < d.n_res; rs += kThreads) {
        const double m = row_dot(v, rs, tmp);
        MRc[rs] = m;
        js += m * m;
      }
      js = block_sum(js, red);
      if (tid == 0) {
        st->alpha = st->ghat_sq / js;
        st->alpha_valid = 1;
      }
      __syncthreads();
    }
    if (tid == 0) {
      const double radius = st->radius, alpha = st->alpha, gnorm = st->ghat_norm, gnn = sqrt(gnn2);
      if (gnn <= radius) {
        s_case = 0;
        s_c[2] = gnn;
      } else if (gnorm * alpha >= radius) {
        s_case = 1;
        s_c[0] = -(radius / gnorm);
        s_c[2] = radius;
      } else {
        const double b_dot_a = -alpha * gdot;
        const double a_sq = pow(alpha * gnorm, 2.0);
        const double bma_sq = a_sq - 2 * b_dot_a + pow(gnn, 2.0);
        const double c = b_dot_a - a_sq;
        const double dd = sqrt(c * c + bma_sq * (pow(radius, 2.0) - a_sq));
        const double beta = (c <= 0) ? (dd - c) / bma_sq : (radius * radius - a_sq) / (dd + c);
        s_case = 2;
        s_c[0] = -alpha * (1.0 - beta);
        s_c[1] = beta;
      }
    }
    __syncthreads();
    const int cs = s_case;
    double sn2 = 0.0;
    for (int k = tid; k < d.n_t; k += kThreads) {
      double s;
      if (cs == 0) s = GN[k];
      else if (cs == 1) s = s_c[0] * GH[k];
      else s = s_c[0] * GH[k] + s_c[1] * GN[k];
      sn2 += s * s;
      STEP[k] = s / DG[k];
    }
    sn2 = block_sum(sn2, red);  // (barrier: STEP is complete)
    const double* R = v.W(W_RES);
    double* MR = v.W(W_MRES);
    double mc = 0.0;
    for (int rs = tid; rs < d.n_res; rs += kThreads) {
      const double m = row_dot(v, rs, STEP);
      MR[rs] = m;
      mc += m * (R[rs] + m / 2.0);
    }
    mc = block_sum(mc, red);
    valid = (-mc > 0.0);
    if (tid == 0) {
      st->dogleg_step_norm = (cs == 2) ? sqrt(sn2) : s_c[2];
      st->model_cost_change = -mc;
    }
  }
  if (!valid) {  // HandleInvalidStep + DoglegStrategy::StepIsInvalid (:635-638)
    if (tid == 0) {
      st->step_valid = 0;
      st->num_consecutive_invalid += 1;
      if (st->num_consecutive_invalid >= P.max_num_consecutive_invalid_steps) {
        st->active = 0;
        st->termination = SWGN_FAILURE;
      } else {
        if (P.strategy == SWGN_LEVENBERG_MARQUARDT) {  // StepIsInvalid = StepRejected (levenberg_marquardt_strategy.h:61-67)
          st->radius = st->radius / st->decrease_factor;
          st->decrease_factor *= 2.0;
        } else {
          st->mu *= 10.0;
          st->reuse = 0;
        }
        st->iter_cost = st->x_cost + st->fixed_cost;
        st->last_successful = 0;
      }
    }
    return;
  }
  // candidate point and |x - x_candidate|
  const double* x = v.W(W_X);
  double* xc = v.W(W_XCAND);
  const double* STEP = v.W(W_STEP);
  const int32_t* col_state = v.I(I_COL_STATE);
  const int32_t* col_gsize = v.I(I_COL_GSIZE);
  const int32_t* col_size = v.I(I_COL_SIZE);
  const int32_t* col_pos = v.I(I_COL_POS);
  double dn2 = 0.0;
  for (int c = tid; c < d.n_cols; c += kThreads) {
    const int so = col_state[c], gs = col_gsize[c];
    block_plus(x + so, STEP + col_pos[c], xc + so, gs, col_size[c]);
    for (int k = 0; k < gs; ++k) {
      const double dx = x[so + k] - xc[so + k];
      dn2 += dx * dx;
    }
  }
  dn2 = block_sum(dn2, red);
  if (tid == 0) {
    st->step_valid = 1;
    st->num_consecutive_invalid = 0;
    st->step_norm = sqrt(dn2);
  }
}

// ---------------------------------------------------------------------------------------------
// k_end: convergence tests on the candidate (:706-748), step quality, accept / reject (:781-826),
// DoglegStrategy::StepAccepted / StepRejected (:612-633), TrustRegionStepEvaluator::StepAccepted.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_end(DeviceBatch b) {
  __shared__ WinDesc sd;
  __shared__ int s_accept;
  const int w = blockIdx.x;
  TRState* st = b.state + w;
  if (!st->active || st->need_solve || !st->step_valid) return;
  const SolverParams& P = b.params;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int accept = 0;
    TRState s = *st;
    const double cost_change = s.x_cost - s.candidate_cost;
    if (s.step_norm <= P.parameter_tolerance * (s.x_norm + P.parameter_tolerance)) {
      s.active = 0;
      s.termination = SWGN_CONVERGENCE;
    } else if (fabs(cost_change) <= P.function_tolerance * s.x_cost) {
      s.active = 0;
      s.termination = SWGN_CONVERGENCE;
    } else {
      double q;
      if (s.candidate_cost >= DBL_MAX) {
        q = -DBL_MAX;
      } else {
        const double rel = (s.se_cur - s.candidate_cost) / s.model_cost_change;
        const double hist = (s.se_ref - s.candidate_cost) / (s.se_acc_ref + s.model_cost_change);
        q = fmax(rel, hist);
      }
      s.relative_decrease = q;
      if (q > P.min_relative_decrease) {
        accept = 1;
        s.accepted = 1;
        if (P.strategy == SWGN_LEVENBERG_MARQUARDT) {  // LevenbergMarquardtStrategy::StepAccepted :151-158
          s.radius = s.radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * q - 1.0, 3.0));
          s.radius = fmin(P.max_radius_lm, s.radius);
          s.decrease_factor = 2.0;
        } else {
          if (q < 0.25) s.radius *= 0.5;
          if (q > 0.75) s.radius = fmax(s.radius, 3.0 * s.dogleg_step_norm);
          s.mu = fmax(P.min_mu, 2.0 * s.mu / 10.0);
          s.reuse = 0;
        }
        s.se_cur = s.candidate_cost;
        s.se_acc_cand += s.model_cost_change;
        s.se_acc_ref += s.model_cost_change;
        if (s.se_cur < s.se_min) {
          s.se_min = s.se_cur;
          s.se_nonmono = 0;
          s.se_cand = s.se_cur;
          s.se_acc_cand = 0.0;
        } else {
          s.se_nonmono += 1;
          if (s.se_cur > s.se_cand) {
            s.se_cand = s.se_cur;
            s.se_acc_cand = 0.0;
          }
        }
        if (s.se_nonmono == 0) {
          s.se_ref = s.se_cand;
          s.se_acc_ref = s.se_acc_cand;
        }
        // The step that exhausts max_num_iterations: Ceres still evaluates residuals and Jacobian at the new point
        // (HandleSuccessfulStep) and then stops in FinalizeIterationAndCheckIfMinimizerCanContinue, which tests the
        // iteration count BEFORE the gradient (trust_region_minimizer.cc:303-330) -- nothing that evaluation produces is
        // read again.  Its cost is the candidate cost (same point, same code), so the evaluation is skipped here, unless
        // the window holds stateful or host-evaluated factors (IMUGNSSFactor back-substitutes its hidden states during a
        // Jacobian evaluation).  The one observable difference: a non-finite Jacobian at that last point (finite
        // residuals) ends as NO_CONVERGENCE with the accepted state instead of FAILURE.
        if (s.iteration >= P.max_num_iterations && b.desc[w].n_chain == 0 && b.desc[w].n_host == 0) {
          s.accepted = 0;  // k_eval(accepted) does not run
          s.x_cost = s.candidate_cost;
          s.iter_cost = s.x_cost + s.fixed_cost;
          s.last_successful = 1;
        }
      } else {
        s.last_successful = 0;
        s.iter_cost = s.candidate_cost + s.fixed_cost;
        if (P.strategy == SWGN_LEVENBERG_MARQUARDT) {  // StepRejected :160-164
          s.radius = s.radius / s.decrease_factor;
          s.decrease_factor *= 2.0;
        } else {
          s.radius *= 0.5;
          s.reuse = 1;
        }
      }
    }
    *st = s;
    s_accept = accept;
  }
  __syncthreads();
  if (!s_accept) return;
  const Win v = load_window(b, w, &sd);
  const double* xc = v.W(W_XCAND);
  double* x = v.W(W_X);
  for (int k = tid; k < sd.n_state; k += kThreads) x[k] = xc[k];
}

// k_finish: the user-visible state is the lowest-cost point seen, or the original parameters when
// the solve failed (CERES/internal/ceres/solver.cc:444-447).
__global__ void __launch_bounds__(kThreads) k_finish(DeviceBatch b) {
  __shared__ WinDesc sd;
  const int w = blockIdx.x;
  const TRState* st = b.state + w;
  const Win v = load_window(b, w, &sd);
  const double* src = v.W(st->termination == SWGN_FAILURE ? W_X0 : W_XBEST);
  double* x = v.W(W_X);
  double* xc = v.W(W_XCAND);
  for (int k = threadIdx.x; k < sd.n_state; k += kThreads) {
    x[k] = src[k];
    xc[k] = src[k];
  }
}

// pack / unpack the user-visible states of all windows into one staging buffer (one H2D / D2H copy
// for the whole batch instead of one per window)
__global__ void __launch_bounds__(kThreads) k_gather_states(DeviceBatch b, double* stage, const int64_t* offs, int to_device) {
  const int w = blockIdx.x;
  const WinDesc& d = b.desc[w];
  double* x = b.wpool + d.woff[W_X];
  double* st = stage + offs[w];
  for (int k = threadIdx.x; k < d.n_state; k += kThreads) {
    if (to_device) x[k] = st[k];
    else st[k] = x[k];
  }
}
void launch_gather_states(const DeviceBatch& b, double* stage, const int64_t* offs, int to_device, cudaStream_t s) {
  k_gather_states<<<b.n_windows, kThreads, 0, s>>>(b, stage, offs, to_device);
}

void launch_begin(const DeviceBatch& b, int tick, cudaStream_t s) { k_begin<<<b.n_windows, kThreads, 0, s>>>(b, tick); }
void launch_step(const DeviceBatch& b, cudaStream_t s) { k_step<<<b.n_windows, kThreads, 0, s>>>(b); }
void launch_end(const DeviceBatch& b, cudaStream_t s) { k_end<<<b.n_windows, kThreads, 0, s>>>(b); }
void launch_finish(const DeviceBatch& b, cudaStream_t s) { k_finish<<<b.n_windows, kThreads, 0, s>>>(b); }

}  // namespace swgn
