// Launch wrappers of the solver kernels.  One CTA per window everywhere: windows are the unit of
// data parallelism, a batch is processed by a grid of n_windows CTAs and every
// kernel masks windows whose trust-region state machine does not need that stage in this tick.
#pragma once
#include <cuda_runtime.h>

#include "device_types.h"

struct swgn_fix_result;

namespace swgn {

struct DeviceBatch {
  int n_windows;
  const WinDesc* desc;   // [n_windows]
  const int32_t* ipool;
  const double* cpool;
  double* wpool;
  TRState* state;        // [n_windows]
  int32_t* counters;     // [2]: windows still active after k_begin (ping-pong by tick parity)
  SolverParams params;
  int max_wbuf;          // max over windows of the warp-chunk scratch (doubles)
  int max_nf;            // max reduced-system size
  int max_prior_n;
  int max_chain;         // max IMUGNSSFactor chains per window (0: k_chain is never launched)
  int max_chain_k;       // max phase biases per chain (shared-memory size of k_chain)
  int chain_epoch;       // bumped by create / update_inputs: chains reload their hidden states and forget history
  int sb_windows;        // windows whose Schur elimination runs streamed (k_schur_stream); the others run k_schur
  int gather_windows;
  unsigned sb_smem;      // dynamic shared memory of k_schur_stream (max over the streamed windows)
  int keep_copy;         // copy S|rhs to W_SCOPY before factorising (staged test entry point)
  long long* debug;      // optional [n_windows * 8] phase timestamps of k_schur (SWGN_DEBUG_TIMELINE=1), else null
};

enum EvalMode { EVAL_INIT = 0, EVAL_ACCEPTED = 1, EVAL_CANDIDATE = 2, EVAL_FORCE = 3 };
// stages of the staged (test) entry points bypass the state machine for one window
enum { RUN_STATE_MACHINE = -1 };

cudaError_t configure_kernels(const DeviceBatch& b);

// evaluation of every factor; launches k_chain first when the batch holds IMUGNSSFactor chains
void launch_eval(const DeviceBatch& b, int mode, int only_window, cudaStream_t s);
void launch_chain(const DeviceBatch& b, int mode, int only_window, cudaStream_t s);
cudaError_t configure_chain(const DeviceBatch& b);
void launch_begin(const DeviceBatch& b, int tick, cudaStream_t s);
void launch_schur(const DeviceBatch& b, int only_window, cudaStream_t s);         // dispatches to the two kernels below
void launch_schur_gather(const DeviceBatch& b, int only_window, cudaStream_t s);
void launch_schur_stream(const DeviceBatch& b, int only_window, cudaStream_t s);
cudaError_t configure_schur_stream(const DeviceBatch& b);
void launch_chol(const DeviceBatch& b, int only_window, cudaStream_t s);
void launch_backsub(const DeviceBatch& b, int only_window, cudaStream_t s);
void launch_step(const DeviceBatch& b, cudaStream_t s);
void launch_end(const DeviceBatch& b, cudaStream_t s);
void launch_finish(const DeviceBatch& b, cudaStream_t s);
// UpdateSchur (RVI/swf/swf_gnss.cpp:25-61): eigen pseudo-inverse Schur reduction of the exported (S, r)
size_t head_marginal_scratch_doubles(int m, int n);
cudaError_t launch_head_marginal(const DeviceBatch& b, int window, int n_f, int n, double* A_dev, double* b_dev, double* scratch,
                                 cudaStream_t s);
// MarginalizationInfo::setmarginalizeinfo (marginalization_factor.cpp:449-475): (A, b) -> prior factor (J0, r0)
size_t prior_sqrt_scratch_doubles(int n);
cudaError_t launch_prior_sqrt(const double* A_dev, const double* b_dev, int n, double* J0_dev, double* r0_dev, double* scratch,
                              cudaStream_t s);
// UpdateSchur + setmarginalizeinfo for every window (n_tail[w] = 0 skips one); off[4w..4w+3] = offsets of A, b, J0|r0, scratch in buf
cudaError_t launch_marginal_priors(const DeviceBatch& b, int max_m, int max_n, const int32_t* n_tail_dev, const int64_t* off_dev, double* buf,
                                   cudaStream_t s);
void launch_tail_information(const DeviceBatch& b, int window, int n_tail, double* A_dev, cudaStream_t s);

// IMU pre-integration (k_preint.cu): one warp per factor; noise4 = ACC_N, GYR_N, ACC_W, GYR_W (host array)
cudaError_t launch_preintegrate(int n_factors, const int32_t* begin_dev, const double* samples_dev, const double* bias_dev,
                                const double* noise4, double* records_dev, int32_t* status_dev, cudaStream_t s);

// K7/K8: batched RTKLIB-style lambda() and the LambdaSearch decision (k_lambda.cu)
void launch_lambda_batch(int n_problems, int m, const int32_t* n_dev, const int64_t* aoff_dev,
                         const int64_t* qoff_dev, const double* a_dev, const double* Q_dev,
                         double* F_dev, double* s_dev, int32_t* info_dev, double* work_dev,
                         const int64_t* woff_dev, cudaStream_t s);
size_t lambda_work_doubles(int n, int m);
// K7 + K8 for a whole batch: tail information and float ambiguities of every window (one CTA each), then the
// LambdaSearch decision of every window (one thread each)
void launch_tail_information_batch(const DeviceBatch& b, int n_tail, double* A_all, double* y_all, int32_t* have_A, cudaStream_t s);
size_t fix_work_doubles(int n);
size_t fix_work_ints(int n);
void launch_ambiguity_fix_batch(int n_windows, int n, const double* A_all, const double* y_all, const int32_t* win_epoch,
                                const int32_t* epoch_begin, const int32_t* obs_amb, const int32_t* obs_sysfreq, const int32_t* last_fix,
                                const int32_t* have_A, int32_t* dd_pairs, double* F, struct ::swgn_fix_result* res, double* work,
                                int32_t* iwork, cudaStream_t s);

}  // namespace swgn
