// Launch wrappers of the solver kernels (kernels.cu).  One CTA per window everywhere: windows are
// the unit of data parallelism, a batch is processed by a grid of n_windows CTAs.
#pragma once
#include <cuda_runtime.h>

#include "device_types.h"

namespace swgn {

struct DeviceBatch {
  int n_windows;
  const WinDesc* desc;   // [n_windows]
  const int32_t* ipool;
  const double* cpool;
  double* wpool;
  TRState* state;        // [n_windows]
  int32_t* counters;     // [4]: 0 = windows still active, 1 = windows needing a solve retry
  SolverParams params;
  int max_buf;           // max over windows of the chunk buffer size (doubles)
  int max_nf;            // max reduced-system size
  int max_prior_n;
};

enum EvalMode { EVAL_INIT = 0, EVAL_ACCEPTED = 1, EVAL_CANDIDATE = 2, EVAL_FORCE = 3 };

size_t schur_smem_bytes(const DeviceBatch& b, int* chunk_warps);
size_t chol_smem_bytes(const DeviceBatch& b);
size_t eval_smem_bytes(const DeviceBatch& b);
cudaError_t configure_kernels(const DeviceBatch& b);

void launch_init(const DeviceBatch& b, cudaStream_t s);
void launch_eval(const DeviceBatch& b, int mode, cudaStream_t s);
void launch_grad(const DeviceBatch& b, int mode, cudaStream_t s);
void launch_begin(const DeviceBatch& b, cudaStream_t s);
void launch_schur(const DeviceBatch& b, int force, cudaStream_t s);
void launch_chol(const DeviceBatch& b, int force, cudaStream_t s);
void launch_backsub(const DeviceBatch& b, int force, cudaStream_t s);
void launch_step(const DeviceBatch& b, cudaStream_t s);
void launch_end(const DeviceBatch& b, cudaStream_t s);
void launch_finish(const DeviceBatch& b, cudaStream_t s);
// staged helpers for tests
void launch_set_lm_diagonal(const DeviceBatch& b, int window, const double* D_dev, cudaStream_t s);

}  // namespace swgn
