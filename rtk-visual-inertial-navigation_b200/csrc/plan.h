// Host-side preprocessing of one window: what Ceres does in TrustRegionPreprocessor::Preprocess
// on every Solve (CERES/internal/ceres/trust_region_preprocessor.cc:360-393) -- reduced program,
// ordering, e-block independence check, lexicographic row order, block-sparse structure, Schur
// chunks -- flattened into the index arrays the kernels consume (device_types.h).
#pragma once
#include <string>
#include <vector>

#include "../../include/swgn.h"
#include "device_types.h"

namespace swgn {

// result of build_stream_plan (plan_stream.cpp): sizes of the on-chip layout of the streamed Schur elimination
struct StreamPlanInfo {
  int ok = 0, fits = 0, nbatch = 0, acc = 0, jcap = 0, rcap = 0, ecap = 0, fcap = 0, reccap = 0, n_fb = 0;
};

struct WindowPlan {
  WinDesc d;                               // counts filled; offsets filled by the batch
  std::vector<int32_t> iarr[NUM_IARR];
  std::vector<double> carr[NUM_CARR];
  int64_t wsize[NUM_WARR];
  std::vector<double> state;               // initial state (graph layout)
  // algorithmic traffic of one Schur elimination on the materialised Jacobian, in doubles
  // (SURVEY.md 8d formula): J blocks at their stored size + residuals + D in, S upper + r + y out
  int64_t schur_doubles;
  int64_t n_mma = 0;                       // tensor-core MMAs of one Schur gather pass (planning statistic)
  StreamPlanInfo sb;
  bool want_stream_plan = false;           // plan the streamed Schur elimination even when it is not enabled (probes, tests)
};
bool stream_enabled();                     // SWGN_SCHUR_STREAM=1

// dynamic shared memory of k_schur_stream for a window (or a batch: pass the maxima) with these sizes
size_t stream_smem_bytes(int nbatch, int acc, int jcap, int rcap, int ecap, int fcap, int reccap);
// streamed Schur plan of one window: fills I_SB_HDR, I_SB_REC, I_ACC_MAP and P->sb from the row / chunk / slot tables
void build_stream_plan(WindowPlan* P, int n_rows, int n_cols, int n_ecols, int n_jac, int n_res, const std::vector<int>& col_size,
                       const std::vector<int>& col_pos);

// sizes (doubles) and packing of the factor constants in device layout; used at plan time and by
// swgn_batch_update_inputs (same structure, new measurements)
void constant_sizes(const swgn_graph* g, int64_t sizes[NUM_CARR]);
void pack_constants(const swgn_graph* g, double* const dst[NUM_CARR]);

// returns SWGN_OK or an error status with *err filled
swgn_status build_plan(const swgn_graph* g, int n_parameter_head, WindowPlan* out, std::string* err);

}  // namespace swgn
