// TEST INFRASTRUCTURE -- see oracle_core.h.  C entry points of liboracle.so for ctypes (tests,
// smoke(), bench.py's cpu_baseline / --impl reference legs).
#include <chrono>
#include <cstdlib>

#include "oracle_core.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

extern "C" {

void* oracle_create(const swgn_graph* g, const swgn_options* o) {
  Solver* s = new Solver();
  if (!s->Build(g, o) || !s->Preprocess()) {
    std::fprintf(stderr, "oracle_create: %s\n", s->error.c_str());
    delete s;
    return nullptr;
  }
  return s;
}
void oracle_destroy(void* h) { delete (Solver*)h; }

// dims: [n_residuals, n_cols(tangent), n_col_blocks, n_eliminate_blocks, n_row_blocks, n_f, n_e,
//        n_state]
void oracle_dims(void* h, int32_t* dims) {
  Solver* s = (Solver*)h;
  dims[0] = s->num_residuals;
  dims[1] = s->num_effective_parameters;
  dims[2] = (int)s->pblocks.size();
  dims[3] = s->num_eliminate_blocks;
  dims[4] = (int)s->rblocks.size();
  dims[5] = s->eliminator.lhs_num_rows;
  dims[6] = s->num_effective_parameters - s->eliminator.lhs_num_rows;
  dims[7] = (int)s->state.size();
}

void oracle_columns(void* h, int32_t* block, int32_t* offset, int32_t* size) {
  Solver* s = (Solver*)h;
  for (size_t i = 0; i < s->pblocks.size(); ++i) {
    block[i] = s->pblocks[i]->graph_index;
    offset[i] = s->pblocks[i]->delta_offset;
    size[i] = s->pblocks[i]->local;
  }
}
void oracle_rows(void* h, int32_t* factor, int32_t* offset) {
  Solver* s = (Solver*)h;
  for (size_t i = 0; i < s->rblocks.size(); ++i) {
    factor[i] = s->rblocks[i]->program_index;
    offset[i] = s->residual_layout[i];
  }
}

static void gather_x(Solver* s, std::vector<double>* x) {
  x->resize(s->num_parameters);
  for (ParamBlock* pb : s->pblocks)
    std::memcpy(x->data() + pb->state_offset, pb->user_state, sizeof(double) * pb->size);
}

// evaluate at the current user state; any output may be NULL.  J is dense row-major
// n_residuals x n_cols.
int oracle_evaluate(void* h, double* cost, double* residuals, double* gradient, double* J) {
  Solver* s = (Solver*)h;
  std::vector<double> x;
  gather_x(s, &x);
  std::vector<double> r(s->num_residuals), g(s->num_effective_parameters);
  double c = 0.0;
  if (!s->Evaluate(x.data(), &c, r.data(), g.data(), true)) return 1;
  if (cost) *cost = c;
  if (residuals) std::memcpy(residuals, r.data(), sizeof(double) * r.size());
  if (gradient) std::memcpy(gradient, g.data(), sizeof(double) * g.size());
  if (J) {
    const int nc = s->num_effective_parameters;
    std::fill(J, J + (size_t)s->num_residuals * nc, 0.0);
    for (const RowBlock& row : s->jac.rows)
      for (const Cell& cell : row.cells) {
        const int cs = s->jac.cols[cell.block_id].size, cp = s->jac.cols[cell.block_id].position;
        for (int rr = 0; rr < row.size; ++rr)
          for (int k = 0; k < cs; ++k)
            J[(size_t)(row.position + rr) * nc + cp + k] = s->jac.values[cell.position + rr * cs + k];
      }
  }
  return 0;
}

// MarginalizationInfo::setmarginalizeinfo(..., Sqrt = true)  RVI/factor/marginalization_factor.cpp:449-475:
// J0 = sqrt(S) V', r0 = S^-1/2 V' b from the eigen-decomposition of A (threshold eps = 1e-8)
void oracle_prior_sqrt(const double* A, const double* b, int n, double* J0, double* r0) {
  Mat Am(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Am(i, j) = A[(size_t)i * n + j];
  std::vector<double> w;
  Mat V;
  eig_sym(Am, &w, &V);
  for (int i = 0; i < n; ++i) {
    const double S = w[i] > 1e-8 ? w[i] : 0.0, Sinv = w[i] > 1e-8 ? 1.0 / w[i] : 0.0;
    double dot = 0.0;
    for (int c = 0; c < n; ++c) {
      J0[(size_t)i * n + c] = std::sqrt(S) * V(c, i);
      dot += V(c, i) * b[c];
    }
    r0[i] = std::sqrt(Sinv) * dot;
  }
}

// batched IMU pre-integration, factor f = samples [begin[f], begin[f+1]); returns the number of failures
int oracle_preintegrate_batch(int n, const int32_t* begin, const double* samples, const double* bias, const double* noise4,
                              double* records) {
  int bad = 0;
  for (int f = 0; f < n; ++f)
    if (!preintegrate(begin[f + 1] - begin[f], samples + (size_t)7 * begin[f], bias + 6 * f, noise4,
                      records + (size_t)SWGN_IMU_STRIDE * f))
      ++bad;
  return bad;
}

// cost-only evaluation (what the trust-region loop does for a candidate point: no Jacobians are
// requested from the cost functions, trust_region_minimizer.cc:761-787)
int oracle_evaluate_cost(void* h, double* cost, double* residuals) {
  Solver* s = (Solver*)h;
  std::vector<double> x;
  gather_x(s, &x);
  std::vector<double> r(s->num_residuals);
  double c = 0.0;
  if (!s->Evaluate(x.data(), &c, r.data(), nullptr, false)) return 1;
  if (cost) *cost = c;
  if (residuals) std::memcpy(residuals, r.data(), sizeof(double) * r.size());
  return 0;
}

// one DENSE_SCHUR linear solve on the linearisation at the current user state
int oracle_linear_solve(void* h, const double* D, double* x_out, double* S, double* rhs) {
  Solver* s = (Solver*)h;
  std::vector<double> x;
  gather_x(s, &x);
  std::vector<double> r(s->num_residuals);
  double c;
  if (!s->Evaluate(x.data(), &c, r.data(), nullptr, true)) return 1;
  std::vector<double> y(s->num_effective_parameters);
  bool exported_only = false;
  bool ok = s->LinearSolve(r.data(), D, y.data(), &exported_only);
  const int n = s->exports.hs_row;
  if (S) std::memcpy(S, s->exports.lhs_out.data(), sizeof(double) * n * n);
  if (rhs) std::memcpy(rhs, s->exports.rhs_out.data(), sizeof(double) * n);
  if (x_out) std::memcpy(x_out, y.data(), sizeof(double) * y.size());
  return ok ? 0 : 2;
}

int oracle_minimize(void* h, swgn_summary* summary) {
  Solver* s = (Solver*)h;
  swgn_summary local;
  if (!summary) summary = &local;
  std::memset(summary, 0, sizeof(*summary));
  return s->Minimize(summary) ? 0 : 1;
}

void oracle_get_state(void* h, double* state) {
  Solver* s = (Solver*)h;
  std::memcpy(state, s->state.data(), sizeof(double) * s->state.size());
}
void oracle_set_state(void* h, const double* state) {
  Solver* s = (Solver*)h;
  std::memcpy(s->state.data(), state, sizeof(double) * s->state.size());
}
// current hidden states of every IMUGNSSFactor chain (16 doubles per hidden frame), in the order
// the chains were given in the graph; returns the total number of hidden frames
int oracle_chain_frames(void* h, double* out) {
  Solver* s = (Solver*)h;
  // residual_blocks is in program order; recover graph order through a stable scan per chain kind
  int total = 0;
  for (auto& rb : s->residual_blocks) {
    const int m = chain_factor_frames(rb->cost.get(), out ? out + 16 * total : nullptr);
    total += m;
  }
  return total;
}

int oracle_num_iteration_records(void* h) { return (int)((Solver*)h)->iterations.size(); }
void oracle_iteration_records(void* h, double* cost, double* radius, int32_t* successful) {
  Solver* s = (Solver*)h;
  for (size_t i = 0; i < s->iterations.size(); ++i) {
    cost[i] = s->iterations[i].cost;
    radius[i] = s->iterations[i].radius;
    successful[i] = s->iterations[i].step_is_successful;
  }
}
// exports of the last linear solve: S, r (lhs_out/rhs_out) and L (lhs_out2); returns hs_row
int oracle_get_exports(void* h, double* S, double* r, double* L) {
  Solver* s = (Solver*)h;
  const int n = s->exports.hs_row;
  if (S && !s->exports.lhs_out.empty())
    std::memcpy(S, s->exports.lhs_out.data(), sizeof(double) * n * n);
  if (r && !s->exports.rhs_out.empty()) std::memcpy(r, s->exports.rhs_out.data(), sizeof(double) * n);
  if (L && s->exports.have_factor) std::memcpy(L, s->exports.lhs_out2.data(), sizeof(double) * n * n);
  return n;
}

// UpdateSchurHessianOnly  RVI/swf/swf_gnss.cpp:65-94: A = L_nn L_nn^T
void oracle_tail_information(const double* L, int n, int n_tail, double* A) {
  const int m = n - n_tail;
  for (int i = 0; i < n_tail; ++i)
    for (int j = 0; j < n_tail; ++j) {
      double s = 0.0;
      for (int k = 0; k < n_tail; ++k)
        s += L[(size_t)(m + i) * n + m + k] * L[(size_t)(m + j) * n + m + k];
      A[(size_t)i * n_tail + j] = s;
    }
}

// UpdateSchur  RVI/swf/swf_gnss.cpp:25-61: Schur-reduce the leading m rows of the exported
// (S, r) with an eigen pseudo-inverse (threshold 1e-8)
void oracle_update_schur(const double* S, const double* r, int n, int n_tail, double* A,
                         double* b) {
  const int m = n - n_tail;
  Mat full(n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) full(i, j) = (j >= i) ? S[(size_t)i * n + j] : S[(size_t)j * n + i];
  if (m == 0) {
    std::memcpy(A, full.data(), sizeof(double) * n * n);
    std::memcpy(b, r, sizeof(double) * n);
    return;
  }
  Mat Amm(m, m);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) Amm(i, j) = full(i, j);
  std::vector<double> w;
  Mat V;
  eig_sym(Amm, &w, &V);
  Mat inv(m, m);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) {
      double s = 0.0;
      for (int k = 0; k < m; ++k)
        if (w[k] > 1e-8) s += V(i, k) * (1.0 / w[k]) * V(j, k);
      inv(i, j) = s;
    }
  Mat Anm(n_tail, m);
  for (int i = 0; i < n_tail; ++i)
    for (int j = 0; j < m; ++j) Anm(i, j) = full(m + i, j);
  Mat T = matmul(Anm, inv);  // n_tail x m
  for (int i = 0; i < n_tail; ++i) {
    for (int j = 0; j < n_tail; ++j) {
      double s = 0.0;
      for (int k = 0; k < m; ++k) s += T(i, k) * full(k, m + j);
      A[(size_t)i * n_tail + j] = full(m + i, m + j) - s;
    }
    double s = 0.0;
    for (int k = 0; k < m; ++k) s += T(i, k) * r[k];
    b[i] = r[m + i] - s;
  }
}

// Schur elimination of a raw block-sparse system (golden fixtures of
// CERES/internal/ceres/linear_least_squares_problems.cc).  Cells of row i are
// cell_col[row_ptr[i]..row_ptr[i+1]) sorted by column block with values (row-major
// row_size x col_size) stored consecutively in `values` in that same order.
int oracle_schur_raw(int n_col_blocks, const int32_t* col_sizes, int n_row_blocks,
                     const int32_t* row_sizes, const int32_t* row_ptr, const int32_t* cell_col,
                     const double* values, const double* b, const double* D, int num_eliminate,
                     double* lhs, double* rhs, double* x) {
  BlockSparse A;
  A.cols.resize(n_col_blocks);
  int pos = 0;
  for (int i = 0; i < n_col_blocks; ++i) {
    A.cols[i].size = col_sizes[i];
    A.cols[i].position = pos;
    pos += col_sizes[i];
  }
  A.num_cols = pos;
  A.rows.resize(n_row_blocks);
  int rpos = 0, vpos = 0;
  for (int i = 0; i < n_row_blocks; ++i) {
    A.rows[i].size = row_sizes[i];
    A.rows[i].position = rpos;
    rpos += row_sizes[i];
    for (int k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
      Cell c;
      c.block_id = cell_col[k];
      c.position = vpos;
      vpos += row_sizes[i] * col_sizes[cell_col[k]];
      A.rows[i].cells.push_back(c);
    }
  }
  A.num_rows = rpos;
  A.values.assign(values, values + vpos);
  SchurEliminator el;
  el.Init(num_eliminate, A);
  const int n = el.lhs_num_rows;
  std::vector<double> S((size_t)n * n), r(n);
  el.Eliminate(A, b, D, S.data(), r.data());
  if (lhs) std::memcpy(lhs, S.data(), sizeof(double) * n * n);
  if (rhs) std::memcpy(rhs, r.data(), sizeof(double) * n);
  if (x) {
    std::fill(x, x + A.num_cols, 0.0);
    double* reduced = x + A.num_cols - n;
    std::vector<double> u = S;
    if (!llt_upper_inplace(u.data(), n)) return 2;
    std::copy(r.begin(), r.end(), reduced);
    llt_upper_solve(u.data(), n, reduced);
    el.BackSubstitute(A, b, D, reduced, x);
  }
  return 0;
}

int oracle_lambda(int n, int m, const double* a, const double* Q, double* F, double* s) {
  return lambda_rtk(n, m, a, Q, F, s);
}
int oracle_matinv(double* A, int n) { return matinv_rtk(A, n); }
// InvertPSDMatrix<Dynamic>(assume_full_rank = true, m) as the Schur eliminator calls it; 1 = factorisation succeeded
int oracle_invert_psd(const double* m, int n, double* inv) { return invert_psd(m, n, inv) ? 1 : 0; }
int oracle_ambiguity_fix(int n, const double* A, const double* y, int n_epochs,
                         const int32_t* epoch_begin, const int32_t* obs_amb,
                         const int32_t* obs_sysfreq, int last_fix, int32_t* dd_pairs, double* F,
                         swgn_fix_result* res) {
  return ambiguity_fix(n, A, y, n_epochs, epoch_begin, obs_amb, obs_sysfreq, last_fix, dd_pairs,
                       F, res);
}
double oracle_distance(const double* rr, const double* rs, double* e) {
  return distance_rtk(rr, rs, e);
}
double oracle_velocity_distance(const double* rr, const double* rs, const double* vr,
                                const double* vs, double* e) {
  return velocity_distance_rtk(rr, rs, vr, vs, e);
}
double oracle_varerr2(double el, double dt, double var) { return varerr2(el, dt, var); }
void oracle_pose_plus(const double* x, const double* d, double* out) { pose_plus(x, d, out); }
void oracle_cauchy(double a, double s, double* rho) { cauchy_loss(a, s, rho); }

// Evaluate one factor kind directly (Jacobian checks): kind 0 proj, 1 imu, 2 gnss(kind2), 3 unit.
// params: concatenated global blocks in factor order; jac_out: concatenated row-major global
// Jacobians.
int oracle_factor_eval(int kind, int kind2, const double* globals /*Pbg3,g3,W4*/,
                       const double* record, const double* params, double* residuals,
                       double* jac_out) {
  AppGlobals g;
  for (int i = 0; i < 3; ++i) {
    g.Pbg[i] = globals[i];
    g.gravity[i] = globals[3 + i];
  }
  for (int i = 0; i < 4; ++i) g.proj_sqrt_info[i] = globals[6 + i];
  std::unique_ptr<CostFunction> f;
  if (kind == 0) f.reset(make_projection_factor(&g, record));
  else if (kind == 1) f.reset(make_imu_factor(&g, record));
  else if (kind == 2) f.reset(make_gnss_factor(kind2, record));
  else f.reset(make_unit_factor(record[0]));
  std::vector<const double*> p;
  std::vector<double*> J;
  const double* pp = params;
  double* jp = jac_out;
  for (int sz : f->block_sizes) {
    p.push_back(pp);
    pp += sz;
    J.push_back(jp);
    if (jp) jp += (size_t)f->num_residuals * sz;
  }
  return f->Evaluate(p.data(), residuals, jac_out ? J.data() : nullptr) ? 0 : 1;
}

// The restated MarginalizationFactor (make_prior_factor) evaluated directly: residuals (n) and the row-major
// n x global-size Jacobians back to back -- the same signature as oracle/ref_marg_shim.cpp's ref_prior_eval.
int oracle_prior_eval(int n_keep, int n, const int32_t* keep_size, const int32_t* keep_idx, const double* x0, const double* J0,
                      const double* r0, const double* x, double* residuals, double* jac_out) {
  std::vector<int> sizes(keep_size, keep_size + n_keep), idx(keep_idx, keep_idx + n_keep);
  std::unique_ptr<CostFunction> f(make_prior_factor(n, sizes, idx, x0, J0, r0));
  std::vector<const double*> p;
  std::vector<double*> J;
  const double* pp = x;
  double* jp = jac_out;
  for (int sz : sizes) {
    p.push_back(pp);
    pp += sz;
    J.push_back(jp);
    if (jp) jp += (size_t)n * sz;
  }
  return f->Evaluate(p.data(), residuals, jac_out ? J.data() : nullptr) ? 0 : 1;
}

// CPU baseline: solve n windows, one window per OpenMP thread; returns wall seconds of the
// minimise phase only (preprocessing excluded, like minimizer_time_in_seconds) and the total
// number of trust-region iterations executed.
double oracle_solve_batch_timed(int n, const swgn_graph* const* graphs, const swgn_options* o,
                                int threads, int64_t* total_iterations, double* states_out,
                                int64_t state_stride) {
  std::vector<Solver*> solvers(n, nullptr);
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < n; ++i) {
    Solver* s = new Solver();
    if (s->Build(graphs[i], o) && s->Preprocess()) solvers[i] = s;
    else delete s;
  }
  int64_t iters = 0;
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic) reduction(+ : iters)
  for (int i = 0; i < n; ++i) {
    if (!solvers[i]) continue;
    swgn_summary sm;
    std::memset(&sm, 0, sizeof(sm));
    solvers[i]->Minimize(&sm);
    iters += sm.num_iterations;
  }
  auto t1 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; ++i) {
    if (solvers[i] && states_out)
      std::memcpy(states_out + (size_t)i * state_stride, solvers[i]->state.data(),
                  sizeof(double) * solvers[i]->state.size());
    delete solvers[i];
  }
  if (total_iterations) *total_iterations = iters;
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
