// TEST INFRASTRUCTURE -- see oracle_core.h.  Restatement of the modified-Ceres pipeline one
// ceres::Solve() runs for the reference (DENSE_SCHUR + TRADITIONAL_DOGLEG), single-threaded.
#include <algorithm>
#include <numeric>

#include "oracle_core.h"

namespace oracle {

// =============================================================================================
// Program construction from the flat graph (what RVI/swf/*.cpp builds through ceres::Problem)
// =============================================================================================
bool Solver::Build(const swgn_graph* g, const swgn_options* o) {
  opt = *o;
  for (int i = 0; i < 3; ++i) {
    globals.Pbg[i] = g->Pbg[i];
    globals.gravity[i] = g->gravity[i];
  }
  for (int i = 0; i < 4; ++i) globals.proj_sqrt_info[i] = g->proj_sqrt_info[i];
  state.assign(g->state, g->state + g->n_state);
  blocks.resize(g->n_blocks);
  for (int i = 0; i < g->n_blocks; ++i) {
    ParamBlock& b = blocks[i];
    b.graph_index = i;
    b.size = g->block_size[i];
    b.manifold = g->block_manifold[i];
    b.local = (b.manifold == SWGN_MANIFOLD_POSE) ? 6 : b.size;
    b.constant = g->block_const[i] != 0;
    b.group = g->block_group[i];
    b.user_state = state.data() + g->block_offset[i];
  }
  // factors in kind-major storage, then arranged in program order
  std::vector<std::unique_ptr<ResidualBlock>> byk[6];
  for (int i = 0; i < g->n_proj; ++i) {
    auto rb = std::make_unique<ResidualBlock>();
    rb->cost.reset(make_projection_factor(&globals, g->proj_uv + 2 * i));
    rb->cauchy_a = g->proj_cauchy_a;
    for (int k = 0; k < 3; ++k) rb->blocks.push_back(&blocks[g->proj_blocks[3 * i + k]]);
    byk[0].push_back(std::move(rb));
  }
  for (int i = 0; i < g->n_imu; ++i) {
    auto rb = std::make_unique<ResidualBlock>();
    rb->cost.reset(make_imu_factor(&globals, g->imu_data + (size_t)SWGN_IMU_STRIDE * i));
    for (int k = 0; k < 4; ++k) rb->blocks.push_back(&blocks[g->imu_blocks[4 * i + k]]);
    byk[1].push_back(std::move(rb));
  }
  for (int i = 0; i < g->n_gnss; ++i) {
    auto rb = std::make_unique<ResidualBlock>();
    rb->cost.reset(make_gnss_factor(g->gnss_kind[i], g->gnss_data + (size_t)SWGN_GNSS_STRIDE * i));
    for (size_t k = 0; k < rb->cost->block_sizes.size(); ++k)
      rb->blocks.push_back(&blocks[g->gnss_blocks[3 * i + k]]);
    byk[2].push_back(std::move(rb));
  }
  for (int i = 0; i < g->n_prior; ++i) {
    auto rb = std::make_unique<ResidualBlock>();
    std::vector<int> sizes, idx;
    for (int k = g->prior_blk_begin[i]; k < g->prior_blk_begin[i + 1]; ++k) {
      sizes.push_back(g->block_size[g->prior_blocks[k]]);
      idx.push_back(g->prior_blk_idx[k]);
      rb->blocks.push_back(&blocks[g->prior_blocks[k]]);
    }
    rb->cost.reset(make_prior_factor(g->prior_n[i], sizes, idx, g->prior_x0 + g->prior_x0_begin[i],
                                     g->prior_J + g->prior_J_begin[i],
                                     g->prior_r0 + g->prior_r_begin[i]));
    byk[3].push_back(std::move(rb));
  }
  for (int i = 0; i < g->n_unit; ++i) {
    auto rb = std::make_unique<ResidualBlock>();
    rb->cost.reset(make_unit_factor(g->unit_istd[i]));
    rb->blocks.push_back(&blocks[g->unit_block[i]]);
    byk[4].push_back(std::move(rb));
  }
  {  // IMUGNSSFactor chains
    size_t frame_n_off = 0, chain_n_off = 0, imu_off = 0;
    for (int i = 0; i < g->n_chain; ++i) {
      const int b0 = g->chain_blk_begin[i], b1 = g->chain_blk_begin[i + 1];
      const int f0 = g->chain_frame_begin[i], m = g->chain_frame_begin[i + 1] - f0, k = b1 - b0 - 4;
      if (k < 0 || m < 1) {
        error = "bad chain factor";
        return false;
      }
      auto rb = std::make_unique<ResidualBlock>();
      rb->cost.reset(make_chain_factor(&globals, m, k, g->chain_frame_data + (size_t)SWGN_CHAIN_FRAME_STRIDE * f0,
                                       g->chain_frame_N + frame_n_off, g->chain_N + chain_n_off,
                                       g->chain_imu_data + imu_off));
      for (int b = b0; b < b1; ++b) rb->blocks.push_back(&blocks[g->chain_blocks[b]]);
      byk[5].push_back(std::move(rb));
      frame_n_off += (size_t)m * 15 * k;
      chain_n_off += (size_t)k * k + k;
      imu_off += (size_t)(m + 1) * SWGN_IMU_STRIDE;
    }
  }
  if (g->is_use) {
    size_t k = 0;
    for (int kind = 0; kind < 6; ++kind)
      for (auto& rb : byk[kind]) rb->is_use = g->is_use[k++] != 0;
  }
  residual_blocks.clear();
  if (g->order) {
    for (int k = 0; k < g->n_order; ++k) {
      uint32_t kind = g->order[k] >> 28, idx = g->order[k] & 0x0fffffffu;
      if (kind > 5 || idx >= byk[kind].size() || !byk[kind][idx]) {  // (host-evaluated factors, kind 6, are not restated)
        error = "bad program order entry";
        return false;
      }
      residual_blocks.push_back(std::move(byk[kind][idx]));
    }
  } else {
    for (int kind = 0; kind < 6; ++kind)
      for (auto& rb : byk[kind]) residual_blocks.push_back(std::move(rb));
  }
  for (size_t i = 0; i < residual_blocks.size(); ++i) residual_blocks[i]->program_index = (int)i;
  for (auto& rb : residual_blocks)
    for (size_t k = 0; k < rb->blocks.size(); ++k)
      if (rb->blocks[k]->size != rb->cost->block_sizes[k]) {
        error = "parameter block size mismatch";
        return false;
      }
  return true;
}

// =============================================================================================
// ResidualBlock::Evaluate   CERES/internal/ceres/residual_block.cc:69-199
// jacobians[i] (LOCAL size, row-major) may be null per block or entirely
// =============================================================================================
static bool EvaluateResidualBlock(const ResidualBlock& rb, bool apply_loss, double* cost,
                                  double* residuals, double** jacobians,
                                  std::vector<double>* scratch) {
  const int nb = (int)rb.blocks.size();
  const int nr = rb.cost->num_residuals;
  std::vector<const double*> params(nb);
  for (int i = 0; i < nb; ++i) params[i] = rb.blocks[i]->user_state;  // state() == user memory here
  size_t need = nr;
  for (int i = 0; i < nb; ++i) need += (size_t)nr * rb.blocks[i]->size;
  if (scratch->size() < need) scratch->resize(need);
  double* sp = scratch->data();
  std::vector<double*> global_j(nb, nullptr);
  if (jacobians) {
    for (int i = 0; i < nb; ++i) {
      if (jacobians[i] && rb.blocks[i]->manifold != SWGN_MANIFOLD_EUCLIDEAN) {
        global_j[i] = sp;
        sp += (size_t)nr * rb.blocks[i]->size;
      } else {
        global_j[i] = jacobians[i];
      }
    }
  }
  bool outputting = residuals != nullptr;
  if (!outputting) residuals = sp;
  if (!rb.cost->Evaluate(params.data(), residuals, jacobians ? global_j.data() : nullptr))
    return false;
  // IsEvaluationValid: finite outputs (residual_block.cc:118-132)
  for (int r = 0; r < nr; ++r)
    if (!std::isfinite(residuals[r])) return false;
  double sq = 0.0;
  for (int r = 0; r < nr; ++r) sq += residuals[r] * residuals[r];
  if (jacobians) {
    for (int i = 0; i < nb; ++i) {
      if (!jacobians[i]) continue;
      const ParamBlock* pb = rb.blocks[i];
      for (size_t k = 0; k < (size_t)nr * pb->size; ++k)
        if (!std::isfinite(global_j[i][k])) return false;
      if (pb->manifold == SWGN_MANIFOLD_POSE) {
        // jacobians[i] = global (nr x 7) * [I6; 0]  (residual_block.cc:137-160,
        // PoseLocalParameterization::ComputeJacobian)
        for (int r = 0; r < nr; ++r)
          for (int c = 0; c < 6; ++c) {
            double s = 0.0;
            for (int k = 0; k < 7; ++k) s += global_j[i][r * 7 + k] * ((k == c) ? 1.0 : 0.0);
            jacobians[i][r * 6 + c] = s;
          }
      }
    }
  }
  if (rb.cauchy_a <= 0.0 || !apply_loss) {
    *cost = 0.5 * sq;
    return true;
  }
  double rho[3];
  cauchy_loss(rb.cauchy_a, sq, rho);
  *cost = 0.5 * rho[0];
  if (!jacobians && !outputting) return true;
  // Corrector  CERES/internal/ceres/corrector.cc:42-156
  double sqrt_rho1 = std::sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (sq == 0.0 || rho[2] <= 0.0) {
    residual_scaling = sqrt_rho1;
    alpha_sq_norm = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq;
  }
  if (jacobians) {
    for (int i = 0; i < nb; ++i) {
      if (!jacobians[i]) continue;
      int nc = rb.blocks[i]->local;
      double* J = jacobians[i];
      if (alpha_sq_norm == 0.0) {
        for (int k = 0; k < nr * nc; ++k) J[k] *= sqrt_rho1;
      } else {
        for (int c = 0; c < nc; ++c) {
          double rtj = 0.0;
          for (int r = 0; r < nr; ++r) rtj += J[r * nc + c] * residuals[r];
          for (int r = 0; r < nr; ++r)
            J[r * nc + c] = sqrt_rho1 * (J[r * nc + c] - alpha_sq_norm * residuals[r] * rtj);
        }
      }
    }
  }
  if (outputting)
    for (int r = 0; r < nr; ++r) residuals[r] *= residual_scaling;
  return true;
}

// =============================================================================================
// Preprocess: reduced program + ordering + row sort + Jacobian structure
// CERES program.cc:286-411, trust_region_preprocessor.cc:154-252, reorder_program.cc:209-326,
// 425-507, block_jacobian_writer.cc:58-210
// =============================================================================================
bool Solver::Preprocess() {
  // ---- RemoveFixedBlocks
  for (auto& b : blocks) b.index = -1;
  rblocks.clear();
  fixed_cost = 0.0;
  std::vector<double> scratch;
  for (auto& rbp : residual_blocks) {
    ResidualBlock* rb = rbp.get();
    bool all_constant = true;
    for (ParamBlock* pb : rb->blocks)
      if (!pb->constant) {
        all_constant = false;
        pb->index = 1;
      }
    if (!all_constant && rb->is_use) {
      rblocks.push_back(rb);
      continue;
    }
    double cost = 0.0;
    if (!EvaluateResidualBlock(*rb, true, &cost, nullptr, nullptr, &scratch)) {
      error = "Evaluation failed during removal of fixed residual blocks";
      return false;
    }
    fixed_cost += cost;
  }
  pblocks.clear();
  for (auto& b : blocks)
    if (b.index != -1) pblocks.push_back(&b);
  if (rblocks.empty() || pblocks.empty()) {
    error = "empty reduced program";
    return false;
  }
  // ---- ordering: every remaining block must be ordered; group order then block index
  for (ParamBlock* pb : pblocks)
    if (pb->group < 0) {
      error = "parameter block missing from the linear solver ordering";
      return false;
    }
  int min_group_all = std::numeric_limits<int>::max(), min_group = min_group_all;
  for (auto& b : blocks)
    if (b.group >= 0) min_group_all = std::min(min_group_all, b.group);
  for (ParamBlock* pb : pblocks) min_group = std::min(min_group, pb->group);
  if (min_group != min_group_all) {
    // Ceres would silently switch to another linear solver here
    // (trust_region_preprocessor.cc:177-183); the reference guards against it with the
    // InitialBlackFactor on blackvalue2 (RVI/swf/swf_gnss.cpp:645-654).
    error = "first elimination group is empty after removing fixed blocks";
    return false;
  }
  std::stable_sort(pblocks.begin(), pblocks.end(), [](const ParamBlock* a, const ParamBlock* b) {
    if (a->group != b->group) return a->group < b->group;
    return a->graph_index < b->graph_index;
  });
  num_eliminate_blocks = 0;
  for (ParamBlock* pb : pblocks)
    if (pb->group == min_group) ++num_eliminate_blocks;
  // SetParameterOffsetsAndIndex
  num_parameters = num_effective_parameters = 0;
  for (size_t i = 0; i < pblocks.size(); ++i) {
    pblocks[i]->index = (int)i;
    pblocks[i]->state_offset = num_parameters;
    pblocks[i]->delta_offset = num_effective_parameters;
    num_parameters += pblocks[i]->size;
    num_effective_parameters += pblocks[i]->local;
  }
  // ---- independence of the first elimination group (program.cc:413-434)
  for (ResidualBlock* rb : rblocks) {
    int count = 0;
    for (ParamBlock* pb : rb->blocks)
      if (!pb->constant && pb->index < num_eliminate_blocks) ++count;
    if (count > 1) {
      error = "The first elimination group is not an independent set";
      return false;
    }
  }
  // ---- LexicographicallyOrderResidualBlocks (buckets filled back to front)
  {
    const int ne = num_eliminate_blocks;
    std::vector<int> hist(ne + 1, 0), minpos(rblocks.size());
    for (size_t i = 0; i < rblocks.size(); ++i) {
      int pos = ne;
      for (ParamBlock* pb : rblocks[i]->blocks)
        if (!pb->constant) pos = std::min(pos, pb->index);
      minpos[i] = pos;
      hist[pos]++;
    }
    for (int e = 0; e < ne; ++e)
      if (hist[e] == 0) {
        error = "e-block without residuals";
        return false;
      }
    std::vector<int> offsets(ne + 1);
    std::partial_sum(hist.begin(), hist.end(), offsets.begin());
    std::vector<ResidualBlock*> re(rblocks.size(), nullptr);
    for (size_t i = 0; i < rblocks.size(); ++i) re[--offsets[minpos[i]]] = rblocks[i];
    rblocks.swap(re);
  }
  // ---- Jacobian structure: E cells first in memory, then F cells
  {
    jac = BlockSparse();
    jac.cols.resize(pblocks.size());
    for (size_t i = 0; i < pblocks.size(); ++i) {
      jac.cols[i].size = pblocks[i]->local;
      jac.cols[i].position = pblocks[i]->delta_offset;
    }
    jac.num_cols = num_effective_parameters;
    int f_pos = 0;
    for (ResidualBlock* rb : rblocks)
      for (ParamBlock* pb : rb->blocks)
        if (!pb->constant && pb->index < num_eliminate_blocks)
          f_pos += rb->cost->num_residuals * pb->local;
    int e_pos = 0, row_pos = 0;
    jac.rows.resize(rblocks.size());
    residual_layout.resize(rblocks.size());
    for (size_t i = 0; i < rblocks.size(); ++i) {
      ResidualBlock* rb = rblocks[i];
      RowBlock& row = jac.rows[i];
      row.size = rb->cost->num_residuals;
      row.position = row_pos;
      residual_layout[i] = row_pos;
      row_pos += row.size;
      for (ParamBlock* pb : rb->blocks) {
        if (pb->constant) continue;
        Cell c;
        c.block_id = pb->index;
        int sz = row.size * pb->local;
        if (pb->index < num_eliminate_blocks) {
          c.position = e_pos;
          e_pos += sz;
        } else {
          c.position = f_pos;
          f_pos += sz;
        }
        row.cells.push_back(c);
      }
      std::sort(row.cells.begin(), row.cells.end(),
                [](const Cell& a, const Cell& b) { return a.block_id < b.block_id; });
    }
    jac.num_rows = num_residuals = row_pos;
    jac.values.assign(f_pos, 0.0);
  }
  eliminator.Init(num_eliminate_blocks, jac);
  return true;
}

void Solver::StateToUser(const double* x) {
  for (ParamBlock* pb : pblocks)
    std::memcpy(pb->user_state, x + pb->state_offset, sizeof(double) * pb->size);
}

void Solver::Plus(const double* x, const double* delta, double* out) const {
  for (const ParamBlock* pb : pblocks) {
    const double* xs = x + pb->state_offset;
    const double* d = delta + pb->delta_offset;
    double* o = out + pb->state_offset;
    if (pb->manifold == SWGN_MANIFOLD_POSE) {
      pose_plus(xs, d, o);
    } else {
      for (int k = 0; k < pb->size; ++k) o[k] = xs[k] + d[k];
    }
  }
}

// ProgramEvaluator::Evaluate   CERES/internal/ceres/program_evaluator.h:139-286
bool Solver::Evaluate(const double* x, double* cost, double* residuals, double* gradient,
                      bool jacobian) {
  StateToUser(x);  // the restatement evaluates straight out of user memory
  if (residuals) std::fill(residuals, residuals + num_residuals, 0.0);
  if (jacobian) std::fill(jac.values.begin(), jac.values.end(), 0.0);
  if (gradient) std::fill(gradient, gradient + num_effective_parameters, 0.0);
  double total = 0.0;
  std::vector<double> scratch, block_res, jbuf;
  for (size_t i = 0; i < rblocks.size(); ++i) {
    ResidualBlock* rb = rblocks[i];
    const int nr = rb->cost->num_residuals;
    double* block_residuals = nullptr;
    if (residuals) {
      block_residuals = residuals + residual_layout[i];
    } else if (gradient) {
      block_res.resize(nr);
      block_residuals = block_res.data();
    }
    std::vector<double*> block_j;
    if (jacobian || gradient) {
      size_t tot = 0;
      for (ParamBlock* pb : rb->blocks) tot += (size_t)nr * pb->local;
      jbuf.assign(tot, 0.0);
      double* p = jbuf.data();
      for (ParamBlock* pb : rb->blocks) {
        block_j.push_back(pb->constant ? nullptr : p);
        p += (size_t)nr * pb->local;
      }
    }
    double block_cost;
    if (!EvaluateResidualBlock(*rb, true, &block_cost, block_residuals,
                               block_j.empty() ? nullptr : block_j.data(), &scratch))
      return false;
    total += block_cost;
    if (jacobian) {
      for (size_t k = 0; k < rb->blocks.size(); ++k) {
        ParamBlock* pb = rb->blocks[k];
        if (pb->constant) continue;
        for (const Cell& c : jac.rows[i].cells)
          if (c.block_id == pb->index) {
            // a parameter block may appear once per residual block (Ceres CHECKs duplicates)
            std::memcpy(jac.values.data() + c.position, block_j[k],
                        sizeof(double) * nr * pb->local);
            break;
          }
      }
    }
    if (gradient) {
      for (size_t k = 0; k < rb->blocks.size(); ++k) {
        ParamBlock* pb = rb->blocks[k];
        if (pb->constant) continue;
        double* gp = gradient + pb->delta_offset;
        for (int r = 0; r < nr; ++r)
          for (int c = 0; c < pb->local; ++c) gp[c] += block_j[k][r * pb->local + c] * block_residuals[r];
      }
    }
  }
  if (cost) *cost = total;
  return true;
}

// =============================================================================================
// SchurEliminator<Dynamic,Dynamic,Dynamic>   CERES/internal/ceres/schur_eliminator_impl.h
// =============================================================================================
void SchurEliminator::Init(int num_e, const BlockSparse& bs) {
  num_eliminate_blocks = num_e;
  const int num_col_blocks = (int)bs.cols.size();
  const int num_row_blocks = (int)bs.rows.size();
  buffer_size = 1;
  chunks.clear();
  lhs_row_layout.assign(num_col_blocks - num_e, 0);
  lhs_num_rows = 0;
  for (int i = num_e; i < num_col_blocks; ++i) {
    lhs_row_layout[i - num_e] = lhs_num_rows;
    lhs_num_rows += bs.cols[i].size;
  }
  int r = 0;
  while (r < num_row_blocks) {                                   // :118-156
    const int chunk_block_id = bs.rows[r].cells.front().block_id;
    if (chunk_block_id >= num_e) break;
    Chunk chunk;
    chunk.start = r;
    chunk.size = 0;
    int bsize = 0;
    const int e_block_size = bs.cols[chunk_block_id].size;
    while (r + chunk.size < num_row_blocks) {
      const RowBlock& row = bs.rows[r + chunk.size];
      if (row.cells.front().block_id != chunk_block_id) break;
      for (size_t c = 1; c < row.cells.size(); ++c) {
        const Cell& cell = row.cells[c];
        if (chunk.buffer_layout.find(cell.block_id) == chunk.buffer_layout.end()) {
          chunk.buffer_layout[cell.block_id] = bsize;
          bsize += e_block_size * bs.cols[cell.block_id].size;
        }
      }
      buffer_size = std::max(bsize, buffer_size);
      ++chunk.size;
    }
    r += chunk.size;
    chunks.push_back(chunk);
  }
  uneliminated_row_begins = chunks.empty() ? 0 : chunks.back().start + chunks.back().size;
}

// small_blas.h semantics: C op= A^T B etc., plain triple loops
static inline void AtB_add(const double* A, int ar, int ac, const double* B, int br, int bc,
                           double* C, int r0, int c0, int ldc, double sign) {
  (void)br;
  for (int i = 0; i < ac; ++i)
    for (int j = 0; j < bc; ++j) {
      double s = 0.0;
      for (int k = 0; k < ar; ++k) s += A[k * ac + i] * B[k * bc + j];
      C[(size_t)(r0 + i) * ldc + c0 + j] += sign * s;
    }
}

void SchurEliminator::Eliminate(const BlockSparse& A, const double* b, const double* D,
                                double* lhs, double* rhs) const {
  const int n = lhs_num_rows;
  std::fill(lhs, lhs + (size_t)n * n, 0.0);
  std::fill(rhs, rhs + n, 0.0);
  const int num_col_blocks = (int)A.cols.size();
  const double* values = A.values.data();
  if (D) {                                                       // :194-215
    for (int i = num_eliminate_blocks; i < num_col_blocks; ++i) {
      int p = lhs_row_layout[i - num_eliminate_blocks];
      for (int k = 0; k < A.cols[i].size; ++k) {
        double d = D[A.cols[i].position + k];
        lhs[(size_t)(p + k) * n + p + k] += d * d;
      }
    }
  }
  std::vector<double> buffer(buffer_size), tmp(buffer_size);
  for (const Chunk& chunk : chunks) {                            // :230-301
    const int e_id = A.rows[chunk.start].cells.front().block_id;
    const int es = A.cols[e_id].size;
    std::fill(buffer.begin(), buffer.end(), 0.0);
    std::vector<double> ete((size_t)es * es, 0.0), g(es, 0.0);
    if (D)
      for (int k = 0; k < es; ++k) {
        double d = D[A.cols[e_id].position + k];
        ete[(size_t)k * es + k] = d * d;
      }
    // ChunkDiagonalBlockAndGradient :444-507
    for (int j = 0; j < chunk.size; ++j) {
      const RowBlock& row = A.rows[chunk.start + j];
      const int rs = row.size;
      if (row.cells.size() > 1) {                                // EBlockRowOuterProduct :667-716
        for (size_t i = 1; i < row.cells.size(); ++i) {
          const int b1 = row.cells[i].block_id - num_eliminate_blocks;
          const int s1 = A.cols[row.cells[i].block_id].size;
          AtB_add(values + row.cells[i].position, rs, s1, values + row.cells[i].position, rs, s1,
                  lhs, lhs_row_layout[b1], lhs_row_layout[b1], n, 1.0);
          for (size_t k = i + 1; k < row.cells.size(); ++k) {
            const int b2 = row.cells[k].block_id - num_eliminate_blocks;
            const int s2 = A.cols[row.cells[k].block_id].size;
            AtB_add(values + row.cells[i].position, rs, s1, values + row.cells[k].position, rs, s2,
                    lhs, lhs_row_layout[b1], lhs_row_layout[b2], n, 1.0);
          }
        }
      }
      const double* E = values + row.cells.front().position;
      AtB_add(E, rs, es, E, rs, es, ete.data(), 0, 0, es, 1.0);
      for (int i = 0; i < es; ++i) {
        double s = 0.0;
        for (int k = 0; k < rs; ++k) s += E[k * es + i] * b[row.position + k];
        g[i] += s;
      }
      for (size_t c = 1; c < row.cells.size(); ++c) {
        const int f_id = row.cells[c].block_id;
        const int fs = A.cols[f_id].size;
        double* bp = buffer.data() + chunk.buffer_layout.at(f_id);
        AtB_add(E, rs, es, values + row.cells[c].position, rs, fs, bp, 0, 0, fs, 1.0);
      }
    }
    std::vector<double> inv((size_t)es * es);
    invert_psd(ete.data(), es, inv.data());                      // :279-280
    std::vector<double> inv_g(es, 0.0);
    for (int i = 0; i < es; ++i)
      for (int k = 0; k < es; ++k) inv_g[i] += inv[(size_t)i * es + k] * g[k];
    // UpdateRhs :381-422
    for (int j = 0; j < chunk.size; ++j) {
      const RowBlock& row = A.rows[chunk.start + j];
      const int rs = row.size;
      const double* E = values + row.cells.front().position;
      std::vector<double> sj(rs);
      for (int k = 0; k < rs; ++k) {
        double s = 0.0;
        for (int i = 0; i < es; ++i) s += E[k * es + i] * inv_g[i];
        sj[k] = b[row.position + k] - s;
      }
      for (size_t c = 1; c < row.cells.size(); ++c) {
        const int f_id = row.cells[c].block_id;
        const int fs = A.cols[f_id].size;
        const double* F = values + row.cells[c].position;
        double* rp = rhs + lhs_row_layout[f_id - num_eliminate_blocks];
        for (int i = 0; i < fs; ++i) {
          double s = 0.0;
          for (int k = 0; k < rs; ++k) s += F[k * fs + i] * sj[k];
          rp[i] += s;
        }
      }
    }
    // ChunkOuterProduct :514-563
    for (auto it1 = chunk.buffer_layout.begin(); it1 != chunk.buffer_layout.end(); ++it1) {
      const int b1 = it1->first - num_eliminate_blocks;
      const int s1 = A.cols[it1->first].size;
      // tmp (s1 x es) = buffer1^T * inv
      for (int i = 0; i < s1; ++i)
        for (int j = 0; j < es; ++j) {
          double s = 0.0;
          for (int k = 0; k < es; ++k) s += buffer[it1->second + k * s1 + i] * inv[(size_t)k * es + j];
          tmp[i * es + j] = s;
        }
      for (auto it2 = it1; it2 != chunk.buffer_layout.end(); ++it2) {
        const int b2 = it2->first - num_eliminate_blocks;
        const int s2 = A.cols[it2->first].size;
        for (int i = 0; i < s1; ++i)
          for (int j = 0; j < s2; ++j) {
            double s = 0.0;
            for (int k = 0; k < es; ++k) s += tmp[i * es + k] * buffer[it2->second + k * s2 + j];
            lhs[(size_t)(lhs_row_layout[b1] + i) * n + lhs_row_layout[b2] + j] -= s;
          }
      }
    }
  }
  // NoEBlockRowsUpdate :569-661
  for (int r = uneliminated_row_begins; r < (int)A.rows.size(); ++r) {
    const RowBlock& row = A.rows[r];
    const int rs = row.size;
    for (size_t i = 0; i < row.cells.size(); ++i) {
      const int b1 = row.cells[i].block_id - num_eliminate_blocks;
      const int s1 = A.cols[row.cells[i].block_id].size;
      AtB_add(values + row.cells[i].position, rs, s1, values + row.cells[i].position, rs, s1, lhs,
              lhs_row_layout[b1], lhs_row_layout[b1], n, 1.0);
      for (size_t k = i + 1; k < row.cells.size(); ++k) {
        const int b2 = row.cells[k].block_id - num_eliminate_blocks;
        const int s2 = A.cols[row.cells[k].block_id].size;
        AtB_add(values + row.cells[i].position, rs, s1, values + row.cells[k].position, rs, s2, lhs,
                lhs_row_layout[b1], lhs_row_layout[b2], n, 1.0);
      }
      const double* F = values + row.cells[i].position;
      double* rp = rhs + lhs_row_layout[b1];
      for (int c = 0; c < s1; ++c) {
        double s = 0.0;
        for (int k = 0; k < rs; ++k) s += F[k * s1 + c] * b[row.position + k];
        rp[c] += s;
      }
    }
  }
}

void SchurEliminator::BackSubstitute(const BlockSparse& A, const double* b, const double* D,
                                     const double* z, double* y) const {  // :309-375
  const double* values = A.values.data();
  for (const Chunk& chunk : chunks) {
    const int e_id = A.rows[chunk.start].cells.front().block_id;
    const int es = A.cols[e_id].size;
    double* y_ptr = y + A.cols[e_id].position;
    std::vector<double> ete((size_t)es * es, 0.0);
    if (D)
      for (int k = 0; k < es; ++k) {
        double d = D[A.cols[e_id].position + k];
        ete[(size_t)k * es + k] = d * d;
      }
    for (int j = 0; j < chunk.size; ++j) {
      const RowBlock& row = A.rows[chunk.start + j];
      const int rs = row.size;
      std::vector<double> sj(b + row.position, b + row.position + rs);
      for (size_t c = 1; c < row.cells.size(); ++c) {
        const int f_id = row.cells[c].block_id;
        const int fs = A.cols[f_id].size;
        const double* F = values + row.cells[c].position;
        const double* zp = z + lhs_row_layout[f_id - num_eliminate_blocks];
        for (int k = 0; k < rs; ++k) {
          double s = 0.0;
          for (int i = 0; i < fs; ++i) s += F[k * fs + i] * zp[i];
          sj[k] -= s;
        }
      }
      const double* E = values + row.cells.front().position;
      for (int i = 0; i < es; ++i) {
        double s = 0.0;
        for (int k = 0; k < rs; ++k) s += E[k * es + i] * sj[k];
        y_ptr[i] += s;
      }
      AtB_add(E, rs, es, E, rs, es, ete.data(), 0, 0, es, 1.0);
    }
    std::vector<double> inv((size_t)es * es), out(es, 0.0);
    invert_psd(ete.data(), es, inv.data());
    for (int i = 0; i < es; ++i)
      for (int k = 0; k < es; ++k) out[i] += inv[(size_t)i * es + k] * y_ptr[k];
    for (int i = 0; i < es; ++i) y_ptr[i] = out[i];
  }
}

// SchurComplementSolver::SolveImpl + DenseSchurComplementSolver::SolveReducedLinearSystem
// CERES/internal/ceres/schur_complement_solver.cc:126-268 (incl. the HXHADD exports)
bool Solver::LinearSolve(const double* residuals, const double* D, double* y,
                         bool* exported_only) {
  const int n = eliminator.lhs_num_rows;
  const int ncols = jac.num_cols;
  std::vector<double> lhs((size_t)n * n), rhs(n);
  std::fill(y, y + ncols, 0.0);
  eliminator.Eliminate(jac, residuals, D, lhs.data(), rhs.data());
  ++num_linear_solves;
  *exported_only = false;
  if (opt.n_parameter_head > 0 && !opt.is_optimize) {            // :172-188
    exports.hs_row = n;
    exports.lhs_out = lhs;
    exports.rhs_out = rhs;
    exports.have_reduced = true;
    *exported_only = true;
    return true;  // LINEAR_SOLVER_SUCCESS with a zero step
  }
  // keep the last reduced system for inspection by tests even when not exported
  exports.hs_row = n;
  exports.lhs_out = lhs;
  exports.rhs_out = rhs;
  double* reduced = y + ncols - n;
  if (n > 0) {
    std::vector<double> u = lhs;
    if (!llt_upper_inplace(u.data(), n)) return false;          // LINEAR_SOLVER_FAILURE
    if (opt.n_parameter_head > 0 && opt.is_optimize) {           // :253-258  lhs_out2 = matrixL()
      exports.lhs_out2.assign((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) exports.lhs_out2[(size_t)i * n + j] = u[(size_t)j * n + i];
      exports.have_factor = true;
    }
    std::copy(rhs.begin(), rhs.end(), reduced);
    llt_upper_solve(u.data(), n, reduced);
  }
  eliminator.BackSubstitute(jac, residuals, D, reduced, y);
  return true;
}

// =============================================================================================
// TrustRegionMinimizer + DoglegStrategy + TrustRegionStepEvaluator
// CERES trust_region_minimizer.cc:67-134 etc., dogleg_strategy.cc, trust_region_step_evaluator.cc
// =============================================================================================
namespace {
double norm2(const std::vector<double>& v) {
  double s = 0.0;
  for (double x : v) s += x * x;
  return std::sqrt(s);
}
void right_multiply(const BlockSparse& A, const double* x, double* y) {  // y += A x
  for (const RowBlock& row : A.rows)
    for (const Cell& c : row.cells) {
      const int cs = A.cols[c.block_id].size, cp = A.cols[c.block_id].position;
      const double* v = A.values.data() + c.position;
      for (int r = 0; r < row.size; ++r) {
        double s = 0.0;
        for (int k = 0; k < cs; ++k) s += v[r * cs + k] * x[cp + k];
        y[row.position + r] += s;
      }
    }
}
void left_multiply(const BlockSparse& A, const double* x, double* y) {  // y += A^T x
  for (const RowBlock& row : A.rows)
    for (const Cell& c : row.cells) {
      const int cs = A.cols[c.block_id].size, cp = A.cols[c.block_id].position;
      const double* v = A.values.data() + c.position;
      for (int r = 0; r < row.size; ++r)
        for (int k = 0; k < cs; ++k) y[cp + k] += v[r * cs + k] * x[row.position + r];
    }
}
void scale_columns(BlockSparse* A, const double* scale) {  // BlockSparseMatrix::ScaleColumns
  for (const RowBlock& row : A->rows)
    for (const Cell& c : row.cells) {
      const int cs = A->cols[c.block_id].size, cp = A->cols[c.block_id].position;
      double* v = A->values.data() + c.position;
      for (int r = 0; r < row.size; ++r)
        for (int k = 0; k < cs; ++k) v[r * cs + k] *= scale[cp + k];
    }
}
void squared_column_norm(const BlockSparse& A, double* x) {
  std::fill(x, x + A.num_cols, 0.0);
  for (const RowBlock& row : A.rows)
    for (const Cell& c : row.cells) {
      const int cs = A.cols[c.block_id].size, cp = A.cols[c.block_id].position;
      const double* v = A.values.data() + c.position;
      for (int r = 0; r < row.size; ++r)
        for (int k = 0; k < cs; ++k) x[cp + k] += v[r * cs + k] * v[r * cs + k];
    }
}
}  // namespace

bool Solver::Minimize(swgn_summary* summary) {
  const int n = num_effective_parameters;
  const int np = num_parameters;
  const int nr = num_residuals;
  iterations.clear();
  num_linear_solves = 0;
  std::vector<double> x(np), candidate_x(np), residuals(nr), gradient(n), delta(n), step(n),
      model_residuals(nr), neg_grad(n), proj(np), best_x(np);
  for (ParamBlock* pb : pblocks)
    std::memcpy(x.data() + pb->state_offset, pb->user_state, sizeof(double) * pb->size);
  best_x = x;
  const std::vector<double> original_x = x;
  double x_norm = norm2(x);
  double x_cost = std::numeric_limits<double>::max();
  double minimum_cost = x_cost, candidate_cost = 0.0, model_cost_change = 0.0;
  int num_consecutive_invalid = 0;
  int termination = SWGN_NO_CONVERGENCE;
  summary->num_successful_steps = summary->num_unsuccessful_steps = 0;
  summary->fixed_cost = fixed_cost;

  // dogleg state (dogleg_strategy.cc:54-73)
  double radius = opt.initial_trust_region_radius;
  double mu = opt.dogleg_min_mu;
  const double min_mu = opt.dogleg_min_mu, max_mu = 1.0, mu_increase = 10.0;
  bool reuse = false;
  double alpha = 0.0, dogleg_step_norm = 0.0;
  std::vector<double> diagonal(n), dgrad(n), gn(n), lm_diag(n);
  // LevenbergMarquardtStrategy state (levenberg_marquardt_strategy.cc:49-62): radius as above, decrease factor,
  // reuse_diagonal; `diagonal` then holds the clamped SQUARED column norms
  const bool lm = opt.trust_region_strategy == SWGN_LEVENBERG_MARQUARDT;
  double decrease_factor = 2.0;
  bool reuse_diagonal = false;

  // Solver::Options::jacobi_scaling (trust_region_minimizer.cc:183,261-276): the scaling vector is computed from the
  // Jacobian of iteration zero and applied to every later Jacobian; the gradient is taken before the scaling
  std::vector<double> jacobian_scaling(n, 1.0);
  bool have_scaling = false;
  IterationRecord it = {};
  auto eval_grad_jac = [&]() -> bool {                           // EvaluateGradientAndJacobian
    if (!Evaluate(x.data(), &x_cost, residuals.data(), gradient.data(), true)) return false;
    if (opt.jacobi_scaling) {
      if (!have_scaling) {
        squared_column_norm(jac, jacobian_scaling.data());
        for (int i = 0; i < n; ++i) jacobian_scaling[i] = 1.0 / (1.0 + std::sqrt(jacobian_scaling[i]));
        have_scaling = true;
      }
      scale_columns(&jac, jacobian_scaling.data());
    }
    it.cost = x_cost + fixed_cost;
    for (int i = 0; i < n; ++i) neg_grad[i] = -gradient[i];
    Plus(x.data(), neg_grad.data(), proj.data());
    double mx = 0.0;
    for (int i = 0; i < np; ++i) mx = std::max(mx, std::fabs(x[i] - proj[i]));
    it.gradient_max_norm = mx;
    return true;
  };

  // IterationZero :195-230
  it.step_is_valid = 0;
  it.step_is_successful = 0;
  if (!eval_grad_jac()) {
    error = "Residual and Jacobian evaluation failed.";
    summary->termination_type = SWGN_FAILURE;
    return false;
  }
  summary->initial_cost = x_cost + fixed_cost;
  it.step_is_valid = 1;
  it.step_is_successful = 1;
  int iteration = 0;

  // step evaluator (max_consecutive_nonmonotonic_steps = 0)
  double se_min = x_cost, se_cur = x_cost, se_ref = x_cost, se_cand = x_cost;
  double se_acc_ref = 0.0, se_acc_cand = 0.0;
  int se_nonmono = 0;

  bool export_return = false;
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue :303-365
    if (it.step_is_successful) {
      ++summary->num_successful_steps;
      if (x_cost < minimum_cost) {
        minimum_cost = x_cost;
        best_x = x;
      }
    } else {
      ++summary->num_unsuccessful_steps;
    }
    it.radius = radius;
    iterations.push_back(it);
    if (iteration >= opt.max_num_iterations) {
      termination = SWGN_NO_CONVERGENCE;
      break;
    }
    if (it.step_is_successful && it.gradient_max_norm <= opt.gradient_tolerance) {
      termination = SWGN_CONVERGENCE;
      break;
    }
    if (radius <= opt.min_trust_region_radius) {
      termination = SWGN_CONVERGENCE;
      break;
    }
    const double prev_gmax = it.gradient_max_norm;
    it = IterationRecord();
    ++iteration;

    // ---- ComputeTrustRegionStep -> DoglegStrategy::ComputeStep (dogleg_strategy.cc:79-165)
    bool linear_failure = false;
    if (lm) {
      // LevenbergMarquardtStrategy::ComputeStep :67-149
      if (!reuse_diagonal) {
        squared_column_norm(jac, diagonal.data());
        for (int i = 0; i < n; ++i) diagonal[i] = std::min(std::max(diagonal[i], opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
      for (int i = 0; i < n; ++i) lm_diag[i] = std::sqrt(diagonal[i] / radius);
      bool exported_only = false;
      bool ok = LinearSolve(residuals.data(), lm_diag.data(), step.data(), &exported_only);
      if (exported_only) export_return = true;
      if (ok)
        for (int i = 0; i < n; ++i)
          if (!std::isfinite(step[i])) ok = false;
      linear_failure = !ok;
      if (ok)
        for (int i = 0; i < n; ++i) step[i] = -step[i];
      reuse_diagonal = true;
    } else if (!reuse) {
      reuse = true;
      squared_column_norm(jac, diagonal.data());
      for (int i = 0; i < n; ++i)
        diagonal[i] = std::sqrt(
            std::min(std::max(diagonal[i], opt.min_lm_diagonal), opt.max_lm_diagonal));
      std::fill(dgrad.begin(), dgrad.end(), 0.0);                 // ComputeGradient :174-179
      left_multiply(jac, residuals.data(), dgrad.data());
      for (int i = 0; i < n; ++i) dgrad[i] /= diagonal[i];
      {                                                           // ComputeCauchyPoint :183-192
        std::vector<double> Jg(nr, 0.0), sg(n);
        for (int i = 0; i < n; ++i) sg[i] = dgrad[i] / diagonal[i];
        right_multiply(jac, sg.data(), Jg.data());
        double gs = 0.0, js = 0.0;
        for (double v : dgrad) gs += v * v;
        for (double v : Jg) js += v * v;
        alpha = gs / js;
      }
      // ComputeGaussNewtonStep :515-610
      linear_failure = true;
      while (mu < max_mu) {
        for (int i = 0; i < n; ++i) lm_diag[i] = diagonal[i] * std::sqrt(mu);
        bool exported_only = false;
        bool ok = LinearSolve(residuals.data(), lm_diag.data(), gn.data(), &exported_only);
        if (exported_only) export_return = true;
        bool valid = ok;
        if (ok)
          for (int i = 0; i < n; ++i)
            if (!std::isfinite(gn[i])) valid = false;
        if (!valid) {
          mu *= mu_increase;
          continue;
        }
        linear_failure = false;
        break;
      }
      if (!linear_failure)
        for (int i = 0; i < n; ++i) gn[i] *= -diagonal[i];
    }
    it.step_is_valid = 0;
    if (!linear_failure && lm) {
      // model cost change  trust_region_minimizer.cc:414-431 (the step is already in the unscaled space)
      std::fill(model_residuals.begin(), model_residuals.end(), 0.0);
      right_multiply(jac, step.data(), model_residuals.data());
      double mc = 0.0;
      for (int i = 0; i < nr; ++i) mc += model_residuals[i] * (residuals[i] + model_residuals[i] / 2.0);
      model_cost_change = -mc;
      it.step_is_valid = model_cost_change > 0.0;
      if (it.step_is_valid) {
        for (int i = 0; i < n; ++i) delta[i] = step[i] * jacobian_scaling[i];  // undo the column scaling :437
        num_consecutive_invalid = 0;
      }
    } else if (!linear_failure) {
      // ComputeTraditionalDoglegStep :199-253
      double gnorm = norm2(dgrad), gnn = norm2(gn);
      if (gnn <= radius) {
        step = gn;
        dogleg_step_norm = gnn;
      } else if (gnorm * alpha >= radius) {
        for (int i = 0; i < n; ++i) step[i] = -(radius / gnorm) * dgrad[i];
        dogleg_step_norm = radius;
      } else {
        double gdot = 0.0;
        for (int i = 0; i < n; ++i) gdot += dgrad[i] * gn[i];
        const double b_dot_a = -alpha * gdot;
        const double a_sq = std::pow(alpha * gnorm, 2.0);
        const double bma_sq = a_sq - 2 * b_dot_a + std::pow(gnn, 2);
        const double c = b_dot_a - a_sq;
        const double d = std::sqrt(c * c + bma_sq * (std::pow(radius, 2.0) - a_sq));
        double beta = (c <= 0) ? (d - c) / bma_sq : (radius * radius - a_sq) / (d + c);
        for (int i = 0; i < n; ++i) step[i] = (-alpha * (1.0 - beta)) * dgrad[i] + beta * gn[i];
        dogleg_step_norm = norm2(step);
      }
      for (int i = 0; i < n; ++i) step[i] /= diagonal[i];
      // model cost change  trust_region_minimizer.cc:414-431
      std::fill(model_residuals.begin(), model_residuals.end(), 0.0);
      right_multiply(jac, step.data(), model_residuals.data());
      double mc = 0.0;
      for (int i = 0; i < nr; ++i) mc += model_residuals[i] * (residuals[i] + model_residuals[i] / 2.0);
      model_cost_change = -mc;
      it.step_is_valid = model_cost_change > 0.0;
      if (it.step_is_valid) {
        for (int i = 0; i < n; ++i) delta[i] = step[i] * jacobian_scaling[i];  // (all ones unless jacobi_scaling)
        num_consecutive_invalid = 0;
      }
    }
    if (!it.step_is_valid) {                                     // HandleInvalidStep :453-486
      if (++num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) {
        termination = SWGN_FAILURE;
        break;
      }
      if (lm) {  // StepIsInvalid = StepRejected(0)  levenberg_marquardt_strategy.h:61-67
        radius = radius / decrease_factor;
        decrease_factor *= 2.0;
        reuse_diagonal = true;
      } else {
        mu *= mu_increase;                                       // StepIsInvalid
        reuse = false;
      }
      it.cost = x_cost + fixed_cost;
      it.cost_change = 0.0;
      it.gradient_max_norm = prev_gmax;
      it.step_norm = 0.0;
      it.relative_decrease = 0.0;
      it.step_is_successful = 0;
      continue;
    }
    // ComputeCandidatePointAndEvaluateCost :761-779
    Plus(x.data(), delta.data(), candidate_x.data());
    if (!Evaluate(candidate_x.data(), &candidate_cost, nullptr, nullptr, false))
      candidate_cost = std::numeric_limits<double>::max();
    // ParameterToleranceReached :706-727
    {
      double s = 0.0;
      for (int i = 0; i < np; ++i) s += (x[i] - candidate_x[i]) * (x[i] - candidate_x[i]);
      it.step_norm = std::sqrt(s);
      if (it.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
        termination = SWGN_CONVERGENCE;
        break;
      }
    }
    // FunctionToleranceReached :730-748
    it.cost_change = x_cost - candidate_cost;
    if (std::fabs(it.cost_change) <= opt.function_tolerance * x_cost) {
      termination = SWGN_CONVERGENCE;
      break;
    }
    // IsStepSuccessful / StepQuality  trust_region_step_evaluator.cc:52-68
    {
      double q;
      if (candidate_cost >= std::numeric_limits<double>::max()) {
        q = std::numeric_limits<double>::lowest();
      } else {
        double rel = (se_cur - candidate_cost) / model_cost_change;
        double hist = (se_ref - candidate_cost) / (se_acc_ref + model_cost_change);
        q = std::max(rel, hist);
      }
      it.relative_decrease = q;
    }
    if (it.relative_decrease > opt.min_relative_decrease) {
      // HandleSuccessfulStep :812-826
      x = candidate_x;
      x_norm = norm2(x);
      if (!eval_grad_jac()) {
        termination = SWGN_FAILURE;
        break;
      }
      it.step_is_successful = 1;
      if (lm) {  // LevenbergMarquardtStrategy::StepAccepted :151-158
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
        radius = std::min(opt.max_trust_region_radius, radius);
        decrease_factor = 2.0;
        reuse_diagonal = false;
      } else {
        // DoglegStrategy::StepAccepted :612-628
        if (it.relative_decrease < 0.25) radius *= 0.5;
        if (it.relative_decrease > 0.75) radius = std::max(radius, 3.0 * dogleg_step_norm);
        mu = std::max(min_mu, 2.0 * mu / mu_increase);
        reuse = false;
      }
      // TrustRegionStepEvaluator::StepAccepted :70-112
      se_cur = candidate_cost;
      se_acc_cand += model_cost_change;
      se_acc_ref += model_cost_change;
      if (se_cur < se_min) {
        se_min = se_cur;
        se_nonmono = 0;
        se_cand = se_cur;
        se_acc_cand = 0.0;
      } else {
        ++se_nonmono;
        if (se_cur > se_cand) {
          se_cand = se_cur;
          se_acc_cand = 0.0;
        }
      }
      if (se_nonmono == 0) {
        se_ref = se_cand;
        se_acc_ref = se_acc_cand;
      }
    } else {
      it.step_is_successful = 0;
      it.cost = candidate_cost + fixed_cost;
      it.gradient_max_norm = prev_gmax;
      if (lm) {  // LevenbergMarquardtStrategy::StepRejected :160-164
        radius = radius / decrease_factor;
        decrease_factor *= 2.0;
        reuse_diagonal = true;
      } else {
        radius *= 0.5;                                           // StepRejected :630-633
        reuse = true;
      }
    }
  }
  (void)export_return;
  // solver.cc: the minimiser's parameter vector (lowest-cost x seen) goes back to the user
  // (solver.cc:444-447: an unusable solution restores the original parameters)
  StateToUser(termination == SWGN_FAILURE ? original_x.data() : best_x.data());
  summary->termination_type = termination;
  summary->num_iterations = iteration;
  summary->num_linear_solves = num_linear_solves;
  // SetSummaryFinalCost (solver.cc:317-328): minimum over the recorded iteration costs
  summary->final_cost = summary->initial_cost;
  for (const IterationRecord& r : iterations) summary->final_cost = std::min(summary->final_cost, r.cost);
  summary->n_e = 0;
  for (int i = 0; i < num_eliminate_blocks; ++i) summary->n_e += pblocks[i]->local;
  summary->n_f = eliminator.lhs_num_rows;
  summary->n_residuals = nr;
  return termination != SWGN_FAILURE;
}

}  // namespace oracle
