// Host-side preprocessing of one window (see plan.h).  Integer/structural work only; all
// arithmetic of the solve happens on the device.
#include "plan.h"

#include <algorithm>
#include <climits>
#include <cstring>
#include <deque>
#include <map>
#include <numeric>

namespace swgn {
namespace {

struct Factor {
  int kind, idx;             // storage position inside its kind table
  std::vector<int> blocks;   // graph block ids in the factor's own parameter order
  int nres;
  bool is_use = true;
  bool active = false;       // part of the reduced program
  int res_off = -1;          // first residual row (active only)
  std::vector<int> jac_off;  // per parameter: offset of its cell in W_JAC, -1 constant/inactive
};

inline int local_size(int size, int manifold) { return manifold == SWGN_MANIFOLD_POSE ? 6 : size; }
inline int64_t align2(int64_t x) { return (x + 1) & ~int64_t(1); }

}  // namespace

void constant_sizes(const swgn_graph* g, int64_t sizes[NUM_CARR]) {
  sizes[C_GLOBALS] = 12;
  sizes[C_PROJ_UV] = 2 * (int64_t)g->n_proj;
  sizes[C_IMU] = (int64_t)IMU_DEV_STRIDE * g->n_imu;
  sizes[C_GNSS] = (int64_t)GNSS_DEV_STRIDE * g->n_gnss;
  int64_t nj = 0, nr = 0, nx = 0;
  for (int i = 0; i < g->n_prior; ++i) {
    nj += (int64_t)g->prior_n[i] * g->prior_n[i];
    nr += g->prior_n[i];
    for (int k = g->prior_blk_begin[i]; k < g->prior_blk_begin[i + 1]; ++k) nx += g->block_size[g->prior_blocks[k]];
  }
  sizes[C_PRIOR_J] = nj;
  sizes[C_PRIOR_R0] = nr;
  sizes[C_PRIOR_X0] = nx;
  sizes[C_UNIT] = g->n_unit;
  int64_t nc = 0;
  for (int i = 0; i < g->n_chain; ++i) {
    const int m = g->chain_frame_begin[i + 1] - g->chain_frame_begin[i];
    const int k = g->chain_blk_begin[i + 1] - g->chain_blk_begin[i] - 4;
    nc += ChainLayout(m, std::max(k, 0)).c_size;
  }
  sizes[C_CHAIN] = nc;
}

static void pack_imu_record(const double* src, double* d) {
  for (int k = 0; k < 24; ++k) d[k] = src[k];
  const double* J = src + SWGN_IMU_JACOBIAN;
  static const int blk[5][2] = {{0, 9}, {0, 12}, {3, 12}, {6, 9}, {6, 12}};  // dp_dba dp_dbg dq_dbg dv_dba dv_dbg
  for (int bI = 0; bI < 5; ++bI)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) d[IMU_DEV_BLOCKS + bI * 9 + r * 3 + c] = J[(blk[bI][0] + r) * 15 + blk[bI][1] + c];
  d[69] = 0.0;
  for (int k = 0; k < 225; ++k) d[IMU_DEV_SQRT + k] = src[SWGN_IMU_SQRT_INFO + k];
  d[295] = 0.0;
}

// Factor constants in the device layout (device_types.h CArr); dst[a] has constant_sizes()[a] room.
void pack_constants(const swgn_graph* g, double* const dst[NUM_CARR]) {
  {
    double* c = dst[C_GLOBALS];
    for (int k = 0; k < 3; ++k) {
      c[k] = g->Pbg[k];
      c[3 + k] = g->gravity[k];
    }
    for (int k = 0; k < 4; ++k) c[6 + k] = g->proj_sqrt_info[k];
    c[10] = g->proj_cauchy_a;
    c[11] = 0.0;
  }
  if (g->n_proj) std::memcpy(dst[C_PROJ_UV], g->proj_uv, sizeof(double) * 2 * (size_t)g->n_proj);
  for (int i = 0; i < g->n_imu; ++i)
    pack_imu_record(g->imu_data + (size_t)SWGN_IMU_STRIDE * i, dst[C_IMU] + (size_t)IMU_DEV_STRIDE * i);
  for (int i = 0; i < g->n_gnss; ++i) {
    const double* src = g->gnss_data + (size_t)SWGN_GNSS_STRIDE * i;
    double* d = dst[C_GNSS] + (size_t)GNSS_DEV_STRIDE * i;
    for (int k = 0; k < 9; ++k) d[k] = src[k];
    d[9] = src[SWGN_GNSS_MEAS];
    d[10] = src[SWGN_GNSS_LAM];
    d[11] = src[SWGN_GNSS_WEIGHT];
  }
  int64_t oJ = 0, oR = 0, oX = 0;
  for (int i = 0; i < g->n_prior; ++i) {
    const int n = g->prior_n[i];
    std::memcpy(dst[C_PRIOR_J] + oJ, g->prior_J + g->prior_J_begin[i], sizeof(double) * (size_t)n * n);
    std::memcpy(dst[C_PRIOR_R0] + oR, g->prior_r0 + g->prior_r_begin[i], sizeof(double) * n);
    const double* x0 = g->prior_x0 + g->prior_x0_begin[i];
    int x0o = 0;
    for (int k = g->prior_blk_begin[i]; k < g->prior_blk_begin[i + 1]; ++k) {
      const int bs = g->block_size[g->prior_blocks[k]];
      for (int q = 0; q < bs; ++q) dst[C_PRIOR_X0][oX + q] = x0[x0o + q];
      oX += bs;
      x0o += bs;
    }
    oJ += (int64_t)n * n;
    oR += n;
  }
  for (int i = 0; i < g->n_unit; ++i) dst[C_UNIT][i] = g->unit_istd[i];
  {  // IMUGNSSFactor chains
    int64_t oc = 0;
    size_t frame_n_off = 0, chain_n_off = 0, imu_off = 0;
    for (int i = 0; i < g->n_chain; ++i) {
      const int f0 = g->chain_frame_begin[i], m = g->chain_frame_begin[i + 1] - f0;
      const int k = std::max(0, g->chain_blk_begin[i + 1] - g->chain_blk_begin[i] - 4);
      const ChainLayout L(m, k);
      double* c = dst[C_CHAIN] + oc;
      std::fill(c, c + L.c_size, 0.0);
      for (int q = 0; q <= m; ++q)
        pack_imu_record(g->chain_imu_data + imu_off + (size_t)SWGN_IMU_STRIDE * q, c + L.c_imu + IMU_DEV_STRIDE * q);
      if (m > 0)
        std::memcpy(c + L.c_frame, g->chain_frame_data + (size_t)SWGN_CHAIN_FRAME_STRIDE * f0,
                    sizeof(double) * (size_t)SWGN_CHAIN_FRAME_STRIDE * (size_t)m);
      for (int q = 0; q < m; ++q)
        std::memcpy(c + L.c_frameN + q * L.pn_stride, g->chain_frame_N + frame_n_off + (size_t)15 * k * q, sizeof(double) * 15 * k);
      std::memcpy(c + L.c_NN, g->chain_N + chain_n_off, sizeof(double) * (size_t)k * k);
      std::memcpy(c + L.c_Nrhs, g->chain_N + chain_n_off + (size_t)k * k, sizeof(double) * k);
      oc += L.c_size;
      frame_n_off += (size_t)m * 15 * k;
      chain_n_off += (size_t)k * k + k;
      imu_off += (size_t)(m + 1) * SWGN_IMU_STRIDE;
    }
  }
}

swgn_status build_plan(const swgn_graph* g, int n_parameter_head, WindowPlan* P, std::string* err) {
  auto fail = [&](swgn_status st, const char* m) {
    if (err) *err = m;
    return st;
  };
  if (!g || g->n_blocks <= 0 || !g->block_size || !g->state) return fail(SWGN_ERR_INVALID, "empty graph");
  const int nb = g->n_blocks;
  for (int i = 0; i < nb; ++i) {
    if (g->block_size[i] <= 0) return fail(SWGN_ERR_INVALID, "non-positive block size");
    if (g->block_manifold[i] == SWGN_MANIFOLD_POSE && g->block_size[i] != 7)
      return fail(SWGN_ERR_INVALID, "pose manifold on a block that is not 7-dim");
    if (g->block_offset[i] < 0 || g->block_offset[i] + g->block_size[i] > g->n_state)
      return fail(SWGN_ERR_INVALID, "block outside the state vector");
  }

  // ---- factors, kind-major storage
  std::vector<Factor> fac;
  int kind_begin[NUM_KINDS + 1] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto add_factor = [&](int kind, int idx, const int32_t* blocks, int n, int nres) -> bool {
    for (int k = 0; k < n; ++k)
      if (blocks[k] < 0 || blocks[k] >= nb) return false;
    fac.emplace_back();
    Factor& f = fac.back();
    f.kind = kind;
    f.idx = idx;
    f.nres = nres;
    f.blocks.assign(blocks, blocks + n);
    f.jac_off.assign(n, -1);
    return true;
  };
  fac.reserve((size_t)std::max(0, g->n_proj) + std::max(0, g->n_imu) + std::max(0, g->n_gnss) + std::max(0, g->n_prior) +
              std::max(0, g->n_unit) + std::max(0, g->n_chain) + std::max(0, g->n_host));
  static const int kGnssArity[6] = {2, 3, 3, 2, 3, 2};
  static const int kGnssSizes[6][3] = {{7, 1, 0}, {7, 1, 1}, {7, 1, 1}, {7, 1, 0}, {9, 1, 7}, {1, 1, 0}};
  for (int i = 0; i < g->n_proj; ++i)
    if (!add_factor(K_PROJ, i, g->proj_blocks + 3 * i, 3, 2)) return fail(SWGN_ERR_INVALID, "bad proj block");
  kind_begin[1] = (int)fac.size();
  for (int i = 0; i < g->n_imu; ++i)
    if (!add_factor(K_IMU, i, g->imu_blocks + 4 * i, 4, 15)) return fail(SWGN_ERR_INVALID, "bad imu block");
  kind_begin[2] = (int)fac.size();
  for (int i = 0; i < g->n_gnss; ++i) {
    int k = g->gnss_kind[i];
    if (k < 0 || k > 5) return fail(SWGN_ERR_INVALID, "bad gnss kind");
    if (!add_factor(K_GNSS, i, g->gnss_blocks + 3 * i, kGnssArity[k], 1)) return fail(SWGN_ERR_INVALID, "bad gnss block");
    for (int p = 0; p < kGnssArity[k]; ++p)
      if (g->block_size[fac.back().blocks[p]] != kGnssSizes[k][p]) return fail(SWGN_ERR_INVALID, "gnss block size mismatch");
  }
  kind_begin[3] = (int)fac.size();
  for (int i = 0; i < g->n_prior; ++i) {
    int b0 = g->prior_blk_begin[i], b1 = g->prior_blk_begin[i + 1];
    if (!add_factor(K_PRIOR, i, g->prior_blocks + b0, b1 - b0, g->prior_n[i])) return fail(SWGN_ERR_INVALID, "bad prior block");
  }
  kind_begin[4] = (int)fac.size();
  for (int i = 0; i < g->n_unit; ++i) {
    if (!add_factor(K_UNIT, i, g->unit_block + i, 1, 1)) return fail(SWGN_ERR_INVALID, "bad unit block");
    if (g->block_size[g->unit_block[i]] != 1) return fail(SWGN_ERR_INVALID, "unit factor on a non-scalar block");
  }
  kind_begin[5] = (int)fac.size();
  static_assert((int)SWGN_CHAIN_FRAME_STRIDE == (int)CHAIN_FRAME_STRIDE, "frame record stride");
  int n_chain_frames = 0, max_chain_k = 0;
  for (int i = 0; i < g->n_chain; ++i) {
    const int b0 = g->chain_blk_begin[i], b1 = g->chain_blk_begin[i + 1];
    const int m = g->chain_frame_begin[i + 1] - g->chain_frame_begin[i], k = b1 - b0 - 4;
    if (k < 0 || m < 1) return fail(SWGN_ERR_INVALID, "chain factor needs 4 + k blocks and at least one hidden frame");
    if (k > MAX_CHAIN_K) return fail(SWGN_ERR_UNSUPPORTED, "chain factor with more than 48 phase biases");
    if (!add_factor(K_CHAIN, i, g->chain_blocks + b0, b1 - b0, 30 + k)) return fail(SWGN_ERR_INVALID, "bad chain block");
    static const int csz[4] = {7, 9, 7, 9};
    for (int p = 0; p < 4 + k; ++p)
      if (g->block_size[fac.back().blocks[p]] != (p < 4 ? csz[p] : 1)) return fail(SWGN_ERR_INVALID, "chain block size mismatch");
    for (int p = 0; p < 4; p += 2)
      if (g->block_manifold[fac.back().blocks[p]] != SWGN_MANIFOLD_POSE) return fail(SWGN_ERR_INVALID, "chain pose block without the pose manifold");
    n_chain_frames += m;
    max_chain_k = std::max(max_chain_k, k);
  }
  kind_begin[6] = (int)fac.size();
  for (int i = 0; i < g->n_host; ++i) {
    if (!g->host_nres || !g->host_blk_begin || !g->host_blocks || !g->host_eval) return fail(SWGN_ERR_INVALID, "host-evaluated factors without tables / callback");
    const int b0 = g->host_blk_begin[i], b1 = g->host_blk_begin[i + 1];
    if (b1 <= b0 || g->host_nres[i] <= 0) return fail(SWGN_ERR_INVALID, "bad host factor");
    if (!add_factor(K_HOST, i, g->host_blocks + b0, b1 - b0, g->host_nres[i])) return fail(SWGN_ERR_INVALID, "bad host factor block");
  }
  kind_begin[7] = (int)fac.size();
  for (const Factor& f : fac) {
    static const int proj_sz[3] = {7, 7, 3}, imu_sz[4] = {7, 9, 7, 9};
    if (f.kind == K_PROJ)
      for (int p = 0; p < 3; ++p)
        if (g->block_size[f.blocks[p]] != proj_sz[p]) return fail(SWGN_ERR_INVALID, "proj block size mismatch");
    if (f.kind == K_IMU)
      for (int p = 0; p < 4; ++p)
        if (g->block_size[f.blocks[p]] != imu_sz[p]) return fail(SWGN_ERR_INVALID, "imu block size mismatch");
    for (size_t a = 0; a < f.blocks.size(); ++a)
      for (size_t b = a + 1; b < f.blocks.size(); ++b)
        if (f.blocks[a] == f.blocks[b]) return fail(SWGN_ERR_INVALID, "duplicate parameter block in a residual block");
  }
  if (g->is_use)
    for (size_t i = 0; i < fac.size(); ++i) fac[i].is_use = g->is_use[i] != 0;

  // ---- program order
  std::vector<int> program;
  if (g->order && g->n_order > 0) {
    std::vector<char> seen(fac.size(), 0);
    for (int k = 0; k < g->n_order; ++k) {
      uint32_t kind = g->order[k] >> 28, idx = g->order[k] & 0x0fffffffu;
      if (kind > 6 || (int)idx >= kind_begin[kind + 1] - kind_begin[kind]) return fail(SWGN_ERR_INVALID, "bad program order entry");
      int f = kind_begin[kind] + (int)idx;
      if (seen[f]) return fail(SWGN_ERR_INVALID, "residual block listed twice in the program order");
      seen[f] = 1;
      program.push_back(f);
    }
    if (program.size() != fac.size()) return fail(SWGN_ERR_INVALID, "program order does not list every residual block");
  } else {
    program.resize(fac.size());
    std::iota(program.begin(), program.end(), 0);
  }
  std::vector<int> program_index(fac.size());
  for (size_t k = 0; k < program.size(); ++k) program_index[program[k]] = (int)k;

  // ---- RemoveFixedBlocks (CERES program.cc:304-411 with the is_use mask, M3)
  std::vector<char> referenced(nb, 0);
  int n_active = 0;
  for (int fi : program) {
    Factor& f = fac[fi];
    bool all_const = true;
    for (int b : f.blocks)
      if (!g->block_const[b]) {
        all_const = false;
        referenced[b] = 1;
      }
    f.active = !all_const && f.is_use;
    n_active += f.active;
  }
  std::vector<int> cols;
  for (int b = 0; b < nb; ++b)
    if (referenced[b]) cols.push_back(b);
  if (cols.empty() || n_active == 0) return fail(SWGN_ERR_INVALID, "nothing to optimise: empty reduced program");

  // ---- ordering (reorder_program.cc:209-245): group ascending, then block index
  int min_group_all = INT32_MAX, min_group = INT32_MAX;
  for (int b = 0; b < nb; ++b)
    if (g->block_group[b] >= 0) min_group_all = std::min(min_group_all, g->block_group[b]);
  for (int b : cols) {
    if (g->block_group[b] < 0) return fail(SWGN_ERR_ORDERING, "a variable parameter block is missing from the ordering");
    min_group = std::min(min_group, g->block_group[b]);
  }
  if (min_group != min_group_all)
    return fail(SWGN_ERR_UNSUPPORTED,
                "first elimination group is empty after removing fixed blocks (Ceres would switch linear solver)");
  std::stable_sort(cols.begin(), cols.end(), [&](int a, int b) {
    if (g->block_group[a] != g->block_group[b]) return g->block_group[a] < g->block_group[b];
    return a < b;
  });
  const int n_cols = (int)cols.size();
  int n_ecols = 0;
  for (int b : cols) n_ecols += (g->block_group[b] == min_group);
  std::vector<int> col_of_block(nb, -1), col_pos(n_cols), col_size(n_cols);
  int n_t = 0, n_e = 0;
  for (int c = 0; c < n_cols; ++c) {
    col_of_block[cols[c]] = c;
    col_size[c] = local_size(g->block_size[cols[c]], g->block_manifold[cols[c]]);
    col_pos[c] = n_t;
    n_t += col_size[c];
    if (c < n_ecols) n_e += col_size[c];
  }
  const int n_f = n_t - n_e;

  // ---- independence of the e-blocks (program.cc:413-434) + lexicographic row order
  // (reorder_program.cc:247-326: buckets filled back to front)
  std::vector<int> active_list, minpos;
  std::vector<int> hist(n_ecols + 1, 0);
  for (int fi : program) {
    Factor& f = fac[fi];
    if (!f.active) continue;
    int cnt = 0, pos = n_ecols;
    for (int b : f.blocks) {
      int c = col_of_block[b];
      if (g->block_const[b] || c < 0) continue;
      if (c < n_ecols) ++cnt;
      pos = std::min(pos, c);
    }
    if (cnt > 1) return fail(SWGN_ERR_ORDERING, "The first elimination group is not an independent set");
    active_list.push_back(fi);
    minpos.push_back(pos);
    hist[pos]++;
  }
  for (int e = 0; e < n_ecols; ++e)
    if (hist[e] == 0) return fail(SWGN_ERR_INVALID, "an eliminated parameter block has no residual block");
  std::vector<int> offsets(n_ecols + 1);
  std::partial_sum(hist.begin(), hist.end(), offsets.begin());
  std::vector<int> rows(active_list.size(), -1);
  for (size_t i = 0; i < active_list.size(); ++i) rows[--offsets[minpos[i]]] = active_list[i];
  const int n_rows = (int)rows.size();

  // ---- cells, Jacobian value offsets, chunks, slots
  std::vector<int32_t>*I = P->iarr;
  for (int a = 0; a < NUM_IARR; ++a) I[a].clear();
  P->n_mma = 0;
  int n_res = 0, n_jac = 0;
  int64_t schur_doubles = 0;
  I[I_ROW_CELL].push_back(0);
  {
    size_t n_cells_max = 0, n_res_max = 0;
    for (int r = 0; r < n_rows; ++r) {
      n_cells_max += fac[rows[r]].blocks.size();
      n_res_max += (size_t)fac[rows[r]].nres;
    }
    for (int a : {I_ROW_RES, I_ROW_NRES, I_ROW_FACTOR, I_ROW_CELL}) I[a].reserve((size_t)n_rows + 1);
    for (int a : {I_CELL_COL, I_CELL_VAL, I_CELL_SLOT, I_CELL_FIRST}) I[a].reserve(n_cells_max);
    I[I_RS_ROW].reserve(n_res_max);
  }
  std::vector<std::pair<int, int>> cells;  // (col, param slot) of the current row
  for (int r = 0; r < n_rows; ++r) {
    Factor& f = fac[rows[r]];
    f.res_off = n_res;
    I[I_ROW_RES].push_back(n_res);
    I[I_ROW_NRES].push_back(f.nres);
    I[I_ROW_FACTOR].push_back(program_index[rows[r]]);
    for (int k = 0; k < f.nres; ++k) I[I_RS_ROW].push_back(r);
    cells.clear();
    for (size_t p = 0; p < f.blocks.size(); ++p) {
      int b = f.blocks[p];
      if (g->block_const[b]) continue;
      cells.push_back({col_of_block[b], (int)p});
    }
    std::sort(cells.begin(), cells.end());
    for (auto& c : cells) {
      I[I_CELL_COL].push_back(c.first);
      I[I_CELL_VAL].push_back(n_jac);
      I[I_CELL_SLOT].push_back(-1);
      I[I_CELL_FIRST].push_back(0);
      f.jac_off[c.second] = n_jac;
      n_jac += f.nres * col_size[c.first];
      schur_doubles += (int64_t)f.nres * col_size[c.first];
    }
    schur_doubles += f.nres;
    I[I_ROW_CELL].push_back((int32_t)I[I_CELL_COL].size());
    n_res += f.nres;
  }
  const int n_cells = (int)I[I_CELL_COL].size();
  // chunks (schur_eliminator_impl.h:118-156).  Per chunk the kernels keep
  //   L  = chol(E'E + D_e^2)           (W_EFAC, es x es)
  //   Wf = L^-1 E'F_f  per f-block     (W_EBUF, es x fs row-major, one "slot" per f-block)
  //   wg = L^-1 E'b                    (W_EBUF, es)
  // so that S_pq -= Wp' Wq, rhs_p -= Wp' wg and y_e = L^-T (wg - sum_f Wf z_f).
  int n_efac = 0, n_ebuf = 0, max_wbuf = 0;
  {
    int r = 0;
    I[I_CHUNK_ROW].push_back(0);
    I[I_CHUNK_SLOT].push_back(0);
    std::vector<int> fcols, slot_off;
    std::vector<char> seen;
    while (r < n_rows) {
      int first_cell = I[I_ROW_CELL][r];
      int e = I[I_CELL_COL][first_cell];
      if (e >= n_ecols) break;
      int r1 = r;
      fcols.clear();
      while (r1 < n_rows && I[I_CELL_COL][I[I_ROW_CELL][r1]] == e) {
        for (int c = I[I_ROW_CELL][r1] + 1; c < I[I_ROW_CELL][r1 + 1]; ++c) fcols.push_back(I[I_CELL_COL][c]);
        ++r1;
      }
      std::sort(fcols.begin(), fcols.end());
      fcols.erase(std::unique(fcols.begin(), fcols.end()), fcols.end());
      const int es = col_size[e];
      if (es > MAX_WARP_E) return fail(SWGN_ERR_UNSUPPORTED, "eliminated parameter block larger than 16 tangent dimensions");
      const int chunk = (int)I[I_CHUNK_ECOL].size();
      slot_off.assign(fcols.size(), 0);
      int ncol = 0;
      for (size_t k = 0; k < fcols.size(); ++k) {
        slot_off[k] = n_ebuf;
        I[I_SLOT_COL].push_back(fcols[k]);
        I[I_SLOT_BUF].push_back(n_ebuf);
        n_ebuf += es * col_size[fcols[k]];
        ncol += col_size[fcols[k]];
      }
      I[I_CHUNK_G].push_back(n_ebuf);
      n_ebuf += es;
      n_ebuf = (int)align2(n_ebuf);
      seen.assign(fcols.size(), 0);
      for (int rr = r; rr < r1; ++rr)
        for (int c = I[I_ROW_CELL][rr] + 1; c < I[I_ROW_CELL][rr + 1]; ++c) {
          int fc = I[I_CELL_COL][c];
          int s = (int)(std::lower_bound(fcols.begin(), fcols.end(), fc) - fcols.begin());
          I[I_CELL_SLOT][c] = slot_off[s];
          I[I_CELL_FIRST][c] = seen[s] ? 0 : 1;
          seen[s] = 1;
        }
      if (es <= 3) {
        I[I_TCHUNK].push_back(chunk);
      } else {
        I[I_WCHUNK].push_back(chunk);
        max_wbuf = std::max(max_wbuf, MAX_WARP_E * MAX_WARP_E + es * (ncol + 1));
      }
      I[I_CHUNK_ECOL].push_back(e);
      I[I_CHUNK_FAC].push_back(n_efac);
      n_efac += (int)align2(es * es);
      I[I_CHUNK_ROW].push_back(r1);
      I[I_CHUNK_SLOT].push_back((int32_t)I[I_SLOT_COL].size());
      r = r1;
    }
  }
  const int n_chunks = (int)I[I_CHUNK_ECOL].size();
  if (n_chunks != n_ecols) return fail(SWGN_ERR_INVALID, "chunk detection does not match the eliminated blocks");
  const int n_slots = (int)I[I_SLOT_COL].size();
  const int n_jac_al = (int)align2(n_jac), n_ebuf_al = (int)align2(n_ebuf);
  if ((int64_t)n_jac_al + n_ebuf_al + n_res >= (1 << 28))
    return fail(SWGN_ERR_TOO_LARGE, "window Jacobian too large for the 28-bit gather offsets");

  // ---- gather streams (device_types.h "gather stream"): output tiles with their <= 4-row terms are dealt to
  // the SCHUR_WARPS warps of the window's CTA, longest first onto the least loaded warp; every warp gets
  // one linear stream of stages (header + SCHUR_STAGE terms of one tile, padded with switched-off terms)
  struct GTerm {
    uint32_t a, b, m, sign, b2;
  };
  struct TileJob {
    uint32_t meta, flags;
    int32_t x, y;  // header words 0 / 1: S offset and first S row, or e-cell output offset and its g offset
    const std::vector<GTerm>* terms;  // shared by all tiles of one cell (owned by term_store)
  };
  std::deque<std::vector<GTerm>> term_store;  // stable addresses
  auto deal_streams = [&](const std::vector<TileJob>& jobs, int arr_stream, int arr_ptr) {
    std::vector<size_t> idx(jobs.size());
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](size_t x, size_t y) { return jobs[x].terms->size() > jobs[y].terms->size(); });
    std::vector<std::vector<size_t>> mine(SCHUR_WARPS);
    std::vector<size_t> load(SCHUR_WARPS, 0);
    for (size_t j : idx) {
      const int wmin = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      mine[wmin].push_back(j);
      load[wmin] += (jobs[j].terms->size() + SCHUR_STAGE - 1) / SCHUR_STAGE + 1;
    }
    std::vector<int32_t>& WS = I[arr_stream];
    constexpr size_t kStageInts = 4 * (SCHUR_STAGE + 1);
    {
      size_t n_st_total = 0;
      for (const TileJob& jb : jobs) n_st_total += jb.terms->empty() ? 1 : (jb.terms->size() + SCHUR_STAGE - 1) / SCHUR_STAGE;
      WS.resize(WS.size() + n_st_total * kStageInts);
    }
    int32_t* out = WS.data();
    for (int wv = 0; wv < SCHUR_WARPS; ++wv) {
      I[arr_ptr].push_back((int32_t)((out - WS.data()) / kStageInts));
      for (size_t j : mine[wv]) {
        const TileJob& jb = jobs[j];
        const GTerm* terms = jb.terms->data();
        const size_t nt = jb.terms->size();
        const size_t n_st = nt == 0 ? 1 : (nt + SCHUR_STAGE - 1) / SCHUR_STAGE;
        for (size_t st = 0; st < n_st; ++st) {
          out[0] = jb.x;
          out[1] = jb.y;
          out[2] = (int32_t)(jb.flags | (st + 1 == n_st ? 1u : 0u));
          out[3] = (int32_t)jb.meta;
          out += 4;
          for (size_t e = st * SCHUR_STAGE; e < (st + 1) * SCHUR_STAGE; ++e, out += 4) {
            if (e < nt) {
              const GTerm& t = terms[e];
              out[0] = (int32_t)(t.a | ((t.m - 1) << 28) | (t.sign << 30));
              out[1] = (int32_t)t.b;
              out[2] = (int32_t)t.b2;
              out[3] = 0;
            } else {  // padding: bit 31 switches the loads off, the MMA adds zero
              out[0] = (int32_t)0x80000000u;
              out[1] = out[2] = out[3] = 0;
            }
          }
        }
      }
    }
    I[arr_ptr].push_back((int32_t)(WS.size() / (4 * (SCHUR_STAGE + 1))));
  };
  // ---- gather tables of the reduced system (device Schur kernel, phase 2).  Every touched block
  // cell (p, q), p <= q, of S lists its terms:
  //   + F_p' F_q of every row holding both cells      (schur_eliminator_impl.h:667-716, 569-661)
  //   - Wp' Wq   of every chunk holding both slots    (:514-563)
  // The rhs rides on the diagonal cells as one extra column: a diagonal cell's term list is exactly
  // the list of rows / chunks that touch block p, and rhs_p = sum_rows F_p' b - sum_chunks Wp' wg
  // (:381-422: b - E inv g expands to this), so its terms carry a third word, the offset of b / wg.
  const int ld = (n_f + 1 + 3) & ~3;
  {
    auto fpos = [&](int c) { return col_pos[c] - n_e; };  // row/col of the f-block inside S
    typedef GTerm T;
    // cell (p, q) -> term list, in a flat n_fb x n_fb table (lexicographic order = the order of a map keyed by
    // (p, q)); two passes: count, then fill into exactly sized lists
    const int n_fb = n_cols - n_ecols;
    auto cid = [&](int p, int q) { return (size_t)(p - n_ecols) * n_fb + (size_t)(q - n_ecols); };
    std::vector<uint32_t> cnt((size_t)n_fb * n_fb, 0);
    for (int r = 0; r < n_rows; ++r)
      for (int c1 = I[I_ROW_CELL][r]; c1 < I[I_ROW_CELL][r + 1]; ++c1) {
        const int p = I[I_CELL_COL][c1];
        if (p < n_ecols) continue;
        for (int c2 = c1; c2 < I[I_ROW_CELL][r + 1]; ++c2) ++cnt[cid(p, I[I_CELL_COL][c2])];
      }
    for (int ch = 0; ch < n_chunks; ++ch)
      for (int s1 = I[I_CHUNK_SLOT][ch]; s1 < I[I_CHUNK_SLOT][ch + 1]; ++s1)
        for (int s2 = s1; s2 < I[I_CHUNK_SLOT][ch + 1]; ++s2) ++cnt[cid(I[I_SLOT_COL][s1], I[I_SLOT_COL][s2])];
    std::vector<std::vector<T>> cells((size_t)n_fb * n_fb);
    for (size_t k = 0; k < cells.size(); ++k)
      if (cnt[k]) cells[k].reserve(cnt[k]);
    const uint32_t res_base = (uint32_t)(n_jac_al + n_ebuf_al);
    for (int r = 0; r < n_rows; ++r) {
      const int nres = I[I_ROW_NRES][r];
      for (int c1 = I[I_ROW_CELL][r]; c1 < I[I_ROW_CELL][r + 1]; ++c1) {
        const int p = I[I_CELL_COL][c1];
        if (p < n_ecols) continue;
        for (int c2 = c1; c2 < I[I_ROW_CELL][r + 1]; ++c2)
          cells[cid(p, I[I_CELL_COL][c2])].push_back({(uint32_t)I[I_CELL_VAL][c1], (uint32_t)I[I_CELL_VAL][c2], (uint32_t)nres, 0u,
                                                      res_base + (uint32_t)I[I_ROW_RES][r]});
      }
    }
    for (int ch = 0; ch < n_chunks; ++ch) {
      const uint32_t es = (uint32_t)col_size[I[I_CHUNK_ECOL][ch]];
      for (int s1 = I[I_CHUNK_SLOT][ch]; s1 < I[I_CHUNK_SLOT][ch + 1]; ++s1) {
        const int p = I[I_SLOT_COL][s1];
        for (int s2 = s1; s2 < I[I_CHUNK_SLOT][ch + 1]; ++s2)
          cells[cid(p, I[I_SLOT_COL][s2])].push_back({(uint32_t)(n_jac_al + I[I_SLOT_BUF][s1]), (uint32_t)(n_jac_al + I[I_SLOT_BUF][s2]), es, 1u,
                                                      (uint32_t)(n_jac_al + I[I_CHUNK_G][ch])});
      }
    }
    // heaviest cells first: the warps pull cells round-robin, so the tail is made of light cells
    struct OrderedCell {
      std::pair<int, int> first;
      const std::vector<T>* second;
      size_t weight;
    };
    std::vector<OrderedCell> order;
    for (int p = n_ecols; p < n_cols; ++p)
      for (int q = n_ecols; q < n_cols; ++q) {
        const std::vector<T>& ts = cells[cid(p, q)];
        if (ts.empty() && p != q) continue;  // diagonal cells always exist (D^2)
        size_t w = 0;
        for (const T& t : ts) w += t.m;
        const int outs = col_size[p] * (col_size[q] + (p == q ? 1 : 0));
        order.push_back({{p, q}, &ts, w * (size_t)((outs + 31) / 32) + 8});
      }
    std::stable_sort(order.begin(), order.end(), [](const OrderedCell& x, const OrderedCell& y) { return x.weight > y.weight; });
    std::vector<TileJob> jobs;
    for (size_t ci = 0; ci < order.size(); ++ci) {
      const int p = order[ci].first.first, q = order[ci].first.second;
      const int ps = col_size[p], qs = col_size[q];
      if (ps > MAX_COL_SIZE || qs > MAX_COL_SIZE) return fail(SWGN_ERR_UNSUPPORTED, "parameter block larger than 63 tangent dimensions");
      const bool diag = p == q;
      // flat term stream: blocks with more than 4 rows are split into K slabs of <= 4 rows, and every
      // entry packs (offset, rows, sign) so that the device loop is one straight-line batch after another
      //   word0 = a | rows-1 << 28 | subtract << 30        word1 = b        [word2 = b2, word3 = 0]
      term_store.emplace_back();
      std::vector<T>& ts = term_store.back();
      {
        size_t n_slabs = 0;
        for (const T& t : *order[ci].second) n_slabs += (t.m + 3) / 4;
        ts.reserve(n_slabs);
      }
      for (const T& t : *order[ci].second)
        for (uint32_t e0 = 0; e0 < t.m; e0 += 4)
          ts.push_back({t.a + e0 * (uint32_t)ps, t.b + e0 * (uint32_t)qs, std::min(4u, t.m - e0), t.sign, t.b2 + e0});
      {
        const int tiles = ((ps + 7) / 8) * ((qs + (diag ? 1 : 0) + 7) / 8);
        P->n_mma += (int64_t)tiles * (int64_t)ts.size();
      }
      for (int ti = 0; ti < ps; ti += 8)
        for (int tj = 0; tj < qs + (diag ? 1 : 0); tj += 8) {
          TileJob jb;
          jb.meta = (uint32_t)ps | ((uint32_t)qs << 6) | ((uint32_t)(ti / 8) << 12) | ((uint32_t)(tj / 8) << 15) | ((diag ? 1u : 0u) << 18);
          jb.flags = 0;
          jb.x = fpos(p) * ld + fpos(q);
          jb.y = fpos(p);
          jb.terms = &ts;
          jobs.push_back(jb);
        }
      const int32_t rec[8] = {ps, qs, fpos(p) * ld + fpos(q), 0, (int32_t)ts.size(), diag ? 1 : 0, 0, 0};
      I[I_SCELL].insert(I[I_SCELL].end(), rec, rec + 8);  // cell directory (statistics / read-backs); the kernel reads the streams
    }
    deal_streams(jobs, I_WSTREAM, I_WSTREAM_PTR);
    // ---- symbolic fill-in of the blocked Cholesky at the granularity k_chol works at: 32-row panels, 16-column
    // groups.  A[gi][gj] = "S or U can be non-zero somewhere in rows of group gi, columns of group gj"; eliminating
    // a panel couples all column groups present in its rows.  The rhs (column n_f) is dense.
    {
      const int ng = (n_f + 1 + 15) / 16, n_panels = (n_f + 31) / 32;
      if (ng <= 64) {
        std::vector<uint64_t> A(ng, 0);
        const int grhs = n_f / 16;
        for (size_t ci = 0; ci < order.size(); ++ci) {
          const int p = order[ci].first.first, q = order[ci].first.second;
          const int rp = fpos(p), cq = fpos(q);
          for (int gi = rp / 16; gi <= (rp + col_size[p] - 1) / 16; ++gi)
            for (int gj = cq / 16; gj <= (cq + col_size[q] - 1) / 16; ++gj) A[gi] |= 1ull << gj;
        }
        for (int g = 0; g < ng; ++g) A[g] |= (1ull << g) | (1ull << grhs);
        for (int k = 0; k < n_panels; ++k) {
          const int g0 = 2 * k, g1 = std::min(ng - 1, 2 * k + 1);
          uint64_t m = (A[g0] | A[g1] | (1ull << g0) | (1ull << g1)) & ~((1ull << g0) - 1);
          I[I_CHOL_MASK].push_back((int32_t)(uint32_t)(m & 0xffffffffu));
          I[I_CHOL_MASK].push_back((int32_t)(uint32_t)(m >> 32));
          for (int gi = g1 + 1; gi < ng; ++gi)
            if (m & (1ull << gi)) A[gi] |= m & ~((1ull << gi) - 1);
        }
      } else {
        for (int k = 0; k < 2 * n_panels; ++k) I[I_CHOL_MASK].push_back(-1);
      }
    }
  }
  // ---- raw products of the larger e-blocks (4..16 tangent dims: the speed-bias blocks), gathered
  // by the same tensor-core stream code as the reduced system (phase 1a): per chunk one "diagonal" cell
  // [E'E | E'b] (es x es+1, upper tiles only, written to W_EFAC / the chunk's g slot) and one cell
  // E'F_f per slot (es x fs, written to the slot's W_EBUF block); terms = the rows of the chunk.
  {
    const uint32_t res_base = (uint32_t)(n_jac_al + n_ebuf_al);
    std::vector<TileJob> jobs;
    int n_ecells = 0;
    auto emit = [&](int ps, int qs, int out, bool diag, int gout, const std::vector<GTerm>& rows) {
      term_store.emplace_back();
      std::vector<GTerm>& ts = term_store.back();
      for (const GTerm& t : rows)
        for (uint32_t e0 = 0; e0 < t.m; e0 += 4)
          ts.push_back({t.a + e0 * (uint32_t)ps, t.b + e0 * (uint32_t)qs, std::min(4u, t.m - e0), 0u, t.b2 + e0});
      ++n_ecells;
      for (int ti = 0; ti < ps; ti += 8)
        for (int tj = 0; tj < qs + (diag ? 1 : 0); tj += 8) {
          if (diag && tj + 7 < ti) continue;  // E'E: upper triangle only
          TileJob jb;
          jb.meta = (uint32_t)ps | ((uint32_t)qs << 6) | ((uint32_t)(ti / 8) << 12) | ((uint32_t)(tj / 8) << 15) | ((diag ? 1u : 0u) << 18);
          jb.flags = 2;  // e-cell target
          jb.x = out;
          jb.y = gout;
          jb.terms = &ts;
          jobs.push_back(jb);
          P->n_mma += (int64_t)ts.size();
        }
    };
    for (int wc = 0; wc < (int)I[I_WCHUNK].size(); ++wc) {
      const int ch = I[I_WCHUNK][wc];
      const int es = col_size[I[I_CHUNK_ECOL][ch]];
      std::vector<GTerm> diag_terms;
      std::map<int, std::vector<GTerm>> slot_terms;  // slot buffer offset -> rows
      for (int r = I[I_CHUNK_ROW][ch]; r < I[I_CHUNK_ROW][ch + 1]; ++r) {
        const uint32_t nres = (uint32_t)I[I_ROW_NRES][r];
        const int c0 = I[I_ROW_CELL][r];
        const uint32_t eoff = (uint32_t)I[I_CELL_VAL][c0];
        diag_terms.push_back({eoff, eoff, nres, 0u, res_base + (uint32_t)I[I_ROW_RES][r]});
        for (int c = c0 + 1; c < I[I_ROW_CELL][r + 1]; ++c)
          slot_terms[I[I_CELL_SLOT][c]].push_back({eoff, (uint32_t)I[I_CELL_VAL][c], nres, 0u, 0u});
      }
      emit(es, es, I[I_CHUNK_FAC][ch], true, I[I_CHUNK_G][ch], diag_terms);
      for (int s = I[I_CHUNK_SLOT][ch]; s < I[I_CHUNK_SLOT][ch + 1]; ++s)
        emit(es, col_size[I[I_SLOT_COL][s]], I[I_SLOT_BUF][s], false, -1, slot_terms[I[I_SLOT_BUF][s]]);
    }
    deal_streams(jobs, I_ESTREAM, I_ESTREAM_PTR);
    I[I_ECELL_G].assign(1, n_ecells);  // (statistic only)
  }
  // ---- row-parallel part of phase 1: rows of "simple" small chunks (every slot fed by exactly one
  // row, e.g. a landmark seen once per keyframe) compute their W block independently
  {
    I[I_ROW_CHUNK].assign(n_rows, -1);
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int es = col_size[I[I_CHUNK_ECOL][ch]];
      bool simple = es <= 3;
      for (int r = I[I_CHUNK_ROW][ch]; r < I[I_CHUNK_ROW][ch + 1]; ++r) {
        I[I_ROW_CHUNK][r] = ch;
        for (int c = I[I_ROW_CELL][r] + 1; c < I[I_ROW_CELL][r + 1]; ++c)
          if (!I[I_CELL_FIRST][c]) simple = false;
      }
      I[I_CHUNK_SIMPLE].push_back(simple ? 1 : 0);
      if (es <= 3) {
        // several rows feed one slot (e.g. all satellites of an epoch share the pose slot of the clock's chunk): a
        // single thread would walk the rows one dependent load chain after the other -> one warp, lanes over rows
        const int nrows_ch = I[I_CHUNK_ROW][ch + 1] - I[I_CHUNK_ROW][ch];
        bool warp_ok = !simple && nrows_ch >= 4 && nrows_ch <= 32 && (I[I_CHUNK_SLOT][ch + 1] - I[I_CHUNK_SLOT][ch]) <= 32;
        for (int r = I[I_CHUNK_ROW][ch]; warp_ok && r < I[I_CHUNK_ROW][ch + 1]; ++r)
          if (I[I_ROW_NRES][r] > 2 || I[I_ROW_CELL][r + 1] - I[I_ROW_CELL][r] - 1 > 4) warp_ok = false;
        I[warp_ok ? I_TCHUNK_W : I_TCHUNK_T].push_back(ch);
      }
      if (simple)
        for (int r = I[I_CHUNK_ROW][ch]; r < I[I_CHUNK_ROW][ch + 1]; ++r)
          if (I[I_ROW_CELL][r + 1] - I[I_ROW_CELL][r] > 1) {
            const int c0 = I[I_ROW_CELL][r], nf_cells = I[I_ROW_CELL][r + 1] - c0 - 1;
            const int32_t rec[8] = {I[I_CELL_VAL][c0], I[I_ROW_NRES][r] | (es << 8) | (nf_cells << 16), I[I_CHUNK_FAC][ch], c0 + 1,
                                    I[I_CELL_VAL][c0 + 1], I[I_CELL_SLOT][c0 + 1], col_size[I[I_CELL_COL][c0 + 1]], 0};
            I[I_SROW].insert(I[I_SROW].end(), rec, rec + 8);
          }
    }
  }
  // CSC
  {
    std::vector<int> cnt(n_cols + 1, 0);
    for (int c = 0; c < n_cells; ++c) cnt[I[I_CELL_COL][c] + 1]++;
    std::partial_sum(cnt.begin(), cnt.end(), cnt.begin());
    I[I_CSC_PTR].assign(cnt.begin(), cnt.end());
    I[I_CSC_ROW].assign(n_cells, 0);
    I[I_CSC_VAL].assign(n_cells, 0);
    I[I_CSC_NRES].assign(n_cells, 0);
    I[I_CSC_RES].assign(n_cells, 0);
    std::vector<int> cur(cnt.begin(), cnt.end() - 1);
    for (int r = 0; r < n_rows; ++r)
      for (int c = I[I_ROW_CELL][r]; c < I[I_ROW_CELL][r + 1]; ++c) {
        int col = I[I_CELL_COL][c];
        I[I_CSC_ROW][cur[col]] = r;
        I[I_CSC_VAL][cur[col]] = I[I_CELL_VAL][c];
        I[I_CSC_NRES][cur[col]] = I[I_ROW_NRES][r];
        I[I_CSC_RES][cur[col]] = I[I_ROW_RES][r];
        cur[col]++;
      }
  }
  for (int c = 0; c < n_cols; ++c) {
    I[I_COL_STATE].push_back(g->block_offset[cols[c]]);
    I[I_COL_SIZE].push_back(col_size[c]);
    I[I_COL_GSIZE].push_back(g->block_size[cols[c]]);
    I[I_COL_POS].push_back(col_pos[c]);
    I[I_COL_BLOCK].push_back(cols[c]);
    for (int k = 0; k < col_size[c]; ++k) I[I_TCOL].push_back(c);
  }

  // ---- factor tables (kind-major storage order); constants are packed by pack_constants()
  auto soff = [&](int b) { return g->block_offset[b]; };
  for (int i = 0; i < g->n_proj; ++i) {
    const Factor& f = fac[kind_begin[0] + i];
    int32_t rec[8] = {soff(f.blocks[0]), soff(f.blocks[1]), soff(f.blocks[2]), f.jac_off[0], f.jac_off[1],
                      f.jac_off[2], f.active ? f.res_off : -1, 0};
    I[I_PROJ].insert(I[I_PROJ].end(), rec, rec + 8);
  }
  for (int i = 0; i < g->n_imu; ++i) {
    const Factor& f = fac[kind_begin[1] + i];
    int32_t rec[12] = {soff(f.blocks[0]), soff(f.blocks[1]), soff(f.blocks[2]), soff(f.blocks[3]),
                       f.jac_off[0], f.jac_off[1], f.jac_off[2], f.jac_off[3], f.active ? f.res_off : -1, 0, 0, 0};
    I[I_IMU].insert(I[I_IMU].end(), rec, rec + 12);
  }
  for (int i = 0; i < g->n_gnss; ++i) {
    const Factor& f = fac[kind_begin[2] + i];
    int32_t rec[8] = {g->gnss_kind[i], soff(f.blocks[0]), soff(f.blocks[1]), f.blocks.size() > 2 ? soff(f.blocks[2]) : 0,
                      f.jac_off[0], f.jac_off[1], f.blocks.size() > 2 ? f.jac_off[2] : -1, f.active ? f.res_off : -1};
    I[I_GNSS].insert(I[I_GNSS].end(), rec, rec + 8);
  }
  int n_prior_blk = 0;
  {
    int64_t oJ = 0, oR = 0, oX = 0;
    for (int i = 0; i < g->n_prior; ++i) {
      const Factor& f = fac[kind_begin[3] + i];
      const int n = g->prior_n[i];
      if (n <= 0) return fail(SWGN_ERR_INVALID, "prior with no rows");
      int32_t rec[8] = {n, (int32_t)f.blocks.size(), f.active ? f.res_off : -1, n_prior_blk, (int32_t)oJ, (int32_t)oR, 0, 0};
      I[I_PRIOR].insert(I[I_PRIOR].end(), rec, rec + 8);
      for (size_t p = 0; p < f.blocks.size(); ++p) {
        int b = f.blocks[p];
        int idx = g->prior_blk_idx[g->prior_blk_begin[i] + p];
        int ls = local_size(g->block_size[b], g->block_manifold[b]);
        if (idx < 0 || idx + ls > n) return fail(SWGN_ERR_INVALID, "prior block column range outside J0");
        int32_t br[6] = {soff(b), g->block_size[b], idx, f.jac_off[p], (int32_t)oX, ls};
        I[I_PRIOR_BLK].insert(I[I_PRIOR_BLK].end(), br, br + 6);
        oX += g->block_size[b];
        ++n_prior_blk;
      }
      oJ += (int64_t)n * n;
      oR += n;
      if (oJ > INT32_MAX) return fail(SWGN_ERR_TOO_LARGE, "prior Jacobians too large");
    }
  }
  for (int i = 0; i < g->n_unit; ++i) {
    const Factor& f = fac[kind_begin[4] + i];
    int32_t rec[4] = {soff(f.blocks[0]), f.jac_off[0], f.active ? f.res_off : -1, 0};
    I[I_UNIT].insert(I[I_UNIT].end(), rec, rec + 4);
  }
  int64_t chain_work = 0;
  {
    int64_t oc = 0;
    int frame0 = 0;
    for (int i = 0; i < g->n_chain; ++i) {
      const Factor& f = fac[kind_begin[5] + i];
      const int m = g->chain_frame_begin[i + 1] - g->chain_frame_begin[i], k = (int)f.blocks.size() - 4;
      const ChainLayout L(m, k);
      if (oc + L.c_size > INT32_MAX || chain_work + L.w_size > INT32_MAX) return fail(SWGN_ERR_TOO_LARGE, "chain factors too large");
      int32_t rec[8] = {m, k, f.active ? f.res_off : -1, (int32_t)(I[I_CHAIN_BLK].size() / 2), (int32_t)oc, (int32_t)chain_work, frame0, 0};
      I[I_CHAIN].insert(I[I_CHAIN].end(), rec, rec + 8);
      for (size_t p = 0; p < f.blocks.size(); ++p) {
        I[I_CHAIN_BLK].push_back(soff(f.blocks[p]));
        I[I_CHAIN_BLK].push_back(f.jac_off[p]);
      }
      oc += L.c_size;
      chain_work += L.w_size;
      frame0 += m;
    }
  }
  int64_t hostbuf = 0;
  for (int i = 0; i < g->n_host; ++i) {
    const Factor& f = fac[kind_begin[6] + i];
    if (hostbuf > INT32_MAX / 2) return fail(SWGN_ERR_TOO_LARGE, "host-evaluated factors too large");
    int32_t rec[4] = {f.active ? f.res_off : -1, (int32_t)(I[I_HOST_BLK].size() / 4), (int32_t)f.blocks.size(), (int32_t)hostbuf};
    I[I_HOST].insert(I[I_HOST].end(), rec, rec + 4);
    hostbuf += f.nres;
    for (size_t p = 0; p < f.blocks.size(); ++p) {
      const int b = f.blocks[p];
      int32_t br[4] = {soff(b), f.jac_off[p], g->block_size[b], local_size(g->block_size[b], g->block_manifold[b])};
      I[I_HOST_BLK].insert(I[I_HOST_BLK].end(), br, br + 4);
      hostbuf += (int64_t)f.nres * g->block_size[b];
    }
  }
  {
    int64_t sizes[NUM_CARR];
    constant_sizes(g, sizes);
    double* ptr[NUM_CARR];
    for (int a = 0; a < NUM_CARR; ++a) {
      P->carr[a].assign((size_t)sizes[a], 0.0);
      ptr[a] = P->carr[a].data();
    }
    pack_constants(g, ptr);
  }

  // ---- streamed Schur elimination (plan_stream.cpp); windows that do not fit its on-chip budget keep the gather kernel
  P->sb = StreamPlanInfo();
  if (P->want_stream_plan || stream_enabled()) build_stream_plan(P, n_rows, n_cols, n_ecols, n_jac, n_res, col_size, col_pos);

  // ---- descriptor
  WinDesc& d = P->d;
  d = WinDesc();
  d.sb_ok = P->sb.ok;
  d.sb_nbatch = P->sb.nbatch;
  d.sb_acc = P->sb.acc;
  d.sb_jcap = P->sb.jcap;
  d.sb_rcap = P->sb.rcap;
  d.sb_ecap = P->sb.ecap;
  d.sb_fcap = P->sb.fcap;
  d.sb_reccap = P->sb.reccap;
  d.n_fb = n_cols - n_ecols;
  d.n_state = g->n_state;
  d.n_cols = n_cols;
  d.n_ecols = n_ecols;
  d.n_e = n_e;
  d.n_f = n_f;
  d.n_t = n_t;
  d.n_res = n_res;
  d.n_rows = n_rows;
  d.n_cells = n_cells;
  d.n_chunks = n_chunks;
  d.n_slots = n_slots;
  d.n_jac = n_jac;
  d.n_ebuf = n_ebuf;
  d.ld = ld;
  d.n_proj = g->n_proj;
  d.n_imu = g->n_imu;
  d.n_gnss = g->n_gnss;
  d.n_prior = g->n_prior;
  d.n_prior_blk = n_prior_blk;
  d.n_unit = g->n_unit;
  d.n_efac = n_efac;
  d.n_tchunks = (int)I[I_TCHUNK].size();
  d.n_tchunks_t = (int)I[I_TCHUNK_T].size();
  d.n_tchunks_w = (int)I[I_TCHUNK_W].size();
  d.n_wchunks = (int)I[I_WCHUNK].size();
  d.n_scells = (int)(I[I_SCELL].size() / 8);
  d.n_sterms = (int)I[I_STERM].size();
  d.n_srows = (int)(I[I_SROW].size() / 8);
  d.n_ecells = I[I_ECELL_G].empty() ? 0 : I[I_ECELL_G][0];
  d.max_wbuf = max_wbuf;
  d.n_wstream = (int)(I[I_WSTREAM].size() / 4);
  d.n_chain = g->n_chain;
  d.n_host = g->n_host;
  d.n_hostbuf = (int32_t)hostbuf;
  d.n_chain_frames = n_chain_frames;
  d.max_chain_k = max_chain_k;
  d.max_prior_n = 0;
  for (int i = 0; i < g->n_prior; ++i) d.max_prior_n = std::max(d.max_prior_n, g->prior_n[i]);
  // tangent size of the trailing parameter_head groups (UpdateSchurHessianOnly's n)
  d.n_head = 0;
  if (n_parameter_head > 0) {
    if (n_parameter_head > n_cols - n_ecols) return fail(SWGN_ERR_INVALID, "n_parameter_head exceeds the retained blocks");
    for (int c = n_cols - n_parameter_head; c < n_cols; ++c) d.n_head += col_size[c];
  }
  P->state.assign(g->state, g->state + g->n_state);
  int64_t* W = P->wsize;
  W[W_X] = W[W_XCAND] = W[W_XBEST] = W[W_X0] = align2(g->n_state);
  W[W_JAC] = n_jac_al;
  W[W_EBUF] = n_ebuf_al;
  W[W_RES] = W[W_MRES] = align2(n_res);
  W[W_DIAG] = W[W_G] = W[W_GHAT] = W[W_GN] = W[W_STEP] = W[W_Y] = W[W_LMD] = W[W_SCALE] = align2(n_t);
  W[W_EFAC] = align2(n_efac);
  W[W_S] = W[W_SCOPY] = align2((int64_t)n_f * d.ld);
  W[W_CHAIN] = align2(chain_work);
  W[W_HOSTBUF] = align2(hostbuf);
  P->schur_doubles = schur_doubles + n_t + (int64_t)n_f * (n_f + 1) / 2 + n_f + n_e;
  return SWGN_OK;
}

}  // namespace swgn
