// Per-epoch GNSS linearisation (include/swgn_gnss.h): host side.
//
//   tracker                = the three std::list<PBtype>[MAXSATNUM*2] of SWFOptimization (RVI/swf/swf.h:274-278)
//   phase A (host)         = SPP correction + ambiguity lookup                    RVI/swf/swf_gnss.cpp:271-325
//   launch (device)        = update_azel + gating residuals, k_gnss_epoch.cu      :346-377, common_function.cpp:394-408
//   phase B (host)         = medians, slip conditions, new ambiguities, counters  :378-500
//   pack                   = AddGnssResidual -> SWGN_GNSS_* records               RVI/swf/swf_core.cpp:87-205
//   pass 1 (device, batch) = export-mode solve, clocks in elimination group 0, reduced system over the keep blocks
//                            -> eigen square root = marg_info_gnss                :504-530, marginalization_factor.cpp:260-377
//   pass 2 (device, batch) = 2-iteration LEVENBERG_MARQUARDT + jacobi_scaling solve, pose / speed-bias / old ambiguities
//                            constant                                             :532-571
// There is no CPU path for the numerical parts: without a device the call fails like swgn_batch_create does.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/swgn_gnss.h"

namespace swgn {
swgn_status set_error(swgn_status st, const std::string& m);
void keep_pool_memory(int device);
swgn_status batch_marginal_priors_to(swgn_batch* b, const int32_t* n_tail, double* const* J0_ptr, double* const* r0_ptr);
cudaError_t launch_gate_residuals(int n, const double* rec_dev, const int32_t* flags_dev, double azelmin, double* out_dev,
                                  cudaStream_t s);
}  // namespace swgn
using swgn::set_error;

struct swgn_gnss_tracker {
  swgn_gnss_config cfg;
  std::vector<swgn_ambiguity> amb[3];
  std::vector<int32_t> lists[3][SWGN_MAXSAT * 2];  // handles in push_back order; back() = the list's last element
  int back(int fam, int idx) const {
    const std::vector<int32_t>& l = lists[fam][idx];
    return l.empty() ? -1 : l.back();
  }
  int push(int fam, int sat, int sys, int f) {  // PBtype n; n.value = 0; n.continue_count = 0; push_back(n)
    swgn_ambiguity a;
    std::memset(&a, 0, sizeof(a));
    a.sys = (uint8_t)sys;
    a.f = (uint8_t)f;
    a.sat = sat;
    a.alive = 1;
    amb[fam].push_back(a);
    const int h = (int)amb[fam].size() - 1;
    lists[fam][sat * 2 + f].push_back(h);
    return h;
  }
};

namespace {
#define CUG(call)                                                                               \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      st = set_error(SWGN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
      goto done;                                                                                \
    }                                                                                           \
  } while (0)

inline double sqr(double x) { return x * x; }
// 1 / sqrt(varerr2(el, dt, var)), gnss_factor.cpp:98-103 (single-precision sinf, as the reference)
inline double rtk_weight(double el, double dt, double mea_var) {
  const double b = 299792458.0 * 5e-12 * dt;
  const double sinel = sinf((float)el);
  return 1.0 / std::sqrt((mea_var / sinel / sinel) + b * b);
}

enum RefType { R_NONE = -1, R_POSE = 0, R_SB = 1, R_BLACK = 2, R_CLK = 3, R_AMB = 4 /* + family */ };
struct Ref {
  int type = R_NONE, id = 0;
};
struct Factor {
  int kind;
  Ref b[3];
  double data[SWGN_GNSS_STRIDE];
};

void gnss_record(Factor* f, int kind, const swgn_obs& d, const double* base, double meas, double lam, double w) {
  f->kind = kind;
  std::memset(f->data, 0, sizeof(f->data));
  for (int i = 0; i < 3; ++i) {
    f->data[SWGN_GNSS_SAT_POS + i] = d.sat_pos[i];
    f->data[SWGN_GNSS_SAT_VEL + i] = kind == SWGN_GNSS_DOPPLER ? d.sat_vel[i] : 0.0;
    f->data[SWGN_GNSS_BASE_POS + i] = base[i];
  }
  f->data[SWGN_GNSS_MEAS] = meas;
  f->data[SWGN_GNSS_LAM] = lam;
  f->data[SWGN_GNSS_WEIGHT] = w;
}

// AddGnssResidual, swf_core.cpp:87-205 (the InitialBlackFactor of :101-103 is the graph's unit factor)
void add_gnss_residual(const swgn_gnss_config& c, const swgn_epoch& e, const swgn_gnss_frame& fr, std::vector<Factor>* out) {
  bool have_base = false;
  Factor f;
  if (c.use_rtk) {
    for (int i = 0; i < e.n_obs; ++i) {
      const swgn_obs& d = e.obs[i];
      for (int q = 0; q < SWGN_NFREQ; ++q) {
        if (d.rtk_n[q] < 0) continue;
        if (d.el < c.azelmin) continue;
        have_base = true;
        const double lam = c.lams[d.sys][q];
        gnss_record(&f, SWGN_GNSS_RTK_CARRIER, d, e.base_xyz, d.rtk_l[q] * lam, lam,
                    rtk_weight(d.el, e.br_time_diff, std::pow(d.rtk_lstd[q] * lam, 2)));
        f.data[SWGN_GNSS_EL] = d.el;
        f.data[SWGN_GNSS_DT] = e.br_time_diff;
        f.data[SWGN_GNSS_VAR] = std::pow(d.rtk_lstd[q] * lam, 2);
        f.b[0] = {R_POSE, 0};
        f.b[1] = {R_AMB + SWGN_AMB_RTK, d.rtk_n[q]};
        f.b[2] = {R_CLK, d.sys * 2 + q};
        out->push_back(f);
      }
    }
  }
  if (c.use_rtd) {
    for (int i = 0; i < e.n_obs; ++i) {
      const swgn_obs& d = e.obs[i];
      for (int q = 0; q < SWGN_NFREQ; ++q) {
        if (d.rtk_p[q] == 0.0 || d.svh != 0 || d.rtk_pstd[q] > 2) continue;
        if (d.el < c.azelmin) continue;
        have_base = true;
        gnss_record(&f, SWGN_GNSS_RTK_PSEUDORANGE, d, e.base_xyz, d.rtk_p[q], 0.0,
                    rtk_weight(d.el, e.br_time_diff, std::pow(d.rtk_pstd[q], 2)));
        f.data[SWGN_GNSS_EL] = d.el;
        f.data[SWGN_GNSS_DT] = e.br_time_diff;
        f.data[SWGN_GNSS_VAR] = std::pow(d.rtk_pstd[q], 2);
        f.b[0] = {R_POSE, 0};
        f.b[1] = {R_CLK, d.sys * 2 + q};
        f.b[2] = Ref();
        out->push_back(f);
      }
    }
  }
  for (int i = 0; i < e.n_obs; ++i) {
    const swgn_obs& d = e.obs[i];
    if (d.svh != 0) continue;
    if (d.el < c.azelmin) continue;
    const double sin_el = std::sin(d.el);
    const double model_var = d.ion_var * 0.125 * 0.125 + d.trop_var * 0.7 * 0.7 + d.sat_var * 0.35 * 0.35;
    if (d.spp_p[0] != 0.0 && d.spp_pstd[0] < 2 && !have_base) {
      double istd = sin_el * sin_el / std::sqrt(sqr(d.spp_pstd[0]) + (model_var + 1));
      if (fr.epochs_since_start < 100) istd *= 10;
      gnss_record(&f, SWGN_GNSS_SPP_PSEUDORANGE, d, e.base_xyz, d.spp_p[0], 0.0, istd);
      f.b[0] = {R_POSE, 0};
      f.b[1] = {R_CLK, 6 + d.sys * 2};
      f.b[2] = Ref();
      out->push_back(f);
    }
    if (c.use_spp_phase && d.spp_l[0] != 0.0 && d.spp_n[0] >= 0) {
      const double lam = c.lams[d.sys][0];
      const double istd = sin_el * sin_el / std::sqrt(sqr(d.spp_lstd[0] * lam) + model_var);
      gnss_record(&f, SWGN_GNSS_SPP_CARRIER, d, e.base_xyz, d.spp_l[0] * lam, lam, istd);
      f.b[0] = {R_POSE, 0};
      f.b[1] = {R_CLK, 6 + d.sys * 2};
      f.b[2] = {R_AMB + SWGN_AMB_SPP, d.spp_n[0]};
      out->push_back(f);
    }
    if (c.use_spp_correction && d.spp_p0[0] != 0.0 && d.pcorr_n[0] >= 0) {
      const double lam = c.lams[d.sys][0];
      const double istd = sin_el * sin_el / std::sqrt(sqr(d.spp_pstd[0]) + model_var);
      gnss_record(&f, SWGN_GNSS_SPP_CARRIER, d, e.base_xyz, d.spp_p0[0], lam, istd);
      f.b[0] = {R_POSE, 0};
      f.b[1] = {R_CLK, 6 + d.sys * 2};
      f.b[2] = {R_AMB + SWGN_AMB_PCORR, d.pcorr_n[0]};
      out->push_back(f);
    }
  }
  if (c.use_doppler) {
    for (int i = 0; i < e.n_obs; ++i) {
      const swgn_obs& d = e.obs[i];
      if (d.spp_d[0] == 0.0 || d.svh != 0) continue;
      if (d.spp_dstd[0] > 2) continue;
      if (d.el < c.azelmin) continue;
      const double lam = c.lams[d.sys][0];
      const double istd = std::sin(d.el) * std::sin(d.el) / (d.spp_dstd[0] * lam);
      gnss_record(&f, SWGN_GNSS_DOPPLER, d, e.base_xyz, d.spp_d[0] * lam, 0.0, istd);
      f.b[0] = {R_SB, 0};
      f.b[1] = {R_CLK, 12};
      f.b[2] = {R_POSE, 0};
      out->push_back(f);
    }
  }
}

// the epoch's block list: pose, speed-bias (only when a Doppler factor reads it), blackvalue, the clock slots in use
// (ascending), then the ambiguities the factors point at -- RTK, SPP, pseudorange-correction, observation order
struct EpochBlocks {
  bool has_pose = false, has_sb = false;
  std::vector<int> clk_slots;
  std::vector<int> amb_family, amb_handle;
  int b_pose = -1, b_sb = -1, b_black = -1, b_clk0 = -1, b_amb0 = -1, n_blocks = 0;
  int clk_block[SWGN_GNSS_NCLK];
  int find_amb(int fam, int h) const {
    for (size_t i = 0; i < amb_family.size(); ++i)
      if (amb_family[i] == fam && amb_handle[i] == h) return b_amb0 + (int)i;
    return -1;
  }
  int block_of(const Ref& r) const {
    switch (r.type) {
      case R_NONE: return -1;
      case R_POSE: return b_pose;
      case R_SB: return b_sb;
      case R_BLACK: return b_black;
      case R_CLK: return clk_block[r.id];
      default: return find_amb(r.type - R_AMB, r.id);
    }
  }
};

void assign_blocks(const swgn_epoch& e, const std::vector<Factor>& fs, EpochBlocks* B) {
  bool clk_used[SWGN_GNSS_NCLK] = {false};
  for (const Factor& f : fs)
    for (int k = 0; k < 3; ++k) {
      if (f.b[k].type == R_POSE) B->has_pose = true;
      if (f.b[k].type == R_SB) B->has_sb = true;
      if (f.b[k].type == R_CLK) clk_used[f.b[k].id] = true;
    }
  auto used = [&](int fam, int h) {
    for (const Factor& f : fs)
      for (int k = 0; k < 3; ++k)
        if (f.b[k].type == R_AMB + fam && f.b[k].id == h) return true;
    return false;
  };
  for (int fam = 0; fam < 3; ++fam)
    for (int i = 0; i < e.n_obs; ++i)
      for (int q = 0; q < SWGN_NFREQ; ++q) {
        const swgn_obs& d = e.obs[i];
        const int h = fam == 0 ? d.rtk_n[q] : fam == 1 ? d.spp_n[q] : d.pcorr_n[q];
        if (h >= 0 && used(fam, h)) {
          bool dup = false;
          for (size_t k = 0; k < B->amb_handle.size(); ++k) dup |= B->amb_family[k] == fam && B->amb_handle[k] == h;
          if (!dup) {
            B->amb_family.push_back(fam);
            B->amb_handle.push_back(h);
          }
        }
      }
  int nb = 0;
  if (B->has_pose) B->b_pose = nb++;
  if (B->has_sb) B->b_sb = nb++;
  B->b_black = nb++;
  B->b_clk0 = nb;
  for (int s = 0; s < SWGN_GNSS_NCLK; ++s) {
    B->clk_block[s] = -1;
    if (clk_used[s]) {
      B->clk_block[s] = nb++;
      B->clk_slots.push_back(s);
    }
  }
  B->b_amb0 = nb;
  nb += (int)B->amb_handle.size();
  B->n_blocks = nb;
}

// one epoch as a swgn_graph: storage + the struct pointing into it
struct EpochGraph {
  EpochBlocks B;
  std::vector<Factor> factors;
  std::vector<int32_t> size, manifold, konst, group, offset;
  std::vector<double> state;
  std::vector<int32_t> kind, blocks;
  std::vector<double> data;
  int32_t unit_block = 0;
  double unit_istd = 1.0;  // InitialBlackFactor(1), swf_core.cpp:102
  swgn_graph g;
  int n_keep_tangent = 0;

  void build(const swgn_gnss_tracker& t, const swgn_epoch& e, const swgn_gnss_frame& fr) {
    add_gnss_residual(t.cfg, e, fr, &factors);
    assign_blocks(e, factors, &B);
    const int nb = B.n_blocks;
    size.assign(nb, 1);
    manifold.assign(nb, SWGN_MANIFOLD_EUCLIDEAN);
    konst.assign(nb, 0);
    group.assign(nb, 1);
    offset.assign(nb, 0);
    if (B.b_pose >= 0) {
      size[B.b_pose] = 7;
      manifold[B.b_pose] = SWGN_MANIFOLD_POSE;
    }
    if (B.b_sb >= 0) size[B.b_sb] = 9;
    for (int s : B.clk_slots) group[B.clk_block[s]] = 0;
    int off = 0;
    for (int b = 0; b < nb; ++b) {
      offset[b] = off;
      off += size[b];
    }
    state.assign(off, 0.0);
    n_keep_tangent = (B.b_pose >= 0 ? 6 : 0) + (B.b_sb >= 0 ? 9 : 0) + 1 + (int)B.amb_handle.size();
    kind.clear();
    blocks.clear();
    data.clear();
    for (const Factor& f : factors) {
      kind.push_back(f.kind);
      for (int k = 0; k < 3; ++k) blocks.push_back(B.block_of(f.b[k]));
      data.insert(data.end(), f.data, f.data + SWGN_GNSS_STRIDE);
    }
    unit_block = B.b_black;
    std::memset(&g, 0, sizeof(g));
  }
  // values: the frame's states; ambiguities at 0 (PhaseBiasSaveAndReset) or at their tracked values
  void fill_state(const swgn_gnss_tracker& t, const swgn_gnss_frame& fr, bool zero_ambiguities) {
    if (B.b_pose >= 0) std::memcpy(&state[offset[B.b_pose]], fr.pose, sizeof(double) * 7);
    if (B.b_sb >= 0) std::memcpy(&state[offset[B.b_sb]], fr.speed_bias, sizeof(double) * 9);
    state[offset[B.b_black]] = fr.blackvalue;
    for (int s : B.clk_slots) state[offset[B.clk_block[s]]] = fr.gnss_dt[s];
    for (size_t i = 0; i < B.amb_handle.size(); ++i)
      state[offset[B.b_amb0 + (int)i]] = zero_ambiguities ? 0.0 : t.amb[B.amb_family[i]][B.amb_handle[i]].value;
  }
  void point() {
    g.n_blocks = B.n_blocks;
    g.block_size = size.data();
    g.block_manifold = manifold.data();
    g.block_const = konst.data();
    g.block_group = group.data();
    g.block_offset = offset.data();
    g.n_state = (int32_t)state.size();
    g.state = state.data();
    g.proj_sqrt_info[0] = g.proj_sqrt_info[3] = 1.0;
    g.n_gnss = (int32_t)kind.size();
    g.gnss_kind = kind.data();
    g.gnss_blocks = blocks.data();
    g.gnss_data = data.data();
    g.n_unit = 1;
    g.unit_block = &unit_block;
    g.unit_istd = &unit_istd;
  }
};

// receivers are independent: the host phases run over them on all host threads
template <class F>
void parallel_for(int n, F fn) {
  const int nt = std::max(1, std::min<int>(n / 64, (int)std::thread::hardware_concurrency()));
  if (nt <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      const int i0 = next.fetch_add(16);
      if (i0 >= n) break;
      for (int i = i0; i < std::min(n, i0 + 16); ++i) fn(i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}

double median_of(std::vector<double>& v) {  // std::sort + v[size / 2], swf_gnss.cpp:381-388
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}
}  // namespace

extern "C" {

void swgn_gnss_config_default(swgn_gnss_config* c) {
  if (!c) return;
  std::memset(c, 0, sizeof(*c));
  c->use_imu = c->use_rtk = c->use_rtd = c->use_doppler = 1;  // rtk_visual_inertial_config.yaml
  c->use_spp_phase = c->use_spp_correction = 0;
  c->phase_all_reset_count = 10;
  c->estimate_pcorrection_period = 500;
  c->azelmin = 25.0 / 180 * 3.1415926535897932;
  const double lams[3][2] = {{0.190293672798364871256993069437, 0.244210213424568250983881512184},
                             {0.19203948631027648, 0.24834936958430670},
                             {0.19029367279836487, 0.24834936958430670}};
  std::memcpy(c->lams, lams, sizeof(lams));
  c->ambiguity_timeout = 10.0;
  c->slip_fraction_rtk = 0.5;
  c->init_max_iterations = 2;
  c->init_constant_after = 10;
  c->init_radius = 1e15;
  c->device = 0;
}

swgn_status swgn_gnss_tracker_create(const swgn_gnss_config* cfg, swgn_gnss_tracker** out) {
  if (!cfg || !out) return set_error(SWGN_ERR_INVALID, "bad arguments");
  swgn_gnss_tracker* t = new swgn_gnss_tracker();
  t->cfg = *cfg;
  *out = t;
  return SWGN_OK;
}
void swgn_gnss_tracker_destroy(swgn_gnss_tracker* t) { delete t; }
int32_t swgn_gnss_tracker_count(const swgn_gnss_tracker* t, int32_t family) {
  return (!t || family < 0 || family > 2) ? -1 : (int32_t)t->amb[family].size();
}
swgn_status swgn_gnss_tracker_get(const swgn_gnss_tracker* t, int32_t family, int32_t handle, swgn_ambiguity* out) {
  if (!t || !out || family < 0 || family > 2 || handle < 0 || handle >= (int32_t)t->amb[family].size())
    return set_error(SWGN_ERR_INVALID, "bad ambiguity handle");
  *out = t->amb[family][handle];
  return SWGN_OK;
}
swgn_status swgn_gnss_tracker_set_value(swgn_gnss_tracker* t, int32_t family, int32_t handle, double value) {
  if (!t || family < 0 || family > 2 || handle < 0 || handle >= (int32_t)t->amb[family].size())
    return set_error(SWGN_ERR_INVALID, "bad ambiguity handle");
  t->amb[family][handle].value = value;
  return SWGN_OK;
}
swgn_status swgn_gnss_tracker_erase(swgn_gnss_tracker* t, int32_t family, int32_t handle) {
  if (!t || family < 0 || family > 2 || handle < 0 || handle >= (int32_t)t->amb[family].size())
    return set_error(SWGN_ERR_INVALID, "bad ambiguity handle");
  swgn_ambiguity& a = t->amb[family][handle];
  if (!a.alive) return SWGN_OK;
  a.alive = 0;
  std::vector<int32_t>& l = t->lists[family][a.sat * 2 + a.f];
  l.erase(std::remove(l.begin(), l.end(), handle), l.end());
  return SWGN_OK;
}

swgn_status swgn_gnss_chain_frame(const swgn_gnss_output* P, const double* pose, const double* speed_bias, int32_t k,
                                  const int32_t* keep_slot, double* frame, double* frame_N, double* chain_N) {
  if (!P || !pose || !speed_bias || k < 0 || !keep_slot || !frame || (k > 0 && (!frame_N || !chain_N)) || P->n <= 0 || !P->J0 || !P->r0)
    return set_error(SWGN_ERR_INVALID, "bad arguments");
  const int n = P->n;
  // information form of the prior (marg_info_gnss->A, ->b)
  std::vector<double> A((size_t)n * n, 0.0), b(n, 0.0);
  for (int r = 0; r < n; ++r)
    for (int i = 0; i < n; ++i) {
      const double ji = P->J0[(size_t)r * n + i];
      if (ji == 0.0) continue;
      b[i] += ji * P->r0[r];
      for (int j = 0; j < n; ++j) A[(size_t)i * n + j] += ji * P->J0[(size_t)r * n + j];
    }
  std::fill(frame, frame + SWGN_CHAIN_FRAME_STRIDE, 0.0);
  std::fill(frame_N, frame_N + (size_t)15 * k, 0.0);
  std::memcpy(frame + SWGN_CHAIN_POSE, pose, sizeof(double) * 7);
  std::memcpy(frame + SWGN_CHAIN_SB, speed_bias, sizeof(double) * 9);
  std::memcpy(frame + SWGN_CHAIN_POSE_LIN, pose, sizeof(double) * 7);       // when the prior has no such block
  std::memcpy(frame + SWGN_CHAIN_SB_LIN, speed_bias, sizeof(double) * 9);
  int xo = 0;
  struct Blk {
    int idx, row, size, slot;  // column in the prior, row in the 15-vector (-1: phase bias), tangent size, phase-bias slot
  };
  std::vector<Blk> blk;
  for (int i = 0; i < P->n_keep; ++i) {
    const int kind = P->keep_kind[i];
    if (kind == SWGN_KEEP_POSE) {
      std::memcpy(frame + SWGN_CHAIN_POSE_LIN, P->x0 + xo, sizeof(double) * 7);
      blk.push_back({P->keep_idx[i], 0, 6, -1});
      xo += 7;
    } else if (kind == SWGN_KEEP_SPEED_BIAS) {
      std::memcpy(frame + SWGN_CHAIN_SB_LIN, P->x0 + xo, sizeof(double) * 9);
      blk.push_back({P->keep_idx[i], 6, 9, -1});
      xo += 9;
    } else {
      if (keep_slot[i] < 0 || keep_slot[i] >= k) return set_error(SWGN_ERR_INVALID, "a size-1 keep block needs a phase-bias slot");
      blk.push_back({P->keep_idx[i], -1, 1, keep_slot[i]});
      xo += 1;
    }
  }
  double* H = frame + SWGN_CHAIN_HESSIAN;
  double* rhs = frame + SWGN_CHAIN_RHS;
  for (const Blk& p : blk)
    for (const Blk& q : blk)
      for (int a = 0; a < p.size; ++a)
        for (int c = 0; c < q.size; ++c) {
          const double v = A[(size_t)(p.idx + a) * n + q.idx + c];
          if (p.row >= 0 && q.row >= 0) H[(p.row + a) * 15 + q.row + c] += v;
          else if (p.row >= 0) frame_N[(size_t)(p.row + a) * k + q.slot] += v;
          else if (q.row < 0) chain_N[(size_t)p.slot * k + q.slot] += v;
        }
  for (const Blk& p : blk)
    for (int a = 0; a < p.size; ++a) {
      if (p.row >= 0) rhs[p.row + a] += b[p.idx + a];
      else chain_N[(size_t)k * k + p.slot] += b[p.idx + a];
    }
  return SWGN_OK;
}

swgn_status swgn_gnss_gate_residuals(int32_t n_obs, const double* rec, double* out, int32_t device) {
  if (n_obs < 0 || (n_obs > 0 && (!rec || !out))) return set_error(SWGN_ERR_INVALID, "bad arguments");
  if (n_obs == 0) return SWGN_OK;
  swgn_status st = SWGN_OK;
  double *d_rec = nullptr, *d_out = nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return set_error(SWGN_ERR_NO_DEVICE, "no usable CUDA device (there is no CPU fallback)");
  CUG(cudaMalloc(&d_rec, sizeof(double) * 16 * n_obs));
  CUG(cudaMalloc(&d_out, sizeof(double) * 3 * n_obs));
  CUG(cudaMemcpy(d_rec, rec, sizeof(double) * 16 * n_obs, cudaMemcpyHostToDevice));
  CUG(swgn::launch_gate_residuals(n_obs, d_rec, nullptr, -1.0, d_out, 0));
  CUG(cudaMemcpy(out, d_out, sizeof(double) * 3 * n_obs, cudaMemcpyDeviceToHost));
done:
  cudaFree(d_rec);
  cudaFree(d_out);
  return st;
}

swgn_status swgn_gnss_epoch_records(const swgn_gnss_tracker* t, const swgn_epoch* e, const swgn_gnss_frame* f,
                                    int32_t* n_factors, int32_t* kind, int32_t* blocks, double* data, int32_t* n_clk,
                                    int32_t* clk_slot, int32_t* n_amb, int32_t* amb_family, int32_t* amb_handle) {
  if (!t || !e || !f || !n_factors) return set_error(SWGN_ERR_INVALID, "bad arguments");
  EpochGraph G;
  G.build(*t, *e, *f);
  *n_factors = (int32_t)G.kind.size();
  if (kind) std::copy(G.kind.begin(), G.kind.end(), kind);
  if (blocks) std::copy(G.blocks.begin(), G.blocks.end(), blocks);
  if (data) std::copy(G.data.begin(), G.data.end(), data);
  if (n_clk) *n_clk = (int32_t)G.B.clk_slots.size();
  if (clk_slot) std::copy(G.B.clk_slots.begin(), G.B.clk_slots.end(), clk_slot);
  if (n_amb) *n_amb = (int32_t)G.B.amb_handle.size();
  if (amb_family) std::copy(G.B.amb_family.begin(), G.B.amb_family.end(), amb_family);
  if (amb_handle) std::copy(G.B.amb_handle.begin(), G.B.amb_handle.end(), amb_handle);
  return SWGN_OK;
}

swgn_status swgn_gnss_preprocess(int32_t n, swgn_gnss_tracker* const* trackers, swgn_epoch* const* epochs,
                                 swgn_gnss_frame* frames, swgn_gnss_output* outputs) {
  if (n <= 0 || !trackers || !epochs || !frames || !outputs) return set_error(SWGN_ERR_INVALID, "bad arguments");
  for (int i = 0; i < n; ++i) {
    if (!trackers[i] || !epochs[i] || epochs[i]->n_obs < 0 || epochs[i]->n_obs > SWGN_MAXOBS || (epochs[i]->n_obs && !epochs[i]->obs))
      return set_error(SWGN_ERR_INVALID, "bad tracker / epoch");
    if (std::memcmp(&trackers[i]->cfg, &trackers[0]->cfg, sizeof(swgn_gnss_config)) != 0)
      return set_error(SWGN_ERR_INVALID, "all trackers of one call must share one configuration");
    for (int k = 0; k < epochs[i]->n_obs; ++k)
      if (epochs[i]->obs[k].sat >= SWGN_MAXSAT || epochs[i]->obs[k].sys > 2)
        return set_error(SWGN_ERR_INVALID, "observation with sat >= MAXSATNUM or sys > 2");
    for (int j = 0; j < i; ++j)
      if (trackers[j] == trackers[i]) return set_error(SWGN_ERR_INVALID, "one tracker may appear once per call");
  }
  const swgn_gnss_config cfg = trackers[0]->cfg;
  const bool dbg = std::getenv("SWGN_GNSS_DEBUG") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[gnss] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  if (cudaSetDevice(cfg.device) != cudaSuccess) return set_error(SWGN_ERR_NO_DEVICE, "no usable CUDA device (there is no CPU fallback)");
  swgn_status st = SWGN_OK;

  // ---- phase A: SPP correction and ambiguity lookup (swf_gnss.cpp:271-325), gating records -------------------
  std::vector<int64_t> obs0(n + 1, 0);
  for (int i = 0; i < n; ++i) obs0[i + 1] = obs0[i] + epochs[i]->n_obs;
  const int64_t n_obs_all = obs0[n];
  std::vector<double> rec((size_t)16 * std::max<int64_t>(n_obs_all, 1), 0.0), gate((size_t)3 * std::max<int64_t>(n_obs_all, 1), 0.0);
  std::vector<int32_t> flags((size_t)std::max<int64_t>(n_obs_all, 1), 0);
  parallel_for(n, [&](int i) {
    swgn_gnss_tracker& T = *trackers[i];
    swgn_epoch& E = *epochs[i];
    const swgn_gnss_frame& F = frames[i];
    if (cfg.use_spp_correction) {
      for (int k = 0; k < E.n_obs; ++k) {
        swgn_obs& d = E.obs[k];
        if (d.spp_p[0] != 0) {
          d.spp_p0[0] = d.spp_p[0];
          const int h = T.back(SWGN_AMB_PCORR, d.sat * 2 + 0);
          if (h >= 0) {
            swgn_ambiguity& a = T.amb[SWGN_AMB_PCORR][h];
            a.last_update_time = E.ros_time;
            if (a.continue_count > cfg.estimate_pcorrection_period) {
              d.spp_p0[0] = 0;
              d.spp_p[0] += a.value * cfg.lams[d.sys][0];
            }
          }
        } else {
          d.spp_p0[0] = 0;
        }
      }
    }
    for (int k = 0; k < E.n_obs; ++k) {
      swgn_obs& d = E.obs[k];
      for (int q = 0; q < SWGN_NFREQ; ++q) d.rtk_n[q] = d.spp_n[q] = d.pcorr_n[q] = -1;
      if (d.svh) continue;
      for (int q = 0; q < SWGN_NFREQ; ++q) {
        auto recent = [&](int fam) {
          const int h = T.back(fam, d.sat * 2 + q);
          return (h >= 0 && E.ros_time - T.amb[fam][h].last_update_time < cfg.ambiguity_timeout) ? h : -1;
        };
        if (d.rtk_l[q] != 0) d.rtk_n[q] = recent(SWGN_AMB_RTK);
        if (d.spp_l[q] != 0) d.spp_n[q] = recent(SWGN_AMB_SPP);
        if (d.spp_p0[q] != 0) d.pcorr_n[q] = recent(SWGN_AMB_PCORR);
      }
      // gating record of frequency 0 (the reference asserts RTK_L[1] == SPP_L[1] == 0, swf_gnss.cpp:397)
      double* r = &rec[(size_t)16 * (obs0[i] + k)];
      for (int c = 0; c < 3; ++c) {
        r[c] = d.sat_pos[c];
        r[3 + c] = F.pose[c] + E.base_xyz[c];
      }
      const double lam = cfg.lams[d.sys][0];
      r[9] = lam;
      int fl = 0;
      if (d.rtk_n[0] >= 0) {
        fl |= 1;
        r[10] = d.rtk_l[0];
        r[11] = T.amb[SWGN_AMB_RTK][d.rtk_n[0]].value;
        r[12] = F.gnss_dt[d.sys * 2 + 0];
      }
      if (d.spp_n[0] >= 0) {
        fl |= 2;
        r[13] = d.spp_l[0];
        r[14] = T.amb[SWGN_AMB_SPP][d.spp_n[0]].value;
        r[15] = F.gnss_dt[6 + d.sys * 2 + 0];
      }
      flags[obs0[i] + k] = fl;
    }
  });

  lap("phase A (host)");
  // ---- device: elevations + gating residuals of every observation of the call ---------------------------------
  {
    double *d_rec = nullptr, *d_out = nullptr;
    int32_t* d_flags = nullptr;
    if (n_obs_all > 0) {
      // stream-ordered allocations: no device-wide synchronisation, the pool keeps the blocks for the next call
      cudaStream_t s = cudaStreamPerThread;
      swgn::keep_pool_memory(cfg.device);
      cudaError_t e = cudaMallocAsync((void**)&d_rec, sizeof(double) * 16 * n_obs_all, s);
      if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_out, sizeof(double) * 3 * n_obs_all, s);
      if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_flags, sizeof(int32_t) * n_obs_all, s);
      lap("  gate: alloc");
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_rec, rec.data(), sizeof(double) * 16 * n_obs_all, cudaMemcpyHostToDevice, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_flags, flags.data(), sizeof(int32_t) * n_obs_all, cudaMemcpyHostToDevice, s);
      lap("  gate: h2d");
      if (e == cudaSuccess) e = swgn::launch_gate_residuals((int)n_obs_all, d_rec, d_flags, cfg.azelmin, d_out, s);
      if (dbg) cudaStreamSynchronize(s);
      lap("  gate: kernel");
      if (e == cudaSuccess) e = cudaMemcpyAsync(gate.data(), d_out, sizeof(double) * 3 * n_obs_all, cudaMemcpyDeviceToHost, s);
      if (d_rec) cudaFreeAsync(d_rec, s);
      if (d_out) cudaFreeAsync(d_out, s);
      if (d_flags) cudaFreeAsync(d_flags, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) return set_error(SWGN_ERR_CUDA, std::string("gating residuals: ") + cudaGetErrorString(e));
    }
  }
  lap("gate residuals (device)");
  // ---- phase B: medians, slip conditions, new ambiguities, counters (swf_gnss.cpp:346-500) ---------------------
  parallel_for(n, [&](int i) {
    swgn_gnss_tracker& T = *trackers[i];
    swgn_epoch& E = *epochs[i];
    const swgn_gnss_frame& F = frames[i];
    swgn_gnss_output& O = outputs[i];
    O.n_new[0] = O.n_new[1] = O.n_new[2] = O.n_slip_rtk = O.n_slip_spp = 0;
    std::vector<double> err2_rtk[6], err2_spp[6];
    double med_rtk[6] = {0}, med_spp[6] = {0};
    for (int k = 0; k < E.n_obs; ++k) {
      swgn_obs& d = E.obs[k];
      if (d.svh) continue;
      const double* gk = &gate[(size_t)3 * (obs0[i] + k)];
      d.el = gk[0];
      for (int q = 0; q < SWGN_NFREQ; ++q) {
        if (d.el < cfg.azelmin) d.rtk_l[q] = d.spp_l[q] = d.spp_p0[q] = 0;
        // the second frequency carries no phase in this estimator (:397); an ambiguity found for it is gated as untouched
        if (q == 0 && d.rtk_n[q] >= 0 && T.amb[SWGN_AMB_RTK][d.rtk_n[q]].slip_count == d.rtk_slip_count[q])
          err2_rtk[d.sys * 2 + q].push_back(gk[1]);
        if (q == 0 && d.spp_n[q] >= 0 && T.amb[SWGN_AMB_SPP][d.spp_n[q]].slip_count == d.spp_slip_count[q])
          err2_spp[d.sys * 2 + q].push_back(gk[2]);
      }
    }
    for (int s = 0; s < 6; ++s) {
      if (!err2_rtk[s].empty()) med_rtk[s] = median_of(err2_rtk[s]);
      if (!err2_spp[s].empty()) med_spp[s] = median_of(err2_spp[s]);
    }
    const bool reset_all = F.not_fix_count > cfg.phase_all_reset_count;
    for (int k = 0; k < E.n_obs; ++k) {
      swgn_obs& d = E.obs[k];
      if (d.svh) continue;
      const double* gk = &gate[(size_t)3 * (obs0[i] + k)];
      for (int q = 0; q < SWGN_NFREQ; ++q) {
        const double lam = cfg.lams[d.sys][q];
        bool condition3 = false, condition4 = false;
        if (d.rtk_l[q] != 0 && cfg.use_imu && cfg.use_rtk && F.nonlinear && F.rover_count > 1 && d.rtk_n[q] >= 0 &&
            T.amb[SWGN_AMB_RTK][d.rtk_n[q]].slip_count == d.rtk_slip_count[q]) {
          if (std::fabs(gk[1] - med_rtk[d.sys * 2 + q]) > lam * cfg.slip_fraction_rtk) {
            condition3 = true;
            O.n_slip_rtk++;
          }
        }
        if (d.spp_l[q] != 0 && cfg.use_imu && cfg.use_spp_phase && F.nonlinear && F.rover_count > 1 && d.spp_n[q] >= 0 &&
            T.amb[SWGN_AMB_SPP][d.spp_n[q]].slip_count == d.spp_slip_count[q]) {
          const double v = T.amb[SWGN_AMB_SPP][d.spp_n[q]].value;
          if (std::fabs((d.spp_l[q] + v) * lam - d.spp_p[q]) * std::sin(d.el) * std::sin(d.el) > 10) condition4 = true;
          if (std::fabs(gk[2] - med_spp[d.sys * 2 + q]) > lam) condition4 = true;
          if (condition4) O.n_slip_spp++;
        }
        if (d.rtk_l[q] != 0) {
          if (d.rtk_n[q] < 0 || T.amb[SWGN_AMB_RTK][d.rtk_n[q]].slip_count != d.rtk_slip_count[q] || condition3 || reset_all) {
            d.rtk_n[q] = T.push(SWGN_AMB_RTK, d.sat, d.sys, q);
            T.amb[SWGN_AMB_RTK][d.rtk_n[q]].slip_count = d.rtk_slip_count[q];
            T.amb[SWGN_AMB_RTK][d.rtk_n[q]].half_flag = d.half_flag[q];
            O.n_new[SWGN_AMB_RTK]++;
          }
          T.amb[SWGN_AMB_RTK][d.rtk_n[q]].last_update_time = E.ros_time;
        }
        if (d.spp_l[q] != 0) {
          if (d.spp_n[q] < 0 || T.amb[SWGN_AMB_SPP][d.spp_n[q]].slip_count != d.spp_slip_count[q] || condition3 || condition4) {
            d.spp_n[q] = T.push(SWGN_AMB_SPP, d.sat, d.sys, q);
            T.amb[SWGN_AMB_SPP][d.spp_n[q]].slip_count = d.spp_slip_count[q];
            T.amb[SWGN_AMB_SPP][d.spp_n[q]].half_flag = d.half_flag[q];
            O.n_new[SWGN_AMB_SPP]++;
          }
          T.amb[SWGN_AMB_SPP][d.spp_n[q]].last_update_time = E.ros_time;
        }
        if (d.spp_p0[q] != 0) {
          if (d.pcorr_n[q] < 0) {
            d.pcorr_n[q] = T.push(SWGN_AMB_PCORR, d.sat, d.sys, q);
            O.n_new[SWGN_AMB_PCORR]++;
          }
          T.amb[SWGN_AMB_PCORR][d.pcorr_n[q]].last_update_time = E.ros_time;
        }
        if (d.rtk_n[q] >= 0) T.amb[SWGN_AMB_RTK][d.rtk_n[q]].continue_count++;
        if (d.spp_n[q] >= 0) T.amb[SWGN_AMB_SPP][d.spp_n[q]].continue_count++;
        if (d.pcorr_n[q] >= 0) T.amb[SWGN_AMB_PCORR][d.pcorr_n[q]].continue_count++;
      }
    }
  });

  lap("phase B (host)");
  // ---- pack (AddGnssResidual) ------------------------------------------------------------------------------
  std::vector<EpochGraph> G(n);
  std::vector<const swgn_graph*> gp;  // epochs with at least one GNSS factor
  std::vector<int> gi;
  std::vector<char> too_small(n, 0);
  parallel_for(n, [&](int i) {
    G[i].build(*trackers[i], *epochs[i], frames[i]);
    swgn_gnss_output& O = outputs[i];
    O.n_factors = (int32_t)G[i].kind.size() + 1;
    O.n_keep = 1 + (G[i].B.b_pose >= 0) + (G[i].B.b_sb >= 0) + (int32_t)G[i].B.amb_handle.size();
    O.n = G[i].n_keep_tangent;
    std::memset(&O.init_summary, 0, sizeof(O.init_summary));
    if (O.n_keep > O.cap_keep || O.n > O.cap_n || !O.keep_kind || !O.keep_handle || !O.keep_idx || !O.x0 || !O.J0 || !O.r0) {
      too_small[i] = 1;
      return;
    }
    int kb = 0, col = 0, xo = 0;
    auto keep = [&](int kind, int handle, int tangent, const double* x, int nx) {
      O.keep_kind[kb] = kind;
      O.keep_handle[kb] = handle;
      O.keep_idx[kb] = col;
      ++kb;
      col += tangent;
      for (int c = 0; c < nx; ++c) O.x0[xo++] = x ? x[c] : 0.0;
    };
    if (G[i].B.b_pose >= 0) keep(SWGN_KEEP_POSE, -1, 6, frames[i].pose, 7);
    if (G[i].B.b_sb >= 0) keep(SWGN_KEEP_SPEED_BIAS, -1, 9, frames[i].speed_bias, 9);
    keep(SWGN_KEEP_BLACK, -1, 1, &frames[i].blackvalue, 1);
    for (size_t a = 0; a < G[i].B.amb_handle.size(); ++a)
      keep(SWGN_KEEP_AMB_RTK + G[i].B.amb_family[a], G[i].B.amb_handle[a], 1, nullptr, 1);
    if (G[i].kind.empty()) {
      // no usable observation: the epoch's prior is the InitialBlackFactor alone (1 x 1: J0 = istd, r0 = istd * x)
      O.J0[0] = G[i].unit_istd;
      O.r0[0] = G[i].unit_istd * frames[i].blackvalue;
    }
  });
  for (int i = 0; i < n; ++i) {
    if (too_small[i]) return set_error(SWGN_ERR_INVALID, "output buffers too small for the epoch's keep blocks");
    if (!G[i].kind.empty()) gi.push_back(i);
  }
  if (gi.empty()) return SWGN_OK;
  lap("pack");

  swgn_options opt;
  swgn_default_options(&opt);
  opt.device = cfg.device;
  swgn_batch* batch = nullptr;
  std::vector<swgn_summary> sums(gi.size());
  std::vector<double> xs;

  // ---- pass 1: linearise at (states, ambiguities = 0), eliminate the clock terms, square root (:504-530) --------
  for (int i : gi) {
    G[i].fill_state(*trackers[i], frames[i], true);
    G[i].point();
    gp.push_back(&G[i].g);
  }
  opt.is_optimize = 0;
  opt.max_num_iterations = 1;
  opt.n_parameter_head = 1;  // the export switch of the modified Ceres needs a head: the one group holding every keep block
  st = swgn_batch_create(&opt, (int32_t)gp.size(), gp.data(), &batch);
  if (st != SWGN_OK) return st;
  lap("pass 1 create");
  st = swgn_batch_solve(batch, sums.data());
  lap("pass 1 solve");
  if (st == SWGN_OK) {  // all priors in two launches; every (J0, r0) is scattered from one pinned staging block into its caller buffer
    std::vector<int32_t> n_tail(gi.size());
    std::vector<double*> Jp(gi.size()), rp(gi.size());
    for (size_t k = 0; k < gi.size(); ++k) {
      const swgn_gnss_output& O = outputs[gi[k]];
      if (sums[k].n_f != O.n) {
        st = set_error(SWGN_ERR_INVALID, "internal: reduced system size differs from the keep blocks");
        break;
      }
      n_tail[k] = O.n;
      Jp[k] = O.J0;
      rp[k] = O.r0;
    }
    if (st == SWGN_OK) st = swgn::batch_marginal_priors_to(batch, n_tail.data(), Jp.data(), rp.data());
  }
  lap("pass 1 priors");
  swgn_batch_destroy(batch);
  batch = nullptr;
  if (st != SWGN_OK) return st;

  // ---- pass 2: initialise new ambiguities / clocks, Ceres' default strategy (:532-571) ---------------------------
  if (cfg.use_spp_phase || cfg.use_rtk) {
    for (int i : gi) {
      G[i].fill_state(*trackers[i], frames[i], false);
      if (G[i].B.b_pose >= 0) G[i].konst[G[i].B.b_pose] = 1;
      if (G[i].B.b_sb >= 0) G[i].konst[G[i].B.b_sb] = 1;
      for (size_t a = 0; a < G[i].B.amb_handle.size(); ++a)
        if (trackers[i]->amb[G[i].B.amb_family[a]][G[i].B.amb_handle[a]].continue_count > cfg.init_constant_after)
          G[i].konst[G[i].B.b_amb0 + (int)a] = 1;
    }
    swgn_default_options(&opt);
    opt.device = cfg.device;
    opt.trust_region_strategy = SWGN_LEVENBERG_MARQUARDT;
    opt.jacobi_scaling = 1;
    opt.initial_trust_region_radius = opt.max_trust_region_radius = cfg.init_radius;
    opt.max_num_iterations = cfg.init_max_iterations;
    opt.n_parameter_head = 0;
    st = swgn_batch_create(&opt, (int32_t)gp.size(), gp.data(), &batch);
    if (st != SWGN_OK) return st;
    lap("pass 2 create");
    st = swgn_batch_solve(batch, sums.data());
    lap("pass 2 solve");
    if (st == SWGN_OK) {
      xs.resize((size_t)swgn_batch_states_size(batch));
      st = swgn_batch_get_states(batch, xs.data());  // the states of all epochs back to back, one copy
    }
    size_t x0 = 0;
    for (size_t k = 0; st == SWGN_OK && k < gi.size(); ++k) {
      const int i = gi[k];
      const double* x = xs.data() + x0;
      x0 += G[i].state.size();
      outputs[i].init_summary = sums[k];
      frames[i].blackvalue = x[G[i].offset[G[i].B.b_black]];
      for (int s : G[i].B.clk_slots) frames[i].gnss_dt[s] = x[G[i].offset[G[i].B.clk_block[s]]];
      for (size_t a = 0; a < G[i].B.amb_handle.size(); ++a)
        trackers[i]->amb[G[i].B.amb_family[a]][G[i].B.amb_handle[a]].value = x[G[i].offset[G[i].B.b_amb0 + (int)a]];
    }
    swgn_batch_destroy(batch);
    lap("pass 2 read-back");
  }
  return st;
}

}  // extern "C"
