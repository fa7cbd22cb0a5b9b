// swgn_marginalize (include/swgn.h): MarginalizationInfo::marginalize + getParameterBlocks for arbitrary drop sets
// (RVI/factor/marginalization_factor.cpp:260-400), as MargFrames uses it for MargImagSecondNew / MargRoverOld
// (RVI/swf/swf.cpp:329-341).  Host side only.  The solver's Schur stage wants a non-empty elimination group of mutually
// independent blocks, which an arbitrary drop set is not; every graph therefore gets one private scalar block with a unit
// factor as its group 0 (it couples to nothing, so eliminating it changes nothing), the drop blocks become group 1 and the
// keep blocks group 2 = the head.  The export pass leaves S over (drop | keep); swgn_batch_get_marginal_priors reduces the
// leading drop rows with the reference's eigen pseudo-inverse and takes the eigen square root -- all on the device.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/swgn.h"

namespace swgn {
swgn_status set_error(swgn_status st, const std::string& m);
swgn_status batch_marginal_priors_to(swgn_batch* b, const int32_t* n_tail, double* const* J0_ptr, double* const* r0_ptr);
}
using swgn::set_error;

namespace {
struct Aug {
  std::vector<int32_t> size, manifold, konst, group, offset, unit_block;
  std::vector<double> state, unit_istd;
  swgn_graph g;
};
}  // namespace

extern "C" swgn_status swgn_marginalize(int32_t device, int32_t n_graphs, const swgn_graph* const* graphs, const uint8_t* const* drop,
                                        swgn_marginalize_output* outputs) {
  if (n_graphs <= 0 || !graphs || !drop || !outputs) return set_error(SWGN_ERR_INVALID, "bad arguments");
  std::vector<Aug> A(n_graphs);
  std::vector<const swgn_graph*> gp(n_graphs);
  for (int w = 0; w < n_graphs; ++w) {
    const swgn_graph* g = graphs[w];
    const std::string tag = "graph " + std::to_string(w) + ": ";
    if (!g || !drop[w] || g->n_blocks <= 0) return set_error(SWGN_ERR_INVALID, tag + "bad arguments");
    if (g->n_order > 0 || g->order || g->is_use || g->n_chain > 0 || g->n_host > 0)
      return set_error(SWGN_ERR_UNSUPPORTED, tag + "order / is_use / chains / host-evaluated factors are not supported here");
    Aug& Q = A[w];
    swgn_marginalize_output& O = outputs[w];
    const int nb = g->n_blocks;
    Q.size.assign(g->block_size, g->block_size + nb);
    Q.manifold.assign(g->block_manifold, g->block_manifold + nb);
    Q.konst.assign(g->block_const, g->block_const + nb);
    Q.group.resize(nb);
    Q.offset.assign(g->block_offset, g->block_offset + nb);
    Q.state.assign(g->state, g->state + g->n_state);
    // blocks no factor touches take no part (the reference only ever sees the blocks of the residual blocks it was given)
    std::vector<char> touched(nb, 0);
    auto touch = [&](const int32_t* blocks, int64_t count) {
      for (int64_t k = 0; k < count; ++k)
        if (blocks[k] >= 0 && blocks[k] < nb) touched[blocks[k]] = 1;
    };
    touch(g->proj_blocks, (int64_t)3 * g->n_proj);
    touch(g->imu_blocks, (int64_t)4 * g->n_imu);
    touch(g->gnss_blocks, (int64_t)3 * g->n_gnss);
    if (g->n_prior > 0) touch(g->prior_blocks, g->prior_blk_begin[g->n_prior]);
    touch(g->unit_block, g->n_unit);
    int n = 0, m = 0, nk = 0;
    for (int b = 0; b < nb; ++b) {
      const int t = g->block_manifold[b] == SWGN_MANIFOLD_POSE ? 6 : g->block_size[b];
      Q.group[b] = drop[w][b] ? 1 : 2;
      if (g->block_const[b] || !touched[b]) continue;  // constant: stays a constant of the factors; untouched: not in the problem
      if (drop[w][b]) {
        m += t;
      } else {
        if (nk >= O.cap_keep || !O.keep_block || !O.keep_idx) return set_error(SWGN_ERR_INVALID, tag + "output buffers too small");
        O.keep_block[nk] = b;
        O.keep_idx[nk] = n;
        ++nk;
        n += t;
      }
    }
    if (nk == 0) return set_error(SWGN_ERR_INVALID, tag + "nothing to keep");
    if (n > O.cap_n || !O.J0 || !O.r0) return set_error(SWGN_ERR_INVALID, tag + "output buffers too small");
    O.n_keep = nk;
    O.n = n;
    O.m = m;
    // the private group-0 block and its unit factor
    Q.size.push_back(1);
    Q.manifold.push_back(SWGN_MANIFOLD_EUCLIDEAN);
    Q.konst.push_back(0);
    Q.group.push_back(0);
    Q.offset.push_back((int32_t)Q.state.size());
    Q.state.push_back(0.0);
    Q.unit_block.assign(g->unit_block, g->unit_block + g->n_unit);
    Q.unit_istd.assign(g->unit_istd, g->unit_istd + g->n_unit);
    Q.unit_block.push_back(nb);
    Q.unit_istd.push_back(1.0);
    Q.g = *g;
    Q.g.n_blocks = nb + 1;
    Q.g.block_size = Q.size.data();
    Q.g.block_manifold = Q.manifold.data();
    Q.g.block_const = Q.konst.data();
    Q.g.block_group = Q.group.data();
    Q.g.block_offset = Q.offset.data();
    Q.g.n_state = (int32_t)Q.state.size();
    Q.g.state = Q.state.data();
    Q.g.n_unit = g->n_unit + 1;
    Q.g.unit_block = Q.unit_block.data();
    Q.g.unit_istd = Q.unit_istd.data();
    gp[w] = &Q.g;
  }
  swgn_options opt;
  swgn_default_options(&opt);
  opt.device = device;
  opt.is_optimize = 0;
  opt.max_num_iterations = 1;
  opt.n_parameter_head = 1;  // group 2
  swgn_batch* batch = nullptr;
  swgn_status st = swgn_batch_create(&opt, n_graphs, gp.data(), &batch);
  if (st != SWGN_OK) return st;
  std::vector<swgn_summary> sums(n_graphs);
  st = swgn_batch_solve(batch, sums.data());
  if (st == SWGN_OK) {
    std::vector<int32_t> n_tail(n_graphs);
    std::vector<double*> Jp(n_graphs), rp(n_graphs);
    for (int w = 0; w < n_graphs; ++w) {
      if (sums[w].n_f != outputs[w].n + outputs[w].m) {
        st = set_error(SWGN_ERR_INVALID, "graph " + std::to_string(w) + ": internal: reduced system size differs from the drop and keep blocks");
        break;
      }
      n_tail[w] = outputs[w].n;
      Jp[w] = outputs[w].J0;
      rp[w] = outputs[w].r0;
    }
    if (st == SWGN_OK) st = swgn::batch_marginal_priors_to(batch, n_tail.data(), Jp.data(), rp.data());
  }
  swgn_batch_destroy(batch);
  return st;
}
