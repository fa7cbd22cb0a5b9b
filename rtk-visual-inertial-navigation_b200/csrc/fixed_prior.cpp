// swgn_fixed_integer_prior (include/swgn.h): the prior rebuild after FIX_CONTINUE_THRESHOLD accepted ambiguity fixes,
// RVI/swf/swf_lambda.cpp:249-355.  Host side only: every job becomes a small factor graph -- the old prior as a dense
// linear factor, FixedIntegerFactors between per-system/frequency dummy scalars (elimination group 0) and the
// ambiguities -- and all jobs go through ONE export-mode pass of the batched solver and ONE batched eigen square root
// (swgn_batch_get_marginal_priors).  No numerical work happens on the host.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/swgn.h"

namespace swgn {
swgn_status set_error(swgn_status st, const std::string& m);
}
using swgn::set_error;

namespace {
struct JobGraph {
  std::vector<int32_t> size, manifold, konst, group, offset;
  std::vector<double> state;
  std::vector<int32_t> gkind, gblocks;
  std::vector<double> gdata;
  int32_t prior_n, prior_blk_begin[2];
  std::vector<int32_t> prior_blocks, prior_blk_idx;
  int64_t zero64 = 0;
  swgn_graph g;
};
}  // namespace

extern "C" swgn_status swgn_fixed_integer_prior(int32_t device, int32_t n_jobs, const swgn_fixed_integer_job* jobs) {
  if (n_jobs <= 0 || !jobs) return set_error(SWGN_ERR_INVALID, "bad arguments");
  std::vector<JobGraph> G(n_jobs);
  std::vector<const swgn_graph*> gp(n_jobs);
  for (int j = 0; j < n_jobs; ++j) {
    const swgn_fixed_integer_job& J = jobs[j];
    if (J.n_keep <= 0 || J.n <= 0 || !J.keep_size || !J.keep_idx || !J.x0 || !J.J0 || !J.r0 || !J.x || J.n_dd < 0 ||
        (J.n_dd > 0 && (!J.dd_keep || !J.F || !J.dd_sysfreq)) || !J.J0_out || !J.r0_out || !(J.istd > 0))
      return set_error(SWGN_ERR_INVALID, "job " + std::to_string(j) + ": bad arguments");
    JobGraph& Q = G[j];
    // blocks: the keep blocks in prior order, then one dummy per system / frequency in use (tf[6], :247)
    int off = 0, tangent = 0;
    for (int k = 0; k < J.n_keep; ++k) {
      const int s = J.keep_size[k];
      if (s <= 0) return set_error(SWGN_ERR_INVALID, "job " + std::to_string(j) + ": bad keep block size");
      Q.size.push_back(s);
      Q.manifold.push_back(s == 7 ? SWGN_MANIFOLD_POSE : SWGN_MANIFOLD_EUCLIDEAN);
      Q.konst.push_back(0);
      Q.group.push_back(1);
      Q.offset.push_back(off);
      Q.prior_blocks.push_back(k);
      Q.prior_blk_idx.push_back(J.keep_idx[k]);
      off += s;
      tangent += s == 7 ? 6 : s;
    }
    if (tangent != J.n) return set_error(SWGN_ERR_INVALID, "job " + std::to_string(j) + ": keep blocks do not add up to n");
    Q.state.assign(J.x, J.x + off);
    int dummy_of[6] = {-1, -1, -1, -1, -1, -1};
    auto fixed_integer = [&](int dummy, int amb_block, double n21) {
      Q.gkind.push_back(SWGN_GNSS_FIXED_INTEGER);
      Q.gblocks.push_back(dummy);  // FixedIntegerFactor <1;1,1>: (N_ref, N_a), r = istd ((N_a - N_ref) - N21)
      Q.gblocks.push_back(amb_block);
      Q.gblocks.push_back(-1);
      double rec[SWGN_GNSS_STRIDE] = {0};
      rec[SWGN_GNSS_MEAS] = n21;
      rec[SWGN_GNSS_WEIGHT] = J.istd;
      Q.gdata.insert(Q.gdata.end(), rec, rec + SWGN_GNSS_STRIDE);
    };
    for (int d = 0; d < J.n_dd; ++d) {
      const int sf = J.dd_sysfreq[d], kp = J.dd_keep[2 * d], kn = J.dd_keep[2 * d + 1];
      if (sf < 0 || sf > 5 || kp < 0 || kp >= J.n_keep || kn < 0 || kn >= J.n_keep || J.keep_size[kp] != 1 || J.keep_size[kn] != 1)
        return set_error(SWGN_ERR_INVALID, "job " + std::to_string(j) + ": bad double difference");
      if (dummy_of[sf] < 0) {  // first double difference of this system / frequency: tie the dummy to the reference ambiguity
        dummy_of[sf] = (int)Q.size.size();
        Q.size.push_back(1);
        Q.manifold.push_back(SWGN_MANIFOLD_EUCLIDEAN);
        Q.konst.push_back(0);
        Q.group.push_back(0);
        Q.offset.push_back((int32_t)Q.state.size());
        Q.state.push_back(0.0);
        fixed_integer(dummy_of[sf], kn, 0.0);
      }
      fixed_integer(dummy_of[sf], kp, J.F[d]);
    }
    std::memset(&Q.g, 0, sizeof(Q.g));
    Q.g.n_blocks = (int32_t)Q.size.size();
    Q.g.block_size = Q.size.data();
    Q.g.block_manifold = Q.manifold.data();
    Q.g.block_const = Q.konst.data();
    Q.g.block_group = Q.group.data();
    Q.g.block_offset = Q.offset.data();
    Q.g.n_state = (int32_t)Q.state.size();
    Q.g.state = Q.state.data();
    Q.g.proj_sqrt_info[0] = Q.g.proj_sqrt_info[3] = 1.0;
    Q.g.n_gnss = (int32_t)Q.gkind.size();
    Q.g.gnss_kind = Q.gkind.data();
    Q.g.gnss_blocks = Q.gblocks.data();
    Q.g.gnss_data = Q.gdata.data();
    Q.prior_n = J.n;
    Q.prior_blk_begin[0] = 0;
    Q.prior_blk_begin[1] = J.n_keep;
    Q.g.n_prior = 1;
    Q.g.prior_n = &Q.prior_n;
    Q.g.prior_blk_begin = Q.prior_blk_begin;
    Q.g.prior_blocks = Q.prior_blocks.data();
    Q.g.prior_blk_idx = Q.prior_blk_idx.data();
    Q.g.prior_x0_begin = &Q.zero64;
    Q.g.prior_x0 = J.x0;
    Q.g.prior_J_begin = &Q.zero64;
    Q.g.prior_J = J.J0;
    Q.g.prior_r_begin = &Q.zero64;
    Q.g.prior_r0 = J.r0;
    gp[j] = &Q.g;
  }
  // jobs without a fixed double difference have no elimination group: their prior is re-linearised on its own, which the
  // solver cannot express as a Schur pass -- the reference never calls the rebuild without a fix either (:249)
  for (int j = 0; j < n_jobs; ++j)
    if (jobs[j].n_dd == 0) return set_error(SWGN_ERR_INVALID, "job " + std::to_string(j) + ": no fixed double difference");
  swgn_options opt;
  swgn_default_options(&opt);
  opt.device = device;
  opt.is_optimize = 0;
  opt.max_num_iterations = 1;
  opt.n_parameter_head = 1;  // the one group holding every keep block
  swgn_batch* batch = nullptr;
  swgn_status st = swgn_batch_create(&opt, n_jobs, gp.data(), &batch);
  if (st != SWGN_OK) return st;
  std::vector<swgn_summary> sums(n_jobs);
  st = swgn_batch_solve(batch, sums.data());
  if (st == SWGN_OK) {
    std::vector<int32_t> n_tail(n_jobs);
    std::vector<int64_t> j_off(n_jobs), r_off(n_jobs);
    int64_t nj = 0, nr = 0;
    for (int j = 0; j < n_jobs; ++j) {
      if (sums[j].n_f != jobs[j].n) {
        st = set_error(SWGN_ERR_INVALID, "internal: reduced system size differs from the prior's");
        break;
      }
      n_tail[j] = jobs[j].n;
      j_off[j] = nj;
      r_off[j] = nr;
      nj += (int64_t)jobs[j].n * jobs[j].n;
      nr += jobs[j].n;
    }
    std::vector<double> Jall((size_t)nj), rall((size_t)nr);
    if (st == SWGN_OK) st = swgn_batch_get_marginal_priors(batch, n_tail.data(), j_off.data(), r_off.data(), Jall.data(), rall.data());
    for (int j = 0; st == SWGN_OK && j < n_jobs; ++j) {
      std::memcpy(jobs[j].J0_out, Jall.data() + j_off[j], sizeof(double) * (size_t)jobs[j].n * jobs[j].n);
      std::memcpy(jobs[j].r0_out, rall.data() + r_off[j], sizeof(double) * jobs[j].n);
    }
  }
  swgn_batch_destroy(batch);
  return st;
}
