// TEST INFRASTRUCTURE.  The reference's own MarginalizationInfo (RVI/factor/marginalization_factor.cpp, compiled where it
// lies against oracle/ref_stubs' Eigen stand-in) fed with the reference's own GNSS factor classes: the pin of the oracle's
// restated marginalize() (oracle/oracle_gnss_epoch.cpp) and, through it, of the per-epoch prior the device produces.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "../include/swgn.h"
#include "factor/gnss_factor.h"
#include "factor/initial_factor.h"
#include "factor/marginalization_factor.h"

extern "C" {
// Factors: kind[i] = SWGN_GNSS_* (record = SWGN_GNSS_STRIDE doubles; RTK kinds are built with el / dt / var and
// use_istd = true) or -1 = InitialBlackFactor(record[0]); blocks[3 * i ..] index the block table (-1 pad).  Blocks:
// size, drop flag, offset into state.  Runs addResidualBlockInfo for every factor, marginalize(true, true),
// getParameterBlocks() and returns n, m, the keep blocks in the reference's order (block index, first column) and the
// linearised factor.  J is n x n row-major.
int ref_marginalize(int n_factors, const int32_t* kind, const int32_t* blocks, const double* records, int n_blocks,
                    const int32_t* block_size, const int32_t* block_drop, const int32_t* block_offset, double* state, int32_t* n_out,
                    int32_t* m_out, int32_t* n_keep, int32_t* keep_block, int32_t* keep_idx, double* J, double* r) {
  MarginalizationInfo* info = new MarginalizationInfo();
  double xyzt[3] = {0, 0, 0};
  // the factor classes keep the POINTERS they are constructed with (in the reference they point into the mea_t): the
  // arrays must outlive marginalize()
  std::vector<double> geo((size_t)9 * n_factors);
  for (int i = 0; i < n_factors; ++i) {
    const double* rec = records + (size_t)SWGN_GNSS_STRIDE * i;
    double *sat = &geo[(size_t)9 * i], *satv = sat + 3, *base = sat + 6;
    for (int c = 0; c < 3; ++c) {
      sat[c] = rec[SWGN_GNSS_SAT_POS + c];
      satv[c] = rec[SWGN_GNSS_SAT_VEL + c];
      base[c] = rec[SWGN_GNSS_BASE_POS + c];
    }
    const double meas = rec[SWGN_GNSS_MEAS], lam = rec[SWGN_GNSS_LAM], wgt = rec[SWGN_GNSS_WEIGHT];
    const double el = rec[SWGN_GNSS_EL], dt = rec[SWGN_GNSS_DT], var = rec[SWGN_GNSS_VAR];
    ceres::CostFunction* f = nullptr;
    switch (kind[i]) {
      case -1: f = new InitialBlackFactor(rec[0]); break;
      case SWGN_GNSS_SPP_PSEUDORANGE: f = new SppPseudorangeFactor(sat, meas, wgt, base); break;
      case SWGN_GNSS_SPP_CARRIER: f = new SppCarrierPhaseFactor(sat, meas, wgt, base, lam); break;
      case SWGN_GNSS_RTK_CARRIER: f = new RTKCarrierPhaseFactor(sat, meas, lam, el, dt, var, base, true, 0, 0); break;
      case SWGN_GNSS_RTK_PSEUDORANGE: f = new RTKPseudorangeFactor(sat, meas, el, dt, var, base); break;
      case SWGN_GNSS_DOPPLER: f = new SppDopplerFactor(satv, sat, xyzt, meas, wgt, base); break;
      default: return 2;
    }
    std::vector<double*> pb;
    std::vector<int> drop;
    for (int k = 0; k < 3; ++k) {
      const int b = blocks[3 * i + k];
      if (b < 0) continue;
      if (block_drop[b]) drop.push_back((int)pb.size());
      pb.push_back(state + block_offset[b]);
    }
    if ((int)pb.size() != (int)f->parameter_block_sizes().size()) return 3;
    info->addResidualBlockInfo(new ResidualBlockInfo(f, nullptr, pb, drop, std::vector<int>{}));
  }
  info->marginalize(true, true);
  std::vector<double*> keep = info->getParameterBlocks();
  *n_out = info->n;
  *m_out = info->m;
  *n_keep = (int)keep.size();
  std::map<double*, int> block_of;
  for (int b = 0; b < n_blocks; ++b) block_of[state + block_offset[b]] = b;
  for (size_t k = 0; k < keep.size(); ++k) {
    keep_block[k] = block_of[keep[k]];
    keep_idx[k] = info->keep_block_idx[k] - info->m;
    if (info->keep_block_size[k] != block_size[keep_block[k]]) return 4;
  }
  const int n = info->n;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) J[(size_t)i * n + j] = info->linearized_jacobians(i, j);
    r[i] = info->linearized_residuals(i);
  }
  delete info;  // owns the ResidualBlockInfo objects and their cost functions (marginalization_factor.cpp:46-56)
  return 0;
}

// RVI/swf/swf_lambda.cpp:249-355 with the reference's own classes: last_marg_info filled through its public members,
// MarginalizationFactor(last_marg_info), FixedIntegerFactor(0 | F, istd) against the tf dummies, marginalize(false, true).
// Arguments as swgn_fixed_integer_job (include/swgn.h); the keep blocks come back in the reference's own order
// (keep_order[k] = index into the job's keep list, keep_col[k] = first column).
int ref_fixed_integer_prior(int n_keep, int n, const int32_t* keep_size, const int32_t* keep_idx, const double* x0, const double* J0,
                            const double* r0, const double* x, int n_dd, const int32_t* dd_keep, const double* F,
                            const int32_t* dd_sysfreq, double istd, int32_t* keep_order, int32_t* keep_col, double* J_out, double* r_out) {
  std::vector<int> xoff;
  int nx = 0;
  for (int k = 0; k < n_keep; ++k) {
    xoff.push_back(nx);
    nx += keep_size[k];
  }
  std::vector<double> store(x, x + nx), lin(x0, x0 + nx);
  double tf[6] = {0};
  bool tfb[6] = {false};
  MarginalizationInfo* last_marg_info = new MarginalizationInfo();
  last_marg_info->n = n;
  last_marg_info->m = 0;
  for (int k = 0; k < n_keep; ++k) {
    last_marg_info->keep_block_size.push_back(keep_size[k]);
    last_marg_info->keep_block_idx.push_back(keep_idx[k]);
    last_marg_info->keep_block_data.push_back(lin.data() + xoff[k]);
    last_marg_info->keep_block_addr.push_back(store.data() + xoff[k]);
  }
  last_marg_info->linearized_jacobians.resize(n, n);
  last_marg_info->linearized_residuals = Eigen::VectorXd(n);
  for (int i = 0; i < n; ++i) {
    last_marg_info->linearized_residuals(i) = r0[i];
    for (int j = 0; j < n; ++j) last_marg_info->linearized_jacobians(i, j) = J0[(size_t)i * n + j];
  }
  MarginalizationInfo* marginalization_info2 = new MarginalizationInfo();
  MarginalizationFactor* factormarge = new MarginalizationFactor(last_marg_info);
  marginalization_info2->addResidualBlockInfo(
      new ResidualBlockInfo(factormarge, NULL, last_marg_info->keep_block_addr, std::vector<int>{}, std::vector<int>{}));
  for (int i = 0; i < n_dd; i++) {
    double* ppoint = store.data() + xoff[dd_keep[2 * i]];
    double* npoint = store.data() + xoff[dd_keep[2 * i + 1]];
    const int sf = dd_sysfreq[i];
    if (tfb[sf] == false) {
      marginalization_info2->addResidualBlockInfo(new ResidualBlockInfo(new FixedIntegerFactor(0, istd), NULL,
                                                                        std::vector<double*>{&tf[sf], npoint}, std::vector<int>{0}, std::vector<int>{}));
      tfb[sf] = true;
    }
    marginalization_info2->addResidualBlockInfo(new ResidualBlockInfo(new FixedIntegerFactor(F[i], istd), NULL,
                                                                      std::vector<double*>{&tf[sf], ppoint}, std::vector<int>{0}, std::vector<int>{}));
  }
  marginalization_info2->marginalize(false, true);
  std::vector<double*> keep = marginalization_info2->getParameterBlocks();
  if (marginalization_info2->n != n || (int)keep.size() != n_keep) return 2;
  for (size_t k = 0; k < keep.size(); ++k) {
    int which = -1;
    for (int q = 0; q < n_keep; ++q)
      if (keep[k] == store.data() + xoff[q]) which = q;
    if (which < 0) return 3;
    keep_order[k] = which;
    keep_col[k] = marginalization_info2->keep_block_idx[k] - marginalization_info2->m;
  }
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) J_out[(size_t)i * n + j] = marginalization_info2->linearized_jacobians(i, j);
    r_out[i] = marginalization_info2->linearized_residuals(i);
  }
  // ~MarginalizationInfo deletes parameter_block_data it allocated and, when m and n are non-zero, its factors; the old
  // prior's keep_block_data point into `lin` and were never registered in parameter_block_data
  delete marginalization_info2;
  delete last_marg_info;
  return 0;
}

// MarginalizationFactor::Evaluate (marginalization_factor.cpp:410-446) on a MarginalizationInfo filled through its public
// members: residuals (n) and the row-major n x global-size Jacobians back to back.
int ref_prior_eval(int n_keep, int n, const int32_t* keep_size, const int32_t* keep_idx, const double* x0, const double* J0,
                   const double* r0, const double* x, double* residuals, double* jac_out) {
  std::vector<int> xoff;
  int nx = 0;
  for (int k = 0; k < n_keep; ++k) {
    xoff.push_back(nx);
    nx += keep_size[k];
  }
  std::vector<double> lin(x0, x0 + nx);
  MarginalizationInfo info;
  info.n = n;
  info.m = 0;
  for (int k = 0; k < n_keep; ++k) {
    info.keep_block_size.push_back(keep_size[k]);
    info.keep_block_idx.push_back(keep_idx[k]);
    info.keep_block_data.push_back(lin.data() + xoff[k]);
  }
  info.linearized_jacobians.resize(n, n);
  info.linearized_residuals = Eigen::VectorXd(n);
  for (int i = 0; i < n; ++i) {
    info.linearized_residuals(i) = r0[i];
    for (int j = 0; j < n; ++j) info.linearized_jacobians(i, j) = J0[(size_t)i * n + j];
  }
  MarginalizationFactor f(&info);
  std::vector<const double*> p;
  std::vector<double*> J;
  double* jp = jac_out;
  for (int k = 0; k < n_keep; ++k) {
    p.push_back(x + xoff[k]);
    J.push_back(jp);
    if (jp) jp += (size_t)n * keep_size[k];
  }
  return f.Evaluate(p.data(), residuals, jac_out ? J.data() : nullptr) ? 0 : 1;
}
}
