// The reference-side binding of the per-epoch GNSS preprocessing (include/swgn_gnss.h): what a maintainer adds to
// SWFOptimization::GnssProcess (RVI/swf/swf_gnss.cpp:175-262) so that GnssPreprocess (:265-587) runs through the library.
// Templates over the reference's own wire structs (ObsMea / mea_t, RVI/gnss/include/common_function.h:72-124), so the file
// compiles wherever that header is present and costs nothing elsewhere; field by field, no layout assumption.
//
//   mea_to_epoch(*rover, &epoch, obs_buffer);                       // before the call
//   swgn_gnss_preprocess(1, &tracker, &epoch_ptr, &frame, &output); // frame = para_pose / para_speed_bias / para_gnss_dt[0] / blackvalue
//   epoch_to_mea(epoch, rover);                                     // el, masked measurements, corrected SPP_P
//   // output.J0 / r0 / x0 / keep_* become the MarginalizationFactor over (pose, speed-bias, blackvalue, ambiguities);
//   // ambiguity values live in the tracker (swgn_gnss_tracker_get / _set_value) instead of the PBtype lists.
#ifndef SWGN_REFERENCE_GNSS_BINDING_H_
#define SWGN_REFERENCE_GNSS_BINDING_H_
#include "swgn_gnss.h"

namespace swgn_binding {
template <class MeaT>
void mea_to_epoch(const MeaT& m, swgn_epoch* e, swgn_obs* obs /* MAXOBS entries */) {
  e->n_obs = m.obs_count;
  e->pad_ = 0;
  e->ros_time = m.ros_time;
  e->br_time_diff = m.br_time_diff;
  for (int c = 0; c < 3; ++c) e->base_xyz[c] = m.base_xyz[c];
  e->obs = obs;
  for (int i = 0; i < m.obs_count; ++i) {
    const auto& d = m.obs_data[i];
    swgn_obs& o = obs[i];
    o = swgn_obs();
    o.sat = d.sat;
    o.sys = d.sys;
    o.svh = d.SVH;
    for (int f = 0; f < SWGN_NFREQ; ++f) {
      o.rtk_slip_count[f] = d.RTK_SLIP_COUNT[f];
      o.spp_slip_count[f] = d.SPP_SLIP_COUNT[f];
      o.half_flag[f] = d.half_flag[f];
      o.spp_p[f] = d.SPP_P[f];
      o.spp_l[f] = d.SPP_L[f];
      o.spp_d[f] = d.SPP_D[f];
      o.spp_lstd[f] = d.SPP_Lstd[f];
      o.spp_pstd[f] = d.SPP_Pstd[f];
      o.spp_dstd[f] = d.SPP_Dstd[f];
      o.rtk_p[f] = d.RTK_P[f];
      o.rtk_l[f] = d.RTK_L[f];
      o.rtk_pstd[f] = d.RTK_Pstd[f];
      o.rtk_lstd[f] = d.RTK_Lstd[f];
      o.spp_p0[f] = d.SPP_P0[f];
      o.rtk_n[f] = o.spp_n[f] = o.pcorr_n[f] = -1;
    }
    for (int c = 0; c < 3; ++c) {
      o.sat_pos[c] = d.satellite_pos[c];
      o.sat_vel[c] = d.satellite_vel[c];
    }
    o.el = d.el;
    o.sat_var = d.sat_var;
    o.ion_var = d.ion_var;
    o.trop_var = d.trop_var;
  }
}
// what GnssPreprocess changes inside the mea_t besides the PBtype pointers
template <class MeaT>
void epoch_to_mea(const swgn_epoch& e, MeaT* m) {
  for (int i = 0; i < e.n_obs; ++i) {
    auto& d = m->obs_data[i];
    const swgn_obs& o = e.obs[i];
    d.el = o.el;
    for (int f = 0; f < SWGN_NFREQ; ++f) {
      d.RTK_L[f] = o.rtk_l[f];
      d.SPP_L[f] = o.spp_l[f];
      d.SPP_P[f] = o.spp_p[f];
      d.SPP_P0[f] = o.spp_p0[f];
    }
  }
}
}  // namespace swgn_binding
#endif
