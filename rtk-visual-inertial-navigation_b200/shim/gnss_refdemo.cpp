// Drop-in evidence for include/swgn_gnss.h: the per-epoch preprocessing reached through the REFERENCE'S OWN wire struct.
// Built by oracle/build_ref.sh into oracle/_ref/libswgn_refdemo.so together with ceres_shim_refdemo.cpp (only where
// /root/reference is present); test infrastructure around the product, nothing in the product links it.
#include <cstdint>
#include <cstring>

#include "swgn_gnss.h"

// ---- per-epoch GNSS preprocessing through the reference's own wire struct ------------------------------------------
// The epoch arrives as the flat swgn_epoch the tests generate, is written into a REAL mea_t (the struct RVI/main3.cpp
// memcpy's out of the ROS message), and from there takes the path a maintainer's GnssProcess would take:
// mea_to_epoch -> swgn_gnss_preprocess -> epoch_to_mea.  Returns the library's status; `mea_bytes` reports sizeof(mea_t).
#include "common_function.h"
#include "reference_gnss_binding.h"

extern "C" int swgn_refdemo_gnss_preprocess(swgn_gnss_tracker* tracker, swgn_epoch* epoch, swgn_gnss_frame* frame, swgn_gnss_output* out,
                                            int* mea_bytes) {
  static mea_t rover;  // ~50 KB
  if (mea_bytes) *mea_bytes = (int)sizeof(mea_t);
  if (epoch->n_obs > MAXOBS) return -1;
  rover.obs_count = epoch->n_obs;
  rover.ros_time = epoch->ros_time;
  rover.br_time_diff = epoch->br_time_diff;
  for (int c = 0; c < 3; ++c) rover.base_xyz[c] = epoch->base_xyz[c];
  for (int i = 0; i < epoch->n_obs; ++i) {
    ObsMea& d = rover.obs_data[i];
    const swgn_obs& o = epoch->obs[i];
    d.sat = o.sat;
    d.sys = o.sys;
    d.SVH = o.svh;
    for (int f = 0; f < NFREQ; ++f) {
      d.RTK_SLIP_COUNT[f] = o.rtk_slip_count[f];
      d.SPP_SLIP_COUNT[f] = o.spp_slip_count[f];
      d.half_flag[f] = o.half_flag[f];
      d.SPP_P[f] = o.spp_p[f];
      d.SPP_L[f] = o.spp_l[f];
      d.SPP_D[f] = o.spp_d[f];
      d.SPP_Lstd[f] = o.spp_lstd[f];
      d.SPP_Pstd[f] = o.spp_pstd[f];
      d.SPP_Dstd[f] = o.spp_dstd[f];
      d.RTK_P[f] = o.rtk_p[f];
      d.RTK_L[f] = o.rtk_l[f];
      d.RTK_Pstd[f] = o.rtk_pstd[f];
      d.RTK_Lstd[f] = o.rtk_lstd[f];
      d.SPP_P0[f] = o.spp_p0[f];
      d.RTK_Npoint[f] = d.SPP_Npoint[f] = d.SPP_Npoint_PCottections[f] = nullptr;
    }
    for (int c = 0; c < 3; ++c) {
      d.satellite_pos[c] = o.sat_pos[c];
      d.satellite_vel[c] = o.sat_vel[c];
    }
    d.el = o.el;
    d.sat_var = o.sat_var;
    d.ion_var = o.ion_var;
    d.trop_var = o.trop_var;
  }
  // the maintainer-side code proper
  static swgn_obs obs[MAXOBS];
  swgn_epoch e;
  swgn_binding::mea_to_epoch(rover, &e, obs);
  swgn_epoch* ep = &e;
  const swgn_status st = swgn_gnss_preprocess(1, &tracker, &ep, frame, out);
  if (st != SWGN_OK) return (int)st;
  swgn_binding::epoch_to_mea(e, &rover);
  // hand the results back to the caller's flat epoch for comparison
  for (int i = 0; i < epoch->n_obs; ++i) {
    epoch->obs[i] = obs[i];
    epoch->obs[i].el = rover.obs_data[i].el;
    epoch->obs[i].rtk_l[0] = rover.obs_data[i].RTK_L[0];
  }
  return 0;
}
