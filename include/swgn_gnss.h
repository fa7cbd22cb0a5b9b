/*
 * swgn_gnss.h -- per-epoch GNSS linearisation in front of the sliding-window solve.
 *
 * What it replaces, RVI/ = /root/reference/rtk_visual_inertial_src/rtk_visual_inertial/src/:
 *   SWFOptimization::GnssPreprocess       RVI/swf/swf_gnss.cpp:265-587
 *   SWFOptimization::AddGnssResidual      RVI/swf/swf_core.cpp:87-205
 *   update_azel / ecef2pos / satazel      RVI/gnss/src/common_function.cpp:91-124,394-408
 *   MarginalizationInfo::marginalize      RVI/factor/marginalization_factor.cpp:260-377 (as called at swf_gnss.cpp:527-530)
 *   wire structs ObsMea / mea_t / PBtype  RVI/gnss/include/common_function.h:47-124 (RVI/main3.cpp:151-172 memcpy's a
 *                                         mea_t out of a std_msgs/ByteMultiArray)
 *
 * One call handles the newest GNSS epoch of MANY receivers (one tracker per receiver): elevations and the cycle-slip
 * gating residuals of every observation in one launch, the ambiguity bookkeeping on the host, then the flattened
 * GNSS-only graphs of all epochs go through the batched solver twice -- an export-mode pass whose reduced system over
 * (pose, speed-bias, blackvalue, ambiguities) with the receiver-clock terms eliminated becomes the epoch's
 * MarginalizationFactor (marg_info_gnss), and the 2-iteration LEVENBERG_MARQUARDT + jacobi_scaling pass that
 * initialises new ambiguities and the clock terms (swf_gnss.cpp:532-571).
 */
#ifndef SWGN_GNSS_H_
#define SWGN_GNSS_H_
#include "swgn.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { SWGN_NFREQ = 2, SWGN_MAXOBS = 64, SWGN_MAXSAT = 107, SWGN_GNSS_NCLK = 13 };

/* ObsMea (common_function.h:72-112) without its three PBtype* arrays: those become the *_n handles below, filled by
   swgn_gnss_preprocess (-1 = nullptr).  Same field meaning and units. */
typedef struct swgn_obs {
  uint8_t sat, sys, svh, pad0_;
  uint8_t rtk_slip_count[SWGN_NFREQ], spp_slip_count[SWGN_NFREQ], half_flag[SWGN_NFREQ], pad1_[6];
  double spp_p[SWGN_NFREQ], spp_l[SWGN_NFREQ], spp_d[SWGN_NFREQ];
  double spp_lstd[SWGN_NFREQ], spp_pstd[SWGN_NFREQ], spp_dstd[SWGN_NFREQ];
  double rtk_p[SWGN_NFREQ], rtk_l[SWGN_NFREQ], rtk_pstd[SWGN_NFREQ], rtk_lstd[SWGN_NFREQ];
  double spp_p0[SWGN_NFREQ];
  double sat_pos[3], sat_vel[3];
  double el;                         /* in/out: overwritten by update_azel */
  double sat_var, ion_var, trop_var;
  int32_t rtk_n[SWGN_NFREQ];         /* RTK_Npoint               */
  int32_t spp_n[SWGN_NFREQ];         /* SPP_Npoint               */
  int32_t pcorr_n[SWGN_NFREQ];       /* SPP_Npoint_PCottections  */
} swgn_obs;

/* mea_t (common_function.h:115-123); marg_info_gnss / residualBlockId are outputs of the call instead */
typedef struct swgn_epoch {
  int32_t n_obs, pad_;
  double ros_time;
  double base_xyz[3];
  double br_time_diff;
  swgn_obs* obs; /* n_obs entries, modified in place like the reference modifies its mea_t */
} swgn_epoch;

/* the switches and constants GnssPreprocess / AddGnssResidual read (RVI/parameter/parameters.h, common_function.h) */
typedef struct swgn_gnss_config {
  int32_t use_imu, use_rtk, use_rtd, use_spp_phase, use_spp_correction, use_doppler;
  int32_t phase_all_reset_count;        /* Phase_ALL_RESET_COUNT (yaml)          */
  int32_t estimate_pcorrection_period;  /* EstimatePcorrectionPerio = 500        */
  double azelmin;                       /* AZELMIN = 25 deg                      */
  double lams[3][SWGN_NFREQ];           /* wavelengths, common_function.cpp:4-8  */
  double ambiguity_timeout;             /* 10 s: an ambiguity not observed for longer is not reused (swf_gnss.cpp:310) */
  double slip_fraction_rtk;             /* 0.5: |residual - median| > lam * this resets an RTK ambiguity (:399)      */
  int32_t init_max_iterations;          /* 2  (swf_gnss.cpp:567)                 */
  int32_t init_constant_after;          /* 10 (continue_count above which an ambiguity is held constant, :546-556) */
  double init_radius;                   /* 1e15 (:566)                           */
  int32_t device;
  int32_t pad_;
} swgn_gnss_config;
void swgn_gnss_config_default(swgn_gnss_config* c);

/* what GnssPreprocess reads from the estimator besides the epoch */
typedef struct swgn_gnss_frame {
  double pose[7];                    /* para_pose[g2f[rover_count-1]]            */
  double speed_bias[9];              /* para_speed_bias[...]                     */
  double gnss_dt[SWGN_GNSS_NCLK];    /* para_gnss_dt[0]: 6 RB-SD clocks, 6 rover-only clocks, Doppler drift; in/out */
  double blackvalue;                 /* in/out                                   */
  int32_t nonlinear;                 /* solver_flag == NonLinear                 */
  int32_t rover_count;               /* GNSS epochs in the window, this one included */
  int32_t epochs_since_start;        /* rover_count_accumulate - rover_count + ir: below 100 the SPP pseudorange weight is x10 */
  int32_t not_fix_count;
} swgn_gnss_frame;

/* PBtype (common_function.h:47-69) */
typedef struct swgn_ambiguity {
  double value;
  double last_update_time;
  int32_t continue_count;
  uint8_t slip_count, half_flag, sys, f;
  int32_t sat;
  int32_t alive; /* 0 after swgn_gnss_tracker_erase */
} swgn_ambiguity;
enum { SWGN_AMB_RTK = 0, SWGN_AMB_SPP = 1, SWGN_AMB_PCORR = 2 };

/* keep blocks of the epoch's prior, in this order: pose, speed-bias, blackvalue, then every ambiguity the epoch's
   observations point at (RTK, SPP, pseudorange-correction; observation order).  The reference's order is the
   iteration order of a std::unordered_map keyed by address (marginalization_factor.h:76-80,
   marginalization_factor.cpp:264-277); only the column permutation of J0 depends on it. */
enum { SWGN_KEEP_POSE = 0, SWGN_KEEP_SPEED_BIAS = 1, SWGN_KEEP_BLACK = 2, SWGN_KEEP_AMB_RTK = 3, SWGN_KEEP_AMB_SPP = 4,
       SWGN_KEEP_AMB_PCORR = 5 };
typedef struct swgn_gnss_output {
  /* capacities of the caller's buffers */
  int32_t cap_keep, cap_n;
  /* marg_info_gnss: r = r0 + J0 (x [-] x0) over the keep blocks */
  int32_t n_keep, n;
  int32_t* keep_kind;   /* SWGN_KEEP_*                                          */
  int32_t* keep_handle; /* ambiguity handle, -1 for the first three             */
  int32_t* keep_idx;    /* first tangent column in J0                           */
  double* x0;           /* cap_n + 3 doubles: global sizes, keep order (ambiguities at 0: PhaseBiasSaveAndReset) */
  double* J0;           /* n x n row-major                                      */
  double* r0;           /* n                                                    */
  /* bookkeeping diagnostics */
  int32_t n_new[3];     /* ambiguities created this epoch per family            */
  int32_t n_slip_rtk;   /* condition3 hits                                      */
  int32_t n_slip_spp;   /* condition4 hits                                      */
  int32_t n_factors;    /* residual blocks of the epoch's GNSS graph (with InitialBlackFactor) */
  swgn_summary init_summary; /* the phase-bias initialisation solve             */
} swgn_gnss_output;

typedef struct swgn_gnss_tracker swgn_gnss_tracker; /* the three std::list<PBtype>[MAXSATNUM*2] of one receiver, swf.h:274-278 */
swgn_status swgn_gnss_tracker_create(const swgn_gnss_config* cfg, swgn_gnss_tracker** out);
void swgn_gnss_tracker_destroy(swgn_gnss_tracker* t);
int32_t swgn_gnss_tracker_count(const swgn_gnss_tracker* t, int32_t family);
swgn_status swgn_gnss_tracker_get(const swgn_gnss_tracker* t, int32_t family, int32_t handle, swgn_ambiguity* out);
swgn_status swgn_gnss_tracker_set_value(swgn_gnss_tracker* t, int32_t family, int32_t handle, double value);
/* the reference erases list entries when their last epoch leaves the window (swf.cpp:372-380) */
swgn_status swgn_gnss_tracker_erase(swgn_gnss_tracker* t, int32_t family, int32_t handle);

/* GnssPreprocess for n (tracker, epoch, frame) triples at once.  All trackers must share one configuration.
   Side effects as in the reference: epoch observations (el, masked measurements, SPP correction, handles), tracker
   lists (new / counted ambiguities, initialised values), frames[i].gnss_dt / blackvalue (initialisation solve). */
swgn_status swgn_gnss_preprocess(int32_t n, swgn_gnss_tracker* const* trackers, swgn_epoch* const* epochs,
                                 swgn_gnss_frame* frames, swgn_gnss_output* outputs);

/* IMUGNSSBase::AddMargInfo (RVI/factor/gnss_imu_factor.cpp:245-352): the epoch's prior becomes one hidden frame of an
   IMUGNSSFactor chain (the chain_* fields of swgn_graph, include/swgn.h).  Host-side scatter of the prior's information
   A = J0'J0, b = J0'r0: the pose (6) and speed-bias (9) blocks fill the frame's 15 x 15 pose_hessians and pose_rhses, their
   coupling with the size-1 keep blocks fills pose_phase_biases_hessians (15 x k), and the size-1 blocks ACCUMULATE into the
   chain's phase_biases_hessians / phase_biases_rhs (k x k followed by k).  keep_slot[i] names, for every keep block of the
   prior, its slot 0..k-1 in the chain's phase-bias list (the reference treats EVERY size-1 keep block as one, blackvalue
   included) or -1 for the pose / speed-bias blocks.  frame: SWGN_CHAIN_FRAME_STRIDE doubles (current hidden states = pose /
   speed_bias, linearisation point = the prior's x0); frame_N: 15 x k row-major, overwritten; chain_N: k x k + k, added to. */
swgn_status swgn_gnss_chain_frame(const swgn_gnss_output* prior, const double* pose, const double* speed_bias, int32_t k,
                                  const int32_t* keep_slot, double* frame, double* frame_N, double* chain_N);

/* ---- staged entry points (parity tests) ----------------------------------------------------------------------- */
/* update_azel + the gating residuals of swf_gnss.cpp:346-377 for n_obs observations in one launch:
   rec = 16 doubles per observation {sat_pos[3], receiver ECEF[3], base[3] (unused), lam, L_rtk, N_rtk, clk_rtk, L_spp,
   N_spp, clk_spp}; out = 3 doubles per observation {el, residual_rtk, residual_spp}. */
swgn_status swgn_gnss_gate_residuals(int32_t n_obs, const double* rec, double* out, int32_t device);
/* AddGnssResidual (swf_core.cpp:87-205) for one preprocessed epoch: the SWGN_GNSS_* records.  blocks are indices into
   the epoch's block list: 0 pose, 1 speed-bias, 2 blackvalue, 3 .. 3+n_clk-1 the clock slots actually used
   (clk_slot[]), then the keep ambiguities in keep order.  Buffers sized for 5 * n_obs factors; the InitialBlackFactor
   is not in the list (it is the graph's unit factor). */
swgn_status swgn_gnss_epoch_records(const swgn_gnss_tracker* t, const swgn_epoch* e, const swgn_gnss_frame* f,
                                    int32_t* n_factors, int32_t* kind, int32_t* blocks, double* data, int32_t* n_clk,
                                    int32_t* clk_slot, int32_t* n_amb, int32_t* amb_family, int32_t* amb_handle);

#ifdef __cplusplus
}
#endif
#endif
