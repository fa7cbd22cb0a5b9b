/*
 * swgn_synth.h -- deterministic synthetic sliding windows.
 *
 * The reference ships no data generator and its data set is external, so benchmark and parity
 * inputs are synthesised: a figure-8 trajectory, 400 Hz IMU samples pre-integrated with the
 * reference's midpoint scheme (RVI/factor/integration_base.cpp:30-113), landmark tracks,
 * RB-SD carrier-phase / pseudorange / Doppler measurements, a dense prior, and the reference's
 * elimination ordering (RVI/swf/swf_gnss.cpp:629-783).  Output is a swgn_graph (include/swgn.h).
 * This library is workload tooling: it contains no solver code and is independent of oracle/.
 */
#ifndef SWGN_SYNTH_H_
#define SWGN_SYNTH_H_
#include <stdint.h>

#include "../../include/swgn.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swgn_synth_config {
  int32_t n_keyframes;     /* visual keyframes, 0.25 s apart                                  */
  int32_t n_landmarks;
  int32_t n_gnss_epochs;   /* one GNSS frame after every 2nd keyframe (+0.05 s); 0 = VI only  */
  int32_t n_sats;          /* split 8:7:5 over GPS/BDS/GAL (scaled)                           */
  uint64_t seed0;          /* window w is seeded with splitmix64(seed0 + w)                   */
  double state_noise;      /* scale of the initial-state perturbation (1.0 = SURVEY values)   */
  int32_t composition;     /* 0 = "B": GNSS epochs are explicit frames with raw GNSS factors;
                              1 = "A" (reference-faithful): the GNSS frames between
                              two keyframes are hidden inside one IMUGNSSFactor chain that carries
                              their pre-linearised GNSS information                            */
  int32_t hidden_per_gap;  /* composition A: GNSS frames per keyframe gap that has any (1..3)  */
  double bias_walk_scale;  /* multiplies the yaml acc_w / gyr_w bias random-walk densities (1.0 =
                              yaml).  With the yaml values the eliminated chain Hessian spans 16
                              decades (1e12 .. 1e-4 = its own rounding noise), and the reference's
                              absolute 1e-8 eigenvalue threshold (gnss_imu_factor.cpp:9,483) lets
                              noise eigenvalues through; the composition-A presets use 30 so that
                              parity tests compare arithmetic, not amplified rounding noise     */
  double hidden_bias_istd; /* composition A: 1/sigma of a weak absolute bias term in the hidden
                              frames' GNSS information (acc bias; gyro bias uses 10x); 0 = none */
  int32_t variant;         /* composition B only, bit flags for the factor kinds the default window does not hold:
                              1 = rover-only SppPseudorangeFactor / SppCarrierPhaseFactor (gnss_factor.cpp:9-80) in place
                                  of the RB-SD pair; 2 = a FixedIntegerFactor (:85-96) between every ambiguity and the
                                  first ambiguity of its constellation, as LambdaSearch injects them after a fix
                                  (swf_lambda.cpp:249-355); 4 = ESTIMATE_EXTRINSIC: the camera extrinsic is a free
                                  parameter block (ordered after the poses, swf_gnss.cpp:710-717) */
  int32_t pad_;
} swgn_synth_config;
enum { SWGN_SYNTH_SPP = 1, SWGN_SYNTH_FIXED_INTEGER = 2, SWGN_SYNTH_FREE_EXTRINSIC = 4 };

typedef struct swgn_synth swgn_synth;

/* cfg 1: 5 KF x 50 LM, VI only; cfg 2: 20 KF x 300 LM x 10 epochs x 20 sats;
   cfg 3: cfg 2 in composition A (hidden GNSS frames, 2 per gap); cfg 4: 6 KF x 40 LM x 2 gaps x 8 sats, A */
void swgn_synth_default_config(int32_t which, swgn_synth_config* c);
swgn_synth* swgn_synth_create(const swgn_synth_config* c, uint64_t window_id);
void swgn_synth_destroy(swgn_synth* s);
const swgn_graph* swgn_synth_graph(const swgn_synth* s);
/* ground-truth state in the graph's state layout */
const double* swgn_synth_truth(const swgn_synth* s);
/* options matching the window (n_parameter_head etc.) */
void swgn_synth_options(const swgn_synth* s, swgn_options* o);
/* structure numbers: [n_frames, n_obs (visual), n_imu, n_gnss, n_ambiguities, first_amb_block,
   n_prior_rows, n_blocks] */
void swgn_synth_info(const swgn_synth* s, int32_t* info8);
/* epoch -> observed ambiguity lists for swgn_ambiguity_fix: returns n_epochs; arrays may be
   NULL to query sizes (n_obs_total returned through *n_obs). */
int32_t swgn_synth_ambiguity_epochs(const swgn_synth* s, int32_t* epoch_begin, int32_t* obs_amb,
                                    int32_t* obs_sysfreq, int32_t* n_obs);
/* true integer ambiguities (n_ambiguities doubles) */
const double* swgn_synth_true_ambiguities(const swgn_synth* s);
/* ground truth of the hidden GNSS frames of the chain factors (16 doubles per frame, graph order);
   returns the number of hidden frames, frames16 may be NULL */
int32_t swgn_synth_chain_truth(const swgn_synth* s, double* frames16);
/* the (unreferenced) parameter blocks of the graph that hold the hidden frames' states in the
   application's memory layout: pose and speed-bias block index per hidden frame */
int32_t swgn_synth_chain_frame_blocks(const swgn_synth* s, int32_t* pose_block, int32_t* sb_block);

/* the generator's own IMU pre-integration on caller-provided samples (7 per sample: dt, acc, gyr; sample 0
   = initial acc/gyr): SWGN_IMU_STRIDE record; returns 0 on success.  Cross-check for the tests. */
int32_t swgn_synth_preintegrate(int32_t n_samples, const double* samples7, const double* bias6,
                                const double* noise4, double* record);

#ifdef __cplusplus
}
#endif
#endif
