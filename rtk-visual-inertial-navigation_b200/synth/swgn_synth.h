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
} swgn_synth_config;

typedef struct swgn_synth swgn_synth;

/* cfg 1: 5 KF x 50 LM, VI only; cfg 2: 20 KF x 300 LM x 10 epochs x 20 sats */
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

#ifdef __cplusplus
}
#endif
#endif
