/*
 * swgn.h -- C ABI of the Blackwell-native sliding-window Gauss-Newton solver.
 *
 * This is the drop-in boundary for the reference's hot path:
 * everything the reference does between `ceres::Solve(options, &my_problem, &summary)`
 * (RVI/swf/swf_image.cpp:219) and the read-backs `UpdateSchur` / `UpdateSchurHessianOnly`
 * (RVI/swf/swf_gnss.cpp:25-94) and `LambdaSearch` (RVI/swf/swf_lambda.cpp:82-245) is reachable
 * through the entry points below.  The header-compatible C++ shim in include/ceres/ sits on top
 * of this ABI; the CUDA implementation (libswgn.so) sits underneath.
 *
 * Conventions: plain C structs, pointers and sizes; the caller owns every host buffer it passes
 * in or receives results in; the library owns device memory behind the opaque handle; every
 * entry point returns a swgn_status (0 = ok) and never throws or aborts across the boundary.
 * There is no CPU implementation behind this ABI: if no CUDA device is usable the create call
 * fails with SWGN_ERR_NO_DEVICE.
 *
 * RVI/   = /root/reference/rtk_visual_inertial_src/rtk_visual_inertial/src/
 * CERES/ = ceres-solver-modified/ inside /root/reference/ceres-solver-modified.tar
 */
#ifndef SWGN_H_
#define SWGN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t swgn_status;
enum {
  SWGN_OK = 0,
  SWGN_ERR_INVALID = 1,      /* malformed graph / argument (Ceres would CHECK-fail) */
  SWGN_ERR_NO_DEVICE = 2,    /* no usable CUDA device: there is no CPU fallback */
  SWGN_ERR_CUDA = 3,         /* a CUDA runtime call failed; see swgn_last_error() */
  SWGN_ERR_ORDERING = 4,     /* group 0 of the ordering is not an independent set
                                (CERES/internal/ceres/program.cc:413-434) */
  SWGN_ERR_TOO_LARGE = 5,    /* reduced system >= 1024 rows with exports requested
                                (CERES/internal/ceres/schur_complement_solver.cc:178,256) */
  SWGN_ERR_UNSUPPORTED = 6
};

/* ---- parameter blocks ------------------------------------------------------------------ */
enum { SWGN_MANIFOLD_EUCLIDEAN = 0,
       /* 7 -> 6: p += dp ; q <- normalize(q * [1, dtheta/2]); layout (px,py,pz,qx,qy,qz,qw)
          RVI/factor/pose_local_parameterization.cpp:5-27 */
       SWGN_MANIFOLD_POSE = 1 };

/* ---- GNSS factor kinds (RVI/factor/gnss_factor.h) ---------------------------------------- */
enum {
  SWGN_GNSS_SPP_PSEUDORANGE = 0, /* <1;7,1>   (pose, clk)        gnss_factor.cpp:9-39    */
  SWGN_GNSS_SPP_CARRIER     = 1, /* <1;7,1,1> (pose, clk, N)     gnss_factor.cpp:45-80   */
  SWGN_GNSS_RTK_CARRIER     = 2, /* <1;7,1,1> (pose, N, clk)     gnss_factor.cpp:105-138 */
  SWGN_GNSS_RTK_PSEUDORANGE = 3, /* <1;7,1>   (pose, clk)        gnss_factor.cpp:140-168 */
  SWGN_GNSS_DOPPLER         = 4, /* <1;9,1,7> (sb, drift, pose)  gnss_factor.cpp:174-212 */
  SWGN_GNSS_FIXED_INTEGER   = 5  /* <1;1,1>   (N_ref, N_a)       gnss_factor.cpp:85-96   */
};
/* per-factor constant record, SWGN_GNSS_STRIDE doubles */
enum {
  SWGN_GNSS_SAT_POS = 0,   /* [3] satellite ECEF position                                   */
  SWGN_GNSS_SAT_VEL = 3,   /* [3] satellite ECEF velocity (Doppler only)                    */
  SWGN_GNSS_BASE_POS = 6,  /* [3] base_pos added to the pose translation                    */
  SWGN_GNSS_MEAS = 9,      /* P1 | L1_lam | D1_lam | N21                                     */
  SWGN_GNSS_LAM = 10,      /* wavelength (carrier types)                                     */
  SWGN_GNSS_WEIGHT = 11,   /* sqrt-information actually multiplied in: istd, or
                              1/sqrt(varerr2(el,dt,var)) evaluated on the host with the
                              reference's single-precision sinf (gnss_factor.cpp:98-103)    */
  SWGN_GNSS_EL = 12, SWGN_GNSS_DT = 13, SWGN_GNSS_VAR = 14, /* provenance of WEIGHT          */
  SWGN_GNSS_STRIDE = 16
};

/* per-IMU-factor constant record (fields of IntegrationBase, RVI/factor/integration_base.h) */
enum {
  SWGN_IMU_DELTA_P = 0,    /* [3]                                                            */
  SWGN_IMU_DELTA_Q = 3,    /* [4] x,y,z,w                                                    */
  SWGN_IMU_DELTA_V = 7,    /* [3]                                                            */
  SWGN_IMU_LIN_BA = 10,    /* [3] linearized_ba                                              */
  SWGN_IMU_LIN_BG = 13,    /* [3] linearized_bg                                              */
  SWGN_IMU_GYRI = 16,      /* [3]                                                            */
  SWGN_IMU_GYRJ = 19,      /* [3]                                                            */
  SWGN_IMU_SUM_DT = 22,
  SWGN_IMU_JACOBIAN = 24,  /* [225] 15x15 row-major d(delta)/d(bias) accumulated Jacobian    */
  SWGN_IMU_SQRT_INFO = 249,/* [225] 15x15 row-major, = LLT(cov^-1).L^T  (get_sqrtinfo)       */
  SWGN_IMU_STRIDE = 474
};

/* per-hidden-frame record of an IMUGNSSFactor chain (fields of IMUGNSSBase,
   RVI/factor/gnss_imu_factor.h:52-77), SWGN_CHAIN_FRAME_STRIDE doubles */
enum {
  SWGN_CHAIN_POSE = 0,       /* [7] gnss_poses[i]: current hidden pose (px,py,pz,qx,qy,qz,qw)  */
  SWGN_CHAIN_SB = 7,         /* [9] gnss_speed_bias[i]: current hidden (v, ba, bg)             */
  SWGN_CHAIN_POSE_LIN = 16,  /* [7] gnss_poses_lin[i]: linearisation point of the GNSS info    */
  SWGN_CHAIN_SB_LIN = 23,    /* [9] gnss_speed_bias_lin[i]                                     */
  SWGN_CHAIN_RHS = 32,       /* [15] pose_rhses[i]                                             */
  SWGN_CHAIN_HESSIAN = 48,   /* [225] pose_hessians[i], 15x15 row-major (symmetric)            */
  SWGN_CHAIN_FRAME_STRIDE = 274
};

/*
 * One sliding window as a flat factor graph: the content of the reference's long-lived
 * `ceres::Problem my_problem` plus `options.linear_solver_ordering` at the moment of Solve.
 * All arrays are caller-owned and only read.
 */
typedef int32_t (*swgn_host_eval_fn)(void* user, int32_t factor, double const* const* parameters, double* residuals,
                                     double** jacobians);

typedef struct swgn_graph {
  /* parameter blocks (identity in the reference = raw double*, here = index) */
  int32_t n_blocks;
  const int32_t* block_size;     /* global size                                              */
  const int32_t* block_manifold; /* SWGN_MANIFOLD_*                                          */
  const int32_t* block_const;    /* != 0: SetParameterBlockConstant                          */
  const int32_t* block_group;    /* ParameterBlockOrdering group id; 0 = eliminated e-blocks;
                                    ties inside a group are broken by block index (the
                                    reference breaks them by pointer value, ordered_groups.h) */
  const int32_t* block_offset;   /* first double of the block inside state[]                 */
  int32_t n_state;
  const double* state;           /* initial values, n_state doubles                          */

  /* application globals the factors read (RVI/parameter/parameters.h:88,94,100) */
  double Pbg[3];                 /* IMU -> GNSS antenna lever arm                            */
  double gravity[3];             /* Rwgw * G, gravity in the ECEF-aligned world frame        */
  double proj_sqrt_info[4];      /* projection_factor::sqrt_info, 2x2 row-major (swf.cpp:47) */
  double proj_cauchy_a;          /* CauchyLoss(a) on every projection factor; <= 0: no loss  */

  /* projection_factor <2;7,7,3>: (pose_j, cam extrinsic, world landmark) */
  int32_t n_proj;
  const int32_t* proj_blocks;    /* 3 per factor                                             */
  const double* proj_uv;         /* 2 per factor: pts.x, pts.y on the normalised plane       */

  /* IMUFactor <15;7,9,7,9>: (pose_i, sb_i, pose_j, sb_j) */
  int32_t n_imu;
  const int32_t* imu_blocks;     /* 4 per factor                                             */
  const double* imu_data;        /* SWGN_IMU_STRIDE per factor                               */

  /* GNSS scalar factors */
  int32_t n_gnss;
  const int32_t* gnss_kind;      /* SWGN_GNSS_*                                              */
  const int32_t* gnss_blocks;    /* 3 per factor in the factor's own parameter order, -1 pad */
  const double* gnss_data;       /* SWGN_GNSS_STRIDE per factor                              */

  /* MarginalizationFactor: r = r0 + J0 * (x [-] x0)  (marginalization_factor.cpp:410-446) */
  int32_t n_prior;
  const int32_t* prior_n;        /* rows (= columns) of J0 per prior                         */
  const int32_t* prior_blk_begin;/* n_prior+1 offsets into prior_blocks/prior_blk_idx        */
  const int32_t* prior_blocks;   /* keep blocks                                              */
  const int32_t* prior_blk_idx;  /* first tangent column of that block in J0                 */
  const int64_t* prior_x0_begin; /* n_prior offsets into prior_x0                            */
  const double* prior_x0;        /* linearisation point, global sizes, keep-block order      */
  const int64_t* prior_J_begin;  /* n_prior offsets into prior_J                             */
  const double* prior_J;         /* J0, n x n row-major                                      */
  const int64_t* prior_r_begin;  /* n_prior offsets into prior_r0                            */
  const double* prior_r0;

  /* InitialBlackFactor <1;1>: r = x * istd (initial_factor.cpp:90-96) */
  int32_t n_unit;
  const int32_t* unit_block;
  const double* unit_istd;

  /* residual-block program order (the order of AddResidualBlock calls).  Entry k encodes
     (kind << 28 | index) with kind 0 proj, 1 imu, 2 gnss, 3 prior, 4 unit, 5 chain, 6 host.  May be
     NULL: then the order is proj, imu, gnss, prior, unit, chain, host.  It only influences summation
     order.  */
  int32_t n_order;
  const uint32_t* order;

  /* ResidualBlock::is_use masks (CERES/internal/ceres/residual_block.h:135), one byte per
     factor in the same kind-major layout as above (proj, imu, gnss, prior, unit, chain); NULL =
     all used. */
  const uint8_t* is_use;

  /* IMUGNSSFactor <30+k; 7,9,7,9,1 x k> (RVI/factor/gnss_imu_factor.cpp:678-835): the m GNSS
     frames between two consecutive keyframes i and j are hidden inside one stateful factor.
     Every Jacobian evaluation re-eliminates the chain  kf_i -IMU- h_0 -IMU- ... h_{m-1} -IMU- kf_j
     (each hidden frame carries its pre-linearised GNSS information over (pose, speed-bias, N))
     frame by frame, and factors the resulting (30+k)^2 information matrix into J = sqrt(S) V^T;
     cost-only evaluations use the linearised residual r - J*INC; hidden states follow by
     back-substitution at the next Jacobian evaluation.
     Parameter order of chain c: chain_blocks[chain_blk_begin[c] ..] = pose_i, sb_i, pose_j, sb_j,
     then k = chain_blk_begin[c+1]-chain_blk_begin[c]-4 scalar phase-bias blocks (gnss_phase_biases).
     Hidden frames of chain c: [chain_frame_begin[c], chain_frame_begin[c+1]) in chain_frame_data
     (SWGN_CHAIN_FRAME_STRIDE each); chain_frame_N holds pose_phase_biases_hessians[i] (15 x k
     row-major) of every hidden frame back to back in the same order; chain_N holds per chain
     phase_biases_hessians (k x k row-major) followed by phase_biases_rhs (k); chain_imu_data
     holds m+1 SWGN_IMU_STRIDE records per chain: imu_factors[0..m-1] then last_imu_factor.
     The "middle marginalisation" link (pose1_pose2_hessians, gnss_imu_factor.cpp:741-760) is not
     represented. */
  int32_t n_chain;
  const int32_t* chain_blk_begin;    /* n_chain + 1                                           */
  const int32_t* chain_blocks;
  const int32_t* chain_frame_begin;  /* n_chain + 1                                           */
  const double* chain_frame_data;
  const double* chain_frame_N;
  const double* chain_N;
  const double* chain_imu_data;

  /* Host-evaluated residual blocks: cost functions the device has no factor kind for (the contract of
     CERES/include/ceres/cost_function.h:116 -- e.g. the reference's initialisation factors RVI/factor/initial_factor.cpp,
     mag_factor.cpp, pose0_factor.cpp).  Factor i has host_nres[i] residuals over the blocks host_blocks[host_blk_begin[i] ..
     host_blk_begin[i+1]); every evaluation of the window calls host_eval on the calling thread of swgn_batch_solve with
     the current values of those blocks (one pointer per block, global sizes) and receives residuals and row-major
     num_residuals x global-size Jacobians (jacobians may be NULL = residuals only), which are uploaded; pose blocks use
     the first 6 columns (the reference's parameterization Jacobian is [I6; 0]).  This costs two host round trips per
     trust-region iteration for the whole batch: meant for small initialisation problems, not for the timed workloads.
     No loss function.  Kind code 6 in `order` / is_use (after the chains). */
  int32_t n_host;
  const int32_t* host_nres;
  const int32_t* host_blk_begin;     /* n_host + 1 */
  const int32_t* host_blocks;
  swgn_host_eval_fn host_eval;       /* return 0 on success */
  void* host_user;
} swgn_graph;

/* ---- solver options: the subset of ceres::Solver::Options the reference sets, with the
   reference's effective defaults ------------------------------------ */
typedef struct swgn_options {
  int32_t max_num_iterations;            /* 8    yaml MAX_NUM_ITERATIONS                     */
  int32_t max_num_consecutive_invalid_steps; /* 5  solver.h:303                              */
  double initial_trust_region_radius;    /* 1e4  solver.h:277                                */
  double max_trust_region_radius;        /* 1e16 solver.h:278                                */
  double min_trust_region_radius;        /* 1e-32 solver.h:282                               */
  double min_relative_decrease;          /* 1e-3 solver.h:286                                */
  double min_lm_diagonal;                /* 1e-6 solver.h:295                                */
  double max_lm_diagonal;                /* 1e32 solver.h:296                                */
  double function_tolerance;             /* 1e-6 solver.h:309                                */
  double gradient_tolerance;             /* 1e-10 solver.h:316                               */
  double parameter_tolerance;            /* 1e-8 solver.h:322                                */
  double dogleg_min_mu;                  /* 1e-12: the reference's modified kMinMu
                                            (CERES/internal/ceres/dogleg_strategy.cc:51)     */
  int32_t is_optimize;                   /* ceres::internal::is_optimize; 0 = export mode:
                                            evaluate + eliminate once, export S and r,
                                            leave the state untouched
                                            (schur_complement_solver.cc:172-188)             */
  int32_t n_parameter_head;              /* ceres::internal::parameter_head.size(); the head
                                            blocks are the LAST n_parameter_head groups of the
                                            ordering (swf_gnss.cpp:775-782)                  */
  int32_t device;                        /* CUDA device ordinal                              */
  int32_t trust_region_strategy;         /* SWGN_DOGLEG (0, what the reference sets for the sliding-window solve,
                                            swf_image.cpp:205) or SWGN_LEVENBERG_MARQUARDT (Ceres' default,
                                            levenberg_marquardt_strategy.cc:66-165: what the reference's per-epoch
                                            GNSS solves run with, swf_gnss.cpp:204-215,563-573)                     */
  int32_t jacobi_scaling;                /* Solver::Options::jacobi_scaling (Ceres' default is true; the sliding-window
                                            solve sets false): columns of J scaled by 1 / (1 + |column|) of the initial
                                            Jacobian (trust_region_minimizer.cc:261-276,437).  Implemented with
                                            LEVENBERG_MARQUARDT; with DOGLEG swgn_batch_create answers SWGN_ERR_UNSUPPORTED */
  int32_t reserved_;
} swgn_options;
enum { SWGN_DOGLEG = 0, SWGN_LEVENBERG_MARQUARDT = 1 };

enum { SWGN_CONVERGENCE = 0, SWGN_NO_CONVERGENCE = 1, SWGN_FAILURE = 2 };

typedef struct swgn_summary {
  double initial_cost;
  double final_cost;
  double fixed_cost;
  int32_t num_successful_steps;   /* counts iteration 0 like Ceres does                      */
  int32_t num_unsuccessful_steps;
  int32_t num_iterations;         /* trust-region iterations executed (excluding iteration 0) */
  int32_t num_linear_solves;      /* Schur eliminate + Cholesky executions                   */
  int32_t termination_type;       /* SWGN_CONVERGENCE / NO_CONVERGENCE / FAILURE             */
  int32_t n_e;                    /* tangent dimension of eliminated blocks                  */
  int32_t n_f;                    /* rows of the reduced system (hs_row)                     */
  int32_t n_residuals;
} swgn_summary;

typedef struct swgn_batch swgn_batch;   /* opaque: n independent windows on one device */

void swgn_default_options(swgn_options* o);
const char* swgn_last_error(void);
const char* swgn_lambda_last_error(void);   /* CUDA error text of the last swgn_lambda_batch /
                                               swgn_ambiguity_fix failure on this thread */
const char* swgn_version(void);
int32_t swgn_device_count(void);

/* Preprocess (reduced program, ordering, chunk structure: CERES trust_region_preprocessor.cc
   :360-393) and upload n_windows graphs.  Windows are independent. */
swgn_status swgn_batch_create(const swgn_options* options, int32_t n_windows,
                              const swgn_graph* const* graphs, swgn_batch** out);
void swgn_batch_destroy(swgn_batch* b);   /* the batch's device / pinned slabs go to a process-wide cache (<= 4 GB device,
                                             <= 1 GB pinned) that later creates draw from; larger slabs are freed */
/* Host-only: run the preprocessing of one window without touching a device and report
   info[0..11] = n_cols, n_ecols, n_e, n_f, n_t, n_res, n_rows, n_chunks, n_jac, n_scells,
   n_sterms, n_srows; info[12..13] = algorithmic Schur bytes (low / high 32 bits); info[14] =
   tensor-core MMAs of one Schur gather pass.  Returns the
   same status codes swgn_batch_create would (SWGN_ERR_ORDERING, ...). */
swgn_status swgn_plan_probe(const swgn_graph* g, int32_t n_parameter_head, int32_t* info16);
/* Host-only: the planner's symbolic block fill-in of the reduced system's Cholesky factorisation -- per
   32-row panel a 64-bit mask of the 16-column groups that can be non-zero in the panel's rows of U (the
   device factorisation skips everything else).  n_panels may be queried with masks NULL. */
swgn_status swgn_plan_chol_masks(const swgn_graph* g, int32_t n_parameter_head, int32_t* n_panels,
                                 uint64_t* masks);
/* Host-only: decode the planner's per-warp gather streams of one window and verify their invariants (operand
   and store ranges, one end flag per tile); out[0..7] = stages, live terms, padding terms and tiles of the
   reduced-system streams, min / max stages per warp, stages and live terms of the e-cell streams. */
swgn_status swgn_plan_stream_check(const swgn_graph* g, int32_t n_parameter_head, int64_t* out8);
/* Host-only: the planner's ordering of one window -- column blocks of the reduced program in elimination order
   (graph block id per column; reorder_program.cc:209-245) and row blocks in Jacobian order (index of the residual
   block in program order per row; reorder_program.cc:247-326).  Either array may be NULL; *n_cols / *n_rows
   receive the counts (what swgn_batch_get_columns / _get_rows return from a device batch). */
swgn_status swgn_plan_order(const swgn_graph* g, int32_t n_parameter_head, int32_t* n_cols, int32_t* col_block,
                            int32_t* n_rows, int32_t* row_factor);
/* Host-only: the plan of the streamed Schur elimination of one window (k_schur_stream: the Jacobian staged batch by
   batch through shared memory by TMA, S accumulated on chip): info[0..9] = fits the on-chip budget (else the gather
   kernel runs the window), batches, accumulator doubles, J / residual / E-buffer / chunk-factor capacity of a stage
   (doubles), largest record package (ints), retained blocks, dynamic shared memory bytes; info[10] = the batch would use it (fits and SWGN_SCHUR_STREAM=1 is set:
   the streamed kernel is opt-in). */
swgn_status swgn_plan_stream_info(const swgn_graph* g, int32_t n_parameter_head, int32_t* info16);
/* Host-only (tests, debugging): copy one of the planner's index arrays (csrc/device_types.h IArr) of a window;
   *n receives its length, out may be NULL to query it. */
swgn_status swgn_plan_array(const swgn_graph* g, int32_t n_parameter_head, int32_t array, int32_t* out, int64_t* n);
int32_t swgn_batch_size(const swgn_batch* b);
/* Return the cached device / pinned slabs of destroyed batches to the driver (all devices); returns the bytes
   released.  Never needed for correctness: the cache is bounded (see swgn_batch_destroy). */
int64_t swgn_release_cached_memory(void);

/* Re-upload initial states only (structure unchanged): state_w has graphs[w]->n_state doubles. */
swgn_status swgn_batch_set_state(swgn_batch* b, int32_t window, const double* state);

/* Same structure, new inputs (one frame later in a replayed sequence): re-packs the factor
   constants (measurements, pre-integration terms, priors) and initial states of all windows from
   `graphs` and uploads them; fails with SWGN_ERR_INVALID when a graph is NULL or its structure -- block table,
   constness, ordering groups, any factor's block list or kind, prior / chain shapes, program order, is_use masks --
   differs from the one the batch was created with (a fingerprint of all of these is kept per window).
   bytes_h2d (may be NULL) receives the bytes copied. */
swgn_status swgn_batch_update_inputs(swgn_batch* b, const swgn_graph* const* graphs,
                                     int64_t* bytes_h2d);

/* Double-buffered variant for a replayed sequence: swgn_batch_prefetch_inputs packs and uploads the NEXT step's inputs
   into a shadow block on a private copy stream -- it may run on another host thread while swgn_batch_solve /
   swgn_batch_get_states of the current step run (not concurrently with swgn_batch_update_inputs or another prefetch of
   the same batch); swgn_batch_commit_inputs, called between two solves, makes the prefetched inputs the live ones
   (constants by address, states by one scatter launch).  Same checks and errors as swgn_batch_update_inputs. */
swgn_status swgn_batch_prefetch_inputs(swgn_batch* b, const swgn_graph* const* graphs, int64_t* bytes_h2d);
swgn_status swgn_batch_commit_inputs(swgn_batch* b);

/* The whole trust-region solve = ceres::Solve with DENSE_SCHUR + DOGLEG
   (CERES trust_region_minimizer.cc:67-134).  summaries may be NULL, else n_windows entries.
   With is_optimize == 0 this is the export-mode solve. */
swgn_status swgn_batch_solve(swgn_batch* b, swgn_summary* summaries);

/* Timing of the last swgn_batch_solve measured with CUDA events on the library's stream:
   total milliseconds, and the part spent in Schur-elimination launches with their count. */
swgn_status swgn_batch_last_timing(const swgn_batch* b, double* total_ms, double* schur_ms,
                                   int32_t* schur_launches, int32_t* kernel_launches);

/* Algorithmic bytes of one Schur elimination of `window` on the materialised Jacobian (J blocks
   at their stored size, residuals and D in; upper triangle of S, r and the back-substituted y
   out): the numerator of the Schur kernel's HBM roofline. */
int64_t swgn_batch_schur_bytes(const swgn_batch* b, int32_t window);

/* Result read-backs (device -> caller buffer). */
swgn_status swgn_batch_get_state(swgn_batch* b, int32_t window, double* state);
/* Whole-batch variants: the states of all windows back to back (window w starts at the sum of
   graphs[0..w-1]->n_state), moved with one copy. */
int64_t swgn_batch_states_size(const swgn_batch* b);
swgn_status swgn_batch_set_states(swgn_batch* b, const double* states);
swgn_status swgn_batch_get_states(swgn_batch* b, double* states);
/* ceres::internal::{lhs_out, rhs_out, hs_row}: reduced system of the last Eliminate, row-major
   n x n with only the upper triangle meaningful; S and r may be NULL to query n. */
swgn_status swgn_batch_get_reduced(swgn_batch* b, int32_t window, double* S, double* r,
                                   int32_t* n);
/* ceres::internal::lhs_out2: lower Cholesky factor of the last reduced solve, n x n row-major,
   strictly-upper part zero. */
swgn_status swgn_batch_get_cholesky(swgn_batch* b, int32_t window, double* L, int32_t* n);
/* UpdateSchurHessianOnly (RVI/swf/swf_gnss.cpp:65-94): A = L_nn L_nn^T for the trailing
   n_tail rows; A is n_tail x n_tail row-major. */
swgn_status swgn_batch_get_tail_information(swgn_batch* b, int32_t window, int32_t n_tail,
                                            double* A);

/* UpdateSchur (RVI/swf/swf_gnss.cpp:25-61): after an export-mode solve, Schur-reduce the leading
   n_f - n_tail rows of the exported (S, r) onto the trailing n_tail rows with the reference's eigen
   pseudo-inverse (eigenvalues <= 1e-8 dropped): A (n_tail x n_tail row-major, full symmetric) and
   b (n_tail) are what the reference turns into its next marginalisation prior. */
swgn_status swgn_batch_get_head_marginal(swgn_batch* b, int32_t window, int32_t n_tail, double* A,
                                         double* bvec);

/* The next window's marginalisation prior in one call: UpdateSchur as above followed by
   MarginalizationInfo::setmarginalizeinfo(..., Sqrt = true) (RVI/factor/marginalization_factor.cpp:449-475):
   A = V S V', J0 = sqrt(S) V' (n_tail x n_tail row-major), r0 = S^-1/2 V' b, eigenvalues <= 1e-8 dropped --
   exactly the prior_J / prior_r0 arrays of swgn_graph (x0 = the current states of the head blocks).
   A and bvec (may be NULL) return the information form. */
swgn_status swgn_batch_get_marginal_prior(swgn_batch* b, int32_t window, int32_t n_tail, double* J0,
                                          double* r0, double* A, double* bvec);

/* The same for every window of the batch in two launches (one CTA per window) and one read-back: window w reduces onto its
   trailing n_tail[w] rows (0 = skip the window); its J0 (n_tail[w]^2 doubles) is written at J0_all + j_off[w], its r0 at
   r0_all + r_off[w]. */
swgn_status swgn_batch_get_marginal_priors(swgn_batch* b, const int32_t* n_tail, const int64_t* j_off, const int64_t* r_off,
                                           double* J0_all, double* r0_all);

/* Hidden GNSS-frame states of the IMUGNSSFactor chains of `window` after the last Jacobian
   evaluation (the reference updates gnss_poses[i] / gnss_speed_bias[i] in user memory,
   gnss_imu_factor.cpp:601-632): 16 doubles (pose 7, speed-bias 9) per hidden frame in graph
   order; n_frames may be queried with frames NULL. */
swgn_status swgn_batch_get_chain_frames(swgn_batch* b, int32_t window, int32_t* n_frames,
                                        double* frames);

/* ---- staged entry points (used by the parity tests; each is one device pass) ------------- */
/* Evaluate r, cost, gradient g = J^T r at the current state (ProgramEvaluator::Evaluate).
   Output buffers may be NULL.  residuals: n_residuals doubles in the library's row order;
   gradient: n_e + n_f doubles in column order. */
swgn_status swgn_batch_evaluate(swgn_batch* b, int32_t window, double* cost, double* residuals,
                                double* gradient);
/* Column order actually used: for every column block its graph block index, tangent offset and
   tangent size; n_cols may be queried with the arrays NULL. */
swgn_status swgn_batch_get_columns(swgn_batch* b, int32_t window, int32_t* n_cols,
                                   int32_t* block, int32_t* offset, int32_t* size);
/* Row order actually used: for every row block the program-order index of its residual block
   (index into swgn_graph.order semantics) and its first residual row. */
swgn_status swgn_batch_get_rows(swgn_batch* b, int32_t window, int32_t* n_rows, int32_t* factor,
                                int32_t* offset);
/* Dense copy of the block-sparse Jacobian (n_residuals x (n_e+n_f), row-major) for tests. */
swgn_status swgn_batch_get_dense_jacobian(swgn_batch* b, int32_t window, double* J);
/* One linear solve on the current linearisation with LM diagonal D (n_e+n_f entries, may be
   NULL): Eliminate -> Cholesky -> BackSubstitute; x receives the n_e+n_f solution of
   min |J x - r|^2 + |D x|^2 (CERES schur_complement_solver.cc:126-202). */
swgn_status swgn_batch_linear_solve(swgn_batch* b, int32_t window, const double* D, double* x);

/* ---- IMU pre-integration ------------------------------------------------------------------- */
/* IntegrationBase (RVI/factor/integration_base.cpp:5-142), batched on the device: factor f integrates
   the samples [sample_begin[f], sample_begin[f+1]) -- 7 doubles per sample: dt, acc[3], gyr[3]; the first
   sample is (acc_0, gyr_0) of the constructor (its dt is ignored), every further one is one
   push_back(dt, acc, gyr) = one midpoint step with Jacobian and covariance propagation.  bias holds
   linearized_ba[3], linearized_bg[3] per factor; noise = ACC_N, GYR_N, ACC_W, GYR_W.  Output: one
   SWGN_IMU_STRIDE record per factor (exactly what swgn_graph.imu_data / chain_imu_data take), with
   sqrt_info = LLT(covariance^-1).L' (get_sqrtinfo); info[f] = 0 ok, 1 covariance not invertible /
   not positive definite (sqrt_info left zero). */
swgn_status swgn_preintegrate_batch(int32_t device, int32_t n_factors, const int32_t* sample_begin,
                                    const double* samples, const double* bias, const double noise[4],
                                    double* records, int32_t* info);

/* ---- ambiguity resolution (K7/K8) --------------------------------------------------------- */
/* RTKLIB-style lambda()/mlambda as shipped in RVI/gnss/src/lambda.cpp:204-235, batched:
   problem k has n[k] float ambiguities a_k (n[k]) and covariance Q_k (n[k] x n[k], column-major),
   concatenated.  Outputs: F (n[k] x m column-major, concatenated), s (m per problem),
   info (0 ok, -1 failure) per problem.  Runs on the device; bit-identical to the reference's
   double-precision arithmetic (compiled without FMA contraction). */
swgn_status swgn_lambda_batch(int32_t device, int32_t n_problems, const int32_t* n, int32_t m,
                              const double* a, const double* Q, double* F, double* s,
                              int32_t* info);

/* Double-difference construction + LAMBDA + ratio test = the decision part of LambdaSearch
   (RVI/swf/swf_lambda.cpp:101-245) for one window, given the ambiguity information matrix A
   (n x n row-major, from swgn_batch_get_tail_information) and the float values y.
   The GNSS epochs of the window (rovers[0..rover_count-1], oldest first) are given as a CSR
   list: epoch e observes obs_amb[epoch_begin[e] .. epoch_begin[e+1]) where obs_amb is the index
   of the observed ambiguity inside A/y (or -1 when that ambiguity is not part of A) and
   obs_sysfreq = sys*2+f in 0..5.  Epochs are visited newest first; last_fix selects the
   0.2 / 1.4 gate (swf_lambda.cpp:163). */
typedef struct swgn_fix_result {
  int32_t n_dd;          /* rows of D                                                        */
  int32_t status;        /* 0 searched, 1 too few ambiguities (n < 6), 2 too few DD rows,
                            3 lambda() failed                                                */
  int32_t search_ok;     /* ratio test decision                                              */
  int32_t n_different;   /* entries where the two best candidates differ                    */
  double s[2];           /* squared norms of the two best candidates                        */
  double s0_partial, s1_partial;
} swgn_fix_result;
swgn_status swgn_ambiguity_fix(int32_t device, int32_t n, const double* A, const double* y,
                               int32_t n_epochs, const int32_t* epoch_begin,
                               const int32_t* obs_amb, const int32_t* obs_sysfreq,
                               int32_t last_fix, int32_t* dd_pairs /* 2 per row: (a, ref) */,
                               double* F /* n_dd x 2 column-major */, swgn_fix_result* result);

/* MarginalizationInfo::marginalize (RVI/factor/marginalization_factor.cpp:260-377) + getParameterBlocks for ARBITRARY drop
   sets -- what MargFrames runs for MargImagSecondNew / MargRoverOld (RVI/swf/swf.cpp:329-341 -> swf_core.cpp:372-390): the
   factors of graph w are linearised at its state, the blocks with drop[w][b] != 0 are marginalised with the reference's
   eigen pseudo-inverse (eigenvalues <= 1e-8 dropped) and the remaining information becomes the prior r = r0 + J0 (x [-] x0)
   over the keep blocks (block-index order; x0 = the graph's state).  On the device this is one export-mode pass per call for
   all graphs -- drop blocks ordered before the keep blocks in the reduced system, which swgn_batch_get_marginal_priors then
   reduces and square-roots.  block_group is ignored; constant blocks stay constants of their factors and blocks no factor touches take
   no part (neither is a keep block).  Restrictions: order / is_use / chains / host factors must be absent, at least one keep
   block. */
typedef struct swgn_marginalize_output {
  int32_t cap_keep, cap_n;   /* capacities of the caller's buffers                                  */
  int32_t n_keep, n, m;      /* keep blocks, their tangent size, tangent size of the dropped blocks */
  int32_t* keep_block;       /* graph block index of every keep block                               */
  int32_t* keep_idx;         /* its first column in J0                                              */
  double* J0;                /* n x n row-major                                                     */
  double* r0;                /* n                                                                   */
} swgn_marginalize_output;
swgn_status swgn_marginalize(int32_t device, int32_t n_graphs, const swgn_graph* const* graphs, const uint8_t* const* drop,
                             swgn_marginalize_output* outputs);

/* The prior rebuild that follows FIX_CONTINUE_THRESHOLD consecutive accepted fixes (RVI/swf/swf_lambda.cpp:249-355): the
   last marginalisation prior, one FixedIntegerFactor(0, istd) tying a dummy scalar to the reference ambiguity of every
   system / frequency with a fixed double difference, and one FixedIntegerFactor(round(F), istd) per fixed double difference
   tying the same dummy to its other ambiguity, are linearised at x (the reference evaluates with the phase biases at zero:
   PhaseBiasSaveAndReset) and the dummies are marginalised out -- on the device: the dummies are elimination group 0 of an
   export-mode pass and the reduced system goes through the eigen square root, exactly like an epoch's clock terms
   (include/swgn_gnss.h).  All jobs of one call share two batched passes.  The new prior has the keep blocks and columns of
   the old one; its linearisation point is x. */
typedef struct swgn_fixed_integer_job {
  int32_t n_keep, n;          /* keep blocks / tangent size of last_marg_info                                   */
  const int32_t* keep_size;   /* global sizes (7 = pose with the reference's parameterization)                  */
  const int32_t* keep_idx;    /* first tangent column of every keep block                                       */
  const double* x0;           /* linearisation point of the old prior, global sizes, keep order                 */
  const double* J0;           /* n x n row-major                                                                */
  const double* r0;           /* n                                                                              */
  const double* x;            /* where to linearise: current values of the keep blocks, global sizes            */
  int32_t n_dd;
  const int32_t* dd_keep;     /* 2 per double difference: keep-block index of the +1 and of the -1 (reference) ambiguity */
  const double* F;            /* the fixed integers (the caller rounds, :286)                                   */
  const int32_t* dd_sysfreq;  /* sys * 2 + f of every double difference, 0..5                                   */
  double istd;                /* 1 / 0.03 (:312,320)                                                            */
  double* J0_out;             /* n x n row-major                                                                */
  double* r0_out;             /* n                                                                              */
} swgn_fixed_integer_job;
swgn_status swgn_fixed_integer_prior(int32_t device, int32_t n_jobs, const swgn_fixed_integer_job* jobs);

/* The same decision for EVERY window of a batch after its solve, without a host loop (BASELINE configs[3] over a batch):
   launch 1 computes, per window, A = the tail information of the last Cholesky factor (UpdateSchurHessianOnly) and y = the
   current values of the n_tail scalar blocks at the end of the ordering (the float ambiguities, parameter_head); launch 2
   runs the double-difference construction, lambda() and the ratio tests, one thread per window.  Epochs are given as a
   CSR of CSRs: window w owns the epochs [win_epoch[w], win_epoch[w+1]) of epoch_begin (total_epochs + 1 global offsets
   into obs_amb / obs_sysfreq, semantics as above); last_fix (may be NULL = all 0) holds one flag per window.
   results: n_windows entries (status 3 also for windows without a factor); dd_pairs (2 * n_tail ints per window) and F
   (2 * n_tail doubles per window, n_dd x 2 column-major) may be NULL. */
swgn_status swgn_batch_ambiguity_fix(swgn_batch* b, int32_t n_tail, const int32_t* win_epoch, const int32_t* epoch_begin,
                                     const int32_t* obs_amb, const int32_t* obs_sysfreq, const int32_t* last_fix,
                                     swgn_fix_result* results, int32_t* dd_pairs, double* F);

#ifdef __cplusplus
}
#endif
#endif /* SWGN_H_ */
