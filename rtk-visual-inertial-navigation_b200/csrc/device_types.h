// Plain-old-data descriptors shared by the host planner and the CUDA kernels.
//
// HBM layout (one swgn_batch): three pools hold every window back to back (window-major, so one
// CTA streams one contiguous region),
//   ipool : int32  index arrays  (structure; read-only after upload)
//   cpool : double factor constants (read-only after upload)
//   wpool : double work arrays (state, residuals, block-sparse Jacobian values, vectors, the
//           reduced system S | rhs, chunk factors and buffers)
// and WinDesc[w] carries the per-window counts plus 64-bit offsets of each array inside its pool.
// All per-window arrays start on 16-byte boundaries so kernels can use 128-bit / bulk accesses.
#pragma once
#include <stdint.h>

namespace swgn {

enum IArr {
  I_COL_STATE = 0,  // [n_cols] offset of the block inside the window state
  I_COL_SIZE,       // [n_cols] tangent (local) size
  I_COL_GSIZE,      // [n_cols] ambient (global) size
  I_COL_POS,        // [n_cols] tangent offset; e-blocks first
  I_COL_BLOCK,      // [n_cols] graph block index (for read-backs)
  I_TCOL,           // [n_t]    column block of every tangent scalar
  I_ROW_RES,        // [n_rows] first residual row
  I_ROW_NRES,       // [n_rows]
  I_ROW_CELL,       // [n_rows+1] CSR into cells
  I_ROW_FACTOR,     // [n_rows] program-order index of the residual block (read-backs)
  I_CELL_COL,       // [n_cells]
  I_CELL_VAL,       // [n_cells] offset of the cell's values (row-major n_res x col_size) in W_JAC
  I_CELL_SLOT,      // [n_cells] offset of the cell's E'F block inside W_EBUF (absolute within the
                    //           window's W_EBUF), -1 for the e-cell / rows without e-block
  I_CELL_FIRST,     // [n_cells] 1: first row of its chunk that touches this slot (store, don't add)
  I_RS_ROW,         // [n_res] row block of every residual scalar
  I_CHUNK_ROW,      // [n_chunks+1] row range of every chunk
  I_CHUNK_ECOL,     // [n_chunks]
  I_CHUNK_SLOT,     // [n_chunks+1] CSR into slots
  I_CHUNK_FAC,      // [n_chunks] offset of chol(E'E + D^2) (es x es row-major, lower) in W_EFAC
  I_CHUNK_G,        // [n_chunks] offset of L^-1 E'b (es) inside W_EBUF
  I_SLOT_COL,       // [n_slots] f column of the slot, ascending inside a chunk
  I_SLOT_BUF,       // [n_slots] offset of the es x fs block L^-1 E'F inside W_EBUF
  I_TCHUNK,         // [n_tchunks] chunks with e-size <= 3 (one thread each)
  I_WCHUNK,         // [n_wchunks] chunks with e-size 4..16 (one warp each)
  I_CSC_PTR,        // [n_cols+1]
  I_CSC_ROW,        // [nnz] row block
  I_CSC_VAL,        // [nnz] value offset
  I_CSC_NRES,       // [nnz] residual rows of that row block
  I_CSC_RES,        // [nnz] first residual row of that row block
  I_SCELL,          // [n_scells * 8] directory of the touched block cells of S (the kernel reads I_WSTREAM): ps, qs,
                    //                S offset (row*ld+col), 0, term count, diag flag (1: p == q:
                    //                add D^2, carries the rhs as an extra column), first run, run count
  I_TCHUNK_W,       // [n_tchunks_w] small chunks (e-size <= 3) whose slots are fed by several rows (epoch clocks: one row per
                    //               satellite and observable): one WARP each in phase 1a, lanes over the rows
  I_STERM,          // [n_sterms] gather terms of the e-cells (see below): (a, b), or (a, b, b2, 0) on diagonal cells
  I_ROW_CHUNK,      // [n_rows] chunk of the row, -1 without e-block
  I_CHUNK_SIMPLE,   // [n_chunks] 1: small e-block and every slot is fed by exactly one row
  I_SROW,           // [n_srows * 8] rows (with f-cells) of the simple chunks, one thread each in phase 1b; a self-contained
                    //               record per row so that the thread needs ONE index load before its data loads:
                    //               E value offset, nres | es << 8 | n_fcells << 16, chunk factor offset (W_EFAC),
                    //               first f-cell (index into I_CELL_*), then for the first f-cell: value offset,
                    //               W_EBUF slot offset, tangent size; 0
  I_ECELL,          // (unused)
  I_ECELL_G,        // [1] number of e-cells (statistic)
  I_ESTREAM,        // per-warp gather streams of the raw chunk products E'[E | b | F] of the 4..16-dim e-blocks
                    // (phase 1a of k_schur); same stage records as I_WSTREAM with header flag bit 1 set:
                    // w0 = output offset (W_EFAC for the [E'E | E'b] cell, W_EBUF for E'F), w1 = W_EBUF offset of E'b
  I_ESTREAM_PTR,    // [SCHUR_WARPS + 1]
  I_CHOL_MASK,      // [2 * ceil(n_f / 32)] per 32-row panel of the reduced system: 64-bit mask (lo, hi) of the 16-column
                    //               groups that can be non-zero in the panel's rows of U (symbolic block fill-in of the
                    //               right-looking Cholesky); k_chol skips the tiles of all-zero groups
  I_TCHUNK_T,       // [n_tchunks_t] the other small chunks: one THREAD each in phase 1a (I_TCHUNK = both lists, for k_backsub)
  I_WSTREAM,        // [n_wstream * 4] per-warp gather streams of the reduced system (phase 2 of k_schur): stages of
                    //                 1 + SCHUR_STAGE 16-byte records, see "gather stream" below
  I_WSTREAM_PTR,    // [SCHUR_WARPS + 1] first STAGE of every warp's stream
  I_PROJ,           // [n_proj * 8]  state_off[3], jac_off[3], res_off, 0
  I_IMU,            // [n_imu * 12]  state_off[4], jac_off[4], res_off, 0,0,0
  I_GNSS,           // [n_gnss * 8]  kind, state_off[3], jac_off[3], res_off
  I_PRIOR,          // [n_prior * 8] n, n_blk, res_off, blk_begin, J_off, r0_off, 0, 0
  I_PRIOR_BLK,      // [n_prior_blk * 6] state_off, gsize, idx, jac_off, x0_off, local size
  I_UNIT,           // [n_unit * 4]  state_off, jac_off, res_off, 0
  I_CHAIN,          // [n_chain * 8] m (hidden frames), k (phase biases), res_off, first entry in I_CHAIN_BLK,
                    //               C_CHAIN offset, W_CHAIN offset, first hidden frame of the window, 0
  I_CHAIN_BLK,      // [(4 + k) * 2 per chain] state_off, jac_off of pose_i, sb_i, pose_j, sb_j, N_0 ..
  I_HOST,           // [n_host * 4] res_off (-1: not in the reduced program), first entry in I_HOST_BLK, number of blocks,
                    //              offset of the factor's record [r | J_0 | J_1 ...] (row-major nres x GLOBAL size) in W_HOSTBUF
  I_HOST_BLK,       // [4 per block] state_off, jac_off (-1 constant), global size, tangent size
  I_SB_HDR,         // [sb_nbatch * SB_HDR_INTS] batch headers of the streamed Schur elimination (see "streamed Schur" below)
  I_SB_REC,         // record packages of the batches, back to back (every package a multiple of 4 ints)
  I_ACC_MAP,        // [n_fb * n_fb] block cell (p, q), p <= q, of the reduced system -> offset of its compact accumulator
                    //               (ps x (qs + [p == q]) doubles, row-major, the extra column of diagonal cells = rhs), -1: untouched
  NUM_IARR
};

// Gather term of the reduced system: S_pq (+/-)= A^T B with A (m x ps) at JW[a], B (m x qs) at
// JW[b], where JW is the concatenation W_JAC | W_EBUF | W_RES of the window.
// Blocks with more than 4 rows are split into slabs of <= 4 rows by the planner, so every entry is one MMA:
//   word0 = a (28 bits) | rows-1 << 28 | subtract << 30 ;  word1 = b ;  diagonal cells: word2 = b2, word3 = 0
// b2 is the offset of the m-vector paired with A for the rhs column: b or w_g.

// Gather stream of the reduced system.  The 8x8 output tiles of all touched block cells of S are dealt to the
// SCHUR_WARPS warps of the window's CTA (longest first onto the least loaded warp); every warp gets ONE linear
// stream of stages, each stage = one 16-byte header + SCHUR_STAGE 16-byte term entries of the same tile:
//   header: w0 = S offset of the block cell (row * ld + col), w1 = first S row of block p,
//           w2 = bit 0: the tile ends with this stage (store it), bit 1: e-cell target, w3 = meta
//   term:   w0 = a (28 bits) | rows-1 << 28 | subtract << 30 | padding << 31,  w1 = b,
//           w2 = b2 (rhs operand of diagonal cells), w3 = 0
// meta = ps | qs << 6 | (ti / 8) << 12 | (tj / 8) << 15 | diag << 18.  A tile's term list is padded to a multiple
// of SCHUR_STAGE with entries whose loads are switched off, so a stage is straight-line code.
enum { SCHUR_WARPS = 8, SCHUR_STAGE = 8 };

// Streamed Schur elimination (k_schur_stream.cu): the window's Jacobian is cut, in row order, into BATCHES of whole
// chunks (all rows of one e-block) and single rows without e-block.  A batch is staged into shared memory by TMA bulk
// copies (its contiguous J segment, its residual segment, its record package), its chunk products / factors / W blocks
// are computed on chip, and its terms F'F / -W'W are accumulated into COMPACT accumulators of the touched block cells
// of S that stay in shared memory for the whole window; S is written to HBM once at the end.
// Operand area OA (doubles): [stage 0: J | res][stage 1: J | res][wbuf: W (E-buffer segment) | chunk factors]; batch k
// uses stage k & 1; every offset inside a record is a 16-bit absolute OA offset.
//   header (SB_HDR_INTS ints): rec_off, rec_len (ints), j_src (doubles, relative to W_JAC, even), j_len (even),
//           r_src (relative to W_RES, even), r_len, eb_src (W_EBUF), eb_len, ef_src (W_EFAC), ef_len,
//           n_tchunk, off_textra, n_mchunk, off_tchunk, off_mchunk (int offsets inside the package), sec_len
//   package: the first sec_len ints travel to shared memory with the batch: [SB_WARPS + 1] pointers of the phase-A run
//           streams, [SB_WARPS + 1] of the phase-C run streams (int offsets inside the package), then the sections
//           below; the run streams follow and are read from L2 through per-warp cp.async rings:
//     tchunk (int4): offset of its trow list, n_rows | es << 16, factor OA | g OA << 16, tangent position of the e-block
//       trow (int4, one per row of the chunk): E OA | nres << 16, residual OA, F OA | W OA << 16 of the first f-cell,
//                      fs | n_fcells << 8 | index of its further f-cells in the textra list << 16
//       textra (int4): F OA, W OA, fs, 0
//     mchunk (2 x int4): es, tangent position, factor OA, n_slots + 1; slot table offset, 0, 0, 0
//       slot (int2): OA offset of the es x fs block (last entry: g, fs = 1), fs
//     run (int4 + n x int2): dst, rhs dst, n_plus | n_minus << 12 | first << 24 | ecell << 25, meta (as in the gather
//       streams); the n_plus terms that add come first, then the n_minus terms that subtract, both counts multiples of 4
//       term: a OA | b OA << 16,  b2 OA | row mask << 16 (bit k: row k of the <= 4-row slab exists); all-zero = padding
enum { SB_WARPS = 16, SB_HDR_INTS = 16, SB_JCAP = 2560, SB_TERMCAP = 4096, SB_RING_BYTES = 512, SB_SMEM_BUDGET = 227 * 1024 - 1024 };

enum CArr {
  C_GLOBALS = 0,  // Pbg[3], gravity[3], proj_sqrt_info[4], cauchy_a, pad -> 12
  C_PROJ_UV,      // [n_proj * 2]
  C_IMU,          // [n_imu * IMU_DEV_STRIDE]
  C_GNSS,         // [n_gnss * GNSS_DEV_STRIDE]
  C_PRIOR_J,      // concatenated n x n row-major
  C_PRIOR_R0,
  C_PRIOR_X0,
  C_UNIT,         // [n_unit]
  C_CHAIN,        // IMUGNSSFactor constants, per chain (ChainLayout): m+1 IMU device records, m frame records
                  // (the ABI's SWGN_CHAIN_FRAME_STRIDE record: initial hidden state, linearisation point, rhs,
                  // Hessian), m pose-N Hessians (15 x k), the N-N Hessian (k x k) and the N rhs (k)
  NUM_CARR
};

enum WArr {
  W_X = 0,   // [n_state] current point
  W_XCAND,   // [n_state] candidate point
  W_XBEST,   // [n_state] lowest-cost point seen (what goes back to the user)
  W_X0,      // [n_state] state at upload (restored when a solve fails)
  W_JAC,     // [n_jac] block-sparse Jacobian values, cells in row order      \  these three are
  W_EBUF,    // [n_ebuf] per chunk: L^-1 E'F slot blocks, then L^-1 E'b        | adjacent: "JW"
  W_RES,     // [n_res]                                                       /
  W_MRES,    // [n_res] J * v scratch (Cauchy point, model residuals)
  W_DIAG,    // [n_t] sqrt(clamp(colnorm^2))
  W_G,       // [n_t] J^T r
  W_GHAT,    // [n_t] J^T r / diag
  W_GN,      // [n_t] scaled Gauss-Newton step
  W_STEP,    // [n_t] trust-region step (unscaled)
  W_Y,       // [n_t] linear solve output [y_e ; z]
  W_LMD,     // [n_t] LM diagonal D = diag * sqrt(mu)
  W_EFAC,    // chunk Cholesky factors
  W_S,       // [n_f * ld] reduced system, row-major upper triangle; column n_f holds the rhs;
             //            overwritten by its Cholesky factor U (S = U^T U) and U^-T rhs
  W_SCOPY,   // [n_f * ld] copy of S|rhs before factorisation (staged test entry point only)
  W_SCALE,   // [n_t] Jacobi scaling 1 / (1 + |column of the initial Jacobian|) (jacobi_scaling only)
  W_HOSTBUF, // residuals and global-size Jacobians of the host-evaluated factors, uploaded before every evaluation
  W_CHAIN,   // mutable state of the IMUGNSSFactor chains, per chain (ChainLayout): hidden frame states, history
             // flag, states of the last Jacobian evaluation, INC, saved elimination blocks (hmn_save,
             // rhsmn_save), schur_jacobian, schur_residual, cost of the current evaluation
  NUM_WARR
};

enum { IMU_DEV_STRIDE = 296, GNSS_DEV_STRIDE = 12 };
// IMU device record: [0..23] as the ABI record's first 24 doubles, [24..68] the five 3x3 blocks
// dp_dba, dp_dbg, dq_dbg, dv_dba, dv_dbg (row-major), [70..294] sqrt_info 15x15 row-major
enum { IMU_DEV_BLOCKS = 24, IMU_DEV_SQRT = 70 };

enum FactorKind { K_PROJ = 0, K_IMU = 1, K_GNSS = 2, K_PRIOR = 3, K_UNIT = 4, K_CHAIN = 5, K_HOST = 6, NUM_KINDS = 7 };

// Offsets (doubles) inside one chain's C_CHAIN constants and W_CHAIN work area; m hidden frames,
// k phase biases, n = 30 + k residuals.
enum { CHAIN_FRAME_STRIDE = 274, MAX_CHAIN_K = 48 };
struct ChainLayout {
  int m, k, n;
  // constants
  int c_imu, c_frame, c_frameN, pn_stride, c_NN, c_Nrhs, c_size;
  // work
  int w_frames, w_flags, w_old, w_inc, w_save, save_stride, w_J, w_r, w_size;
#if defined(__CUDACC__)
  __host__ __device__
#endif
  ChainLayout(int m_, int k_) : m(m_), k(k_), n(30 + k_) {
    auto al = [](int x) { return (x + 1) & ~1; };
    c_imu = 0;
    c_frame = (m + 1) * 296;
    c_frameN = c_frame + m * CHAIN_FRAME_STRIDE;
    pn_stride = al(15 * k);
    c_NN = c_frameN + m * pn_stride;
    c_Nrhs = c_NN + al(k * k);
    c_size = c_Nrhs + al(k);
    w_frames = 0;                       // m x 16: pose 7, speed-bias 9
    w_flags = m * 16;                   // [0] history_flag, [1] input epoch, [2] cost, [3] evaluation failed
    w_old = w_flags + 4;                // Pi 7, Bi 9, Pj 7, Bj 9, N k
    w_inc = w_old + 32 + al(k);         // n
    w_save = w_inc + al(n);             // per frame: Amm_inv 225 | H12 225 | H1N 15k | H10 225 | rhs 15
    save_stride = al(3 * 225 + 15 * k + 15);
    w_J = w_save + m * save_stride;     // n x n schur_jacobian
    w_r = w_J + al(n * n);              // n schur_residual
    w_size = w_r + al(n);
  }
};

enum { MAX_WARP_E = 16, MAX_COL_SIZE = 63 };

struct WinDesc {
  int32_t n_state, n_cols, n_ecols, n_e, n_f, n_t, n_res, n_rows, n_cells, n_chunks, n_slots;
  int32_t n_jac, n_ebuf, ld, n_proj, n_imu, n_gnss, n_prior, n_prior_blk, n_unit, n_efac, n_head;
  int32_t n_tchunks, n_wchunks, n_scells, n_sterms, n_srows, max_prior_n, max_wbuf, n_ecells;
  int32_t n_chain, n_chain_frames, max_chain_k, n_wstream;
  int32_t n_tchunks_t, n_tchunks_w;
  // streamed Schur elimination: sb_ok = the window fits the on-chip budget (else the gather kernel k_schur runs it)
  int32_t sb_ok, sb_nbatch, sb_acc, sb_jcap, sb_rcap, sb_ecap, sb_fcap, sb_reccap, n_fb, n_host;
  int32_t n_hostbuf, pad1_;
  int64_t ioff[NUM_IARR];
  int64_t coff[NUM_CARR];
  int64_t woff[NUM_WARR];
};

// per-window trust-region state (TrustRegionMinimizer + DoglegStrategy + StepEvaluator members)
struct TRState {
  double x_cost, candidate_cost, minimum_cost, model_cost_change, fixed_cost, initial_cost;
  double final_cost;           // running minimum of the recorded iteration costs
  double radius, mu, alpha, dogleg_step_norm, x_norm;
  double gradient_max_norm;
  double ghat_norm;            // |J^T r / d|
  double relative_decrease, step_norm, iter_cost;
  double se_min, se_cur, se_ref, se_cand, se_acc_ref, se_acc_cand;
  int32_t se_nonmono;
  int32_t iteration, num_successful, num_unsuccessful, num_consecutive_invalid;
  int32_t num_linear_solves;
  int32_t reuse;               // DoglegStrategy::reuse_
  int32_t active;              // minimiser still running
  int32_t termination;         // SWGN_CONVERGENCE / NO_CONVERGENCE / FAILURE
  int32_t last_successful;     // iteration_summary_.step_is_successful of the finished iteration
  int32_t need_solve;          // this iteration still needs a (re)try of the linear solve
  int32_t solve_ok;            // the current Gauss-Newton step W_GN is valid
  int32_t chol_ok;             // the reduced factorisation of this tick succeeded
  int32_t step_valid;          // model_cost_change > 0
  int32_t accepted;            // step accepted at the end of this iteration
  int32_t have_factor;         // W_S currently holds a Cholesky factor (lhs_out2 available)
  int32_t have_reduced;        // W_S currently holds S | rhs of an export-mode eliminate
  int32_t alpha_valid;         // the Cauchy-point alpha of the current linearisation has been computed (lazily, k_step)
  double decrease_factor;      // LevenbergMarquardtStrategy::decrease_factor_
  double ghat_sq;              // |J^T r / d|^2, the numerator of alpha
};

struct SolverParams {
  int32_t max_num_iterations, max_num_consecutive_invalid_steps;
  double initial_radius, max_radius, min_radius, min_relative_decrease;
  double min_lm_diagonal, max_lm_diagonal;
  double function_tolerance, gradient_tolerance, parameter_tolerance, min_mu;
  int32_t is_optimize, n_parameter_head, export_mode;
  int32_t strategy;            // SWGN_DOGLEG / SWGN_LEVENBERG_MARQUARDT
  int32_t jacobi_scaling, pad;
  double max_radius_lm;        // (= max_radius; LevenbergMarquardtStrategy clamps on acceptance)
};

}  // namespace swgn
