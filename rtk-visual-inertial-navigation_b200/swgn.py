"""ctypes binding of the C ABI in include/swgn.h (libswgn.so) and of the synthetic-window
generator (libswgn_synth.so).  Harness-side only: tests/, bench.py and __graft_entry__ use it to
reach the product exactly the way a C/C++ host would.  There is no Python compute path and no CPU
fallback: loading fails loudly when the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
P = C.POINTER


class Graph(C.Structure):
    _fields_ = [
        ("n_blocks", i32), ("block_size", P(i32)), ("block_manifold", P(i32)),
        ("block_const", P(i32)), ("block_group", P(i32)), ("block_offset", P(i32)),
        ("n_state", i32), ("state", P(f64)),
        ("Pbg", f64 * 3), ("gravity", f64 * 3), ("proj_sqrt_info", f64 * 4), ("proj_cauchy_a", f64),
        ("n_proj", i32), ("proj_blocks", P(i32)), ("proj_uv", P(f64)),
        ("n_imu", i32), ("imu_blocks", P(i32)), ("imu_data", P(f64)),
        ("n_gnss", i32), ("gnss_kind", P(i32)), ("gnss_blocks", P(i32)), ("gnss_data", P(f64)),
        ("n_prior", i32), ("prior_n", P(i32)), ("prior_blk_begin", P(i32)),
        ("prior_blocks", P(i32)), ("prior_blk_idx", P(i32)),
        ("prior_x0_begin", P(i64)), ("prior_x0", P(f64)),
        ("prior_J_begin", P(i64)), ("prior_J", P(f64)),
        ("prior_r_begin", P(i64)), ("prior_r0", P(f64)),
        ("n_unit", i32), ("unit_block", P(i32)), ("unit_istd", P(f64)),
        ("n_order", i32), ("order", P(C.c_uint32)),
        ("is_use", P(C.c_uint8)),
        ("n_chain", i32), ("chain_blk_begin", P(i32)), ("chain_blocks", P(i32)),
        ("chain_frame_begin", P(i32)), ("chain_frame_data", P(f64)), ("chain_frame_N", P(f64)),
        ("chain_N", P(f64)), ("chain_imu_data", P(f64)),
        ("n_host", i32), ("host_nres", P(i32)), ("host_blk_begin", P(i32)), ("host_blocks", P(i32)),
        ("host_eval", C.c_void_p), ("host_user", C.c_void_p),
    ]


HOST_EVAL_FN = C.CFUNCTYPE(i32, C.c_void_p, i32, P(P(f64)), P(f64), P(P(f64)))


class Options(C.Structure):
    _fields_ = [
        ("max_num_iterations", i32), ("max_num_consecutive_invalid_steps", i32),
        ("initial_trust_region_radius", f64), ("max_trust_region_radius", f64),
        ("min_trust_region_radius", f64), ("min_relative_decrease", f64),
        ("min_lm_diagonal", f64), ("max_lm_diagonal", f64),
        ("function_tolerance", f64), ("gradient_tolerance", f64), ("parameter_tolerance", f64),
        ("dogleg_min_mu", f64),
        ("is_optimize", i32), ("n_parameter_head", i32), ("device", i32), ("trust_region_strategy", i32),
        ("jacobi_scaling", i32), ("reserved_", i32),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("initial_cost", f64), ("final_cost", f64), ("fixed_cost", f64),
        ("num_successful_steps", i32), ("num_unsuccessful_steps", i32),
        ("num_iterations", i32), ("num_linear_solves", i32), ("termination_type", i32),
        ("n_e", i32), ("n_f", i32), ("n_residuals", i32),
    ]


class FixResult(C.Structure):
    _fields_ = [("n_dd", i32), ("status", i32), ("search_ok", i32), ("n_different", i32),
                ("s", f64 * 2), ("s0_partial", f64), ("s1_partial", f64)]


class SynthConfig(C.Structure):
    _fields_ = [("n_keyframes", i32), ("n_landmarks", i32), ("n_gnss_epochs", i32),
                ("n_sats", i32), ("seed0", C.c_uint64), ("state_noise", f64),
                ("composition", i32), ("hidden_per_gap", i32), ("bias_walk_scale", f64),
                ("hidden_bias_istd", f64), ("variant", i32), ("pad_", i32)]


SYNTH_SPP, SYNTH_FIXED_INTEGER, SYNTH_FREE_EXTRINSIC = 1, 2, 4


def _dp(a):
    return a.ctypes.data_as(P(f64))


def _ip(a):
    return a.ctypes.data_as(P(i32))


_synth = None
_lib = None


def synth_lib():
    global _synth
    if _synth is None:
        L = C.CDLL(os.path.join(HERE, "libswgn_synth.so"))
        L.swgn_synth_create.restype = C.c_void_p
        L.swgn_synth_create.argtypes = [P(SynthConfig), C.c_uint64]
        L.swgn_synth_destroy.argtypes = [C.c_void_p]
        L.swgn_synth_graph.restype = P(Graph)
        L.swgn_synth_graph.argtypes = [C.c_void_p]
        L.swgn_synth_truth.restype = P(f64)
        L.swgn_synth_truth.argtypes = [C.c_void_p]
        L.swgn_synth_options.argtypes = [C.c_void_p, P(Options)]
        L.swgn_synth_info.argtypes = [C.c_void_p, P(i32)]
        L.swgn_synth_default_config.argtypes = [i32, P(SynthConfig)]
        L.swgn_synth_ambiguity_epochs.restype = i32
        L.swgn_synth_ambiguity_epochs.argtypes = [C.c_void_p, P(i32), P(i32), P(i32), P(i32)]
        L.swgn_synth_true_ambiguities.restype = P(f64)
        L.swgn_synth_true_ambiguities.argtypes = [C.c_void_p]
        L.swgn_synth_chain_truth.restype = i32
        L.swgn_synth_chain_truth.argtypes = [C.c_void_p, P(f64)]
        _synth = L
    return _synth


class SynthWindow:
    """One synthetic window.  which=1: 5 KF x 50 LM VI-only; which=2: the
    20 KF x 300 LM x 10 GNSS-epoch BASELINE window; which=3: the same window in composition A
    (GNSS frames hidden inside IMUGNSSFactor chains); which=4: a small composition-A window."""

    def __init__(self, which=2, window_id=0, **overrides):
        L = synth_lib()
        cfg = SynthConfig()
        L.swgn_synth_default_config(which, C.byref(cfg))
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.h = L.swgn_synth_create(C.byref(cfg), window_id)
        if not self.h:
            raise RuntimeError("swgn_synth_create failed")
        self.graph_p = L.swgn_synth_graph(self.h)
        self.graph = self.graph_p.contents
        info = (i32 * 8)()
        L.swgn_synth_info(self.h, info)
        (self.n_frames, self.n_obs, self.n_imu, self.n_gnss, self.n_amb, self.first_amb_block,
         self.n_prior_rows, self.n_blocks) = list(info)
        self.n_state = self.graph.n_state

    def options(self):
        o = Options()
        synth_lib().swgn_synth_options(self.h, C.byref(o))
        return o

    def state0(self):
        return np.ctypeslib.as_array(self.graph.state, shape=(self.n_state,)).copy()

    def truth(self):
        return np.ctypeslib.as_array(synth_lib().swgn_synth_truth(self.h), shape=(self.n_state,)).copy()

    def block_offsets(self):
        return np.ctypeslib.as_array(self.graph.block_offset, shape=(self.n_blocks,)).copy()

    def ambiguity_epochs(self):
        L = synth_lib()
        n_obs = i32()
        ne = L.swgn_synth_ambiguity_epochs(self.h, None, None, None, C.byref(n_obs))
        eb = np.zeros(ne + 1, np.int32)
        oa = np.zeros(max(n_obs.value, 1), np.int32)
        sf = np.zeros(max(n_obs.value, 1), np.int32)
        L.swgn_synth_ambiguity_epochs(self.h, _ip(eb), _ip(oa), _ip(sf), None)
        return eb, oa[:n_obs.value], sf[:n_obs.value]

    def chain_truth(self):
        """Ground truth of the hidden GNSS frames of the IMUGNSSFactor chains, (n_frames, 16)."""
        n = synth_lib().swgn_synth_chain_truth(self.h, None)
        out = np.zeros((max(n, 1), 16))
        synth_lib().swgn_synth_chain_truth(self.h, _dp(out))
        return out[:n]

    def chain_frames0(self):
        """Initial hidden-frame states as given to the solver, (n_frames, 16)."""
        g = self.graph
        if g.n_chain == 0:
            return np.zeros((0, 16))
        n = g.chain_frame_begin[g.n_chain]
        a = np.ctypeslib.as_array(g.chain_frame_data, shape=(n, 274))
        return a[:, :16].copy()

    def true_ambiguities(self):
        return np.ctypeslib.as_array(synth_lib().swgn_synth_true_ambiguities(self.h),
                                     shape=(self.n_amb,)).copy()

    def __del__(self):
        try:
            if self.h:
                synth_lib().swgn_synth_destroy(self.h)
                self.h = None
        except Exception:
            pass


def state_error_by_kind(w, x, xo):
    """max |dx| / max(1, |x|) of two state vectors of SynthWindow w per kind of parameter block: pose translation,
    pose quaternion, speed, accelerometer bias, gyro bias, landmark, ambiguity, other scalars (clocks, drifts)."""
    g = w.graph
    out = {}

    def put(name, a, b):
        if len(a):
            e = float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))
            out[name] = max(out.get(name, 0.0), e)
    for bi in range(g.n_blocks):
        o, s = g.block_offset[bi], g.block_size[bi]
        a, b = x[o:o + s], xo[o:o + s]
        if s == 7:
            put("position", a[:3], b[:3])
            put("quaternion", a[3:], b[3:])
        elif s == 9:
            put("speed", a[:3], b[:3])
            put("acc bias", a[3:6], b[3:6])
            put("gyro bias", a[6:], b[6:])
        elif s == 3:
            put("landmark", a, b)
        elif w.n_amb and w.first_amb_block <= bi < w.first_amb_block + w.n_amb:
            put("ambiguity", a, b)
        else:
            put("scalar", a, b)
    return out


def lib():
    """libswgn.so: the CUDA product.  Raises if it is not built -- there is no fallback."""
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libswgn.so")
        if not os.path.exists(path):
            raise RuntimeError("libswgn.so is not built (run __graft_entry__.build()); "
                               "there is no CPU fallback for the solver")
        L = C.CDLL(path)
        L.swgn_last_error.restype = C.c_char_p
        L.swgn_version.restype = C.c_char_p
        L.swgn_default_options.argtypes = [P(Options)]
        L.swgn_batch_create.argtypes = [P(Options), i32, P(P(Graph)), P(C.c_void_p)]
        L.swgn_batch_destroy.argtypes = [C.c_void_p]
        L.swgn_plan_probe.argtypes = [P(Graph), i32, P(i32)]
        L.swgn_plan_order.argtypes = [P(Graph), i32, P(i32), P(i32), P(i32), P(i32)]
        L.swgn_batch_size.argtypes = [C.c_void_p]
        L.swgn_release_cached_memory.restype = C.c_int64
        L.swgn_release_cached_memory.argtypes = []
        L.swgn_batch_set_state.argtypes = [C.c_void_p, i32, P(f64)]
        L.swgn_batch_solve.argtypes = [C.c_void_p, P(Summary)]
        L.swgn_batch_update_inputs.argtypes = [C.c_void_p, P(P(Graph)), P(i64)]
        L.swgn_batch_prefetch_inputs.argtypes = [C.c_void_p, P(P(Graph)), P(i64)]
        L.swgn_batch_commit_inputs.argtypes = [C.c_void_p]
        L.swgn_batch_last_timing.argtypes = [C.c_void_p, P(f64), P(f64), P(i32), P(i32)]
        L.swgn_batch_get_state.argtypes = [C.c_void_p, i32, P(f64)]
        L.swgn_batch_schur_bytes.restype = i64
        L.swgn_batch_schur_bytes.argtypes = [C.c_void_p, i32]
        L.swgn_batch_states_size.restype = i64
        L.swgn_batch_states_size.argtypes = [C.c_void_p]
        L.swgn_batch_set_states.argtypes = [C.c_void_p, P(f64)]
        L.swgn_batch_get_states.argtypes = [C.c_void_p, P(f64)]
        L.swgn_batch_get_reduced.argtypes = [C.c_void_p, i32, P(f64), P(f64), P(i32)]
        L.swgn_batch_get_cholesky.argtypes = [C.c_void_p, i32, P(f64), P(i32)]
        L.swgn_batch_get_tail_information.argtypes = [C.c_void_p, i32, i32, P(f64)]
        L.swgn_batch_evaluate.argtypes = [C.c_void_p, i32, P(f64), P(f64), P(f64)]
        L.swgn_batch_get_columns.argtypes = [C.c_void_p, i32, P(i32), P(i32), P(i32), P(i32)]
        L.swgn_batch_get_rows.argtypes = [C.c_void_p, i32, P(i32), P(i32), P(i32)]
        L.swgn_batch_get_dense_jacobian.argtypes = [C.c_void_p, i32, P(f64)]
        L.swgn_batch_linear_solve.argtypes = [C.c_void_p, i32, P(f64), P(f64)]
        L.swgn_preintegrate_batch.argtypes = [i32, i32, P(i32), P(f64), P(f64), P(f64), P(f64), P(i32)]
        L.swgn_batch_get_head_marginal.argtypes = [C.c_void_p, i32, i32, P(f64), P(f64)]
        L.swgn_batch_get_marginal_prior.argtypes = [C.c_void_p, i32, i32, P(f64), P(f64), P(f64), P(f64)]
        L.swgn_batch_get_marginal_priors.argtypes = [C.c_void_p, P(i32), P(i64), P(i64), P(f64), P(f64)]
        L.swgn_batch_get_chain_frames.argtypes = [C.c_void_p, i32, P(i32), P(f64)]
        L.swgn_lambda_batch.argtypes = [i32, i32, P(i32), i32, P(f64), P(f64), P(f64), P(f64), P(i32)]
        L.swgn_ambiguity_fix.argtypes = [i32, i32, P(f64), P(f64), i32, P(i32), P(i32), P(i32),
                                         i32, P(i32), P(f64), P(FixResult)]
        L.swgn_batch_ambiguity_fix.argtypes = [C.c_void_p, i32, P(i32), P(i32), P(i32), P(i32), P(i32), P(FixResult), P(i32), P(f64)]
        _lib = L
    return _lib


def plan_probe(graph_p, n_parameter_head=0):
    """Host-only preprocessing of one window; returns (status, dict)."""
    info = (i32 * 16)()
    st = lib().swgn_plan_probe(graph_p, n_parameter_head, info)
    keys = ["n_cols", "n_ecols", "n_e", "n_f", "n_t", "n_res", "n_rows", "n_chunks", "n_jac",
            "n_scells", "n_sterms", "n_srows"]
    d = {k: info[i] for i, k in enumerate(keys)}
    d["schur_bytes"] = (info[12] & 0xffffffff) | (info[13] << 32)
    d["n_mma"] = info[14]
    return st, d


def plan_chol_masks(graph_p, n_parameter_head=0):
    """Host-only: symbolic fill-in masks of the reduced system (one uint64 per 32-row panel)."""
    L = lib()
    L.swgn_plan_chol_masks.argtypes = [P(Graph), i32, P(i32), P(C.c_uint64)]
    n = i32()
    st = L.swgn_plan_chol_masks(graph_p, n_parameter_head, C.byref(n), None)
    if st != 0:
        raise RuntimeError("swgn_plan_chol_masks failed: %s" % L.swgn_last_error().decode())
    m = np.zeros(max(n.value, 1), np.uint64)
    L.swgn_plan_chol_masks(graph_p, n_parameter_head, C.byref(n), m.ctypes.data_as(P(C.c_uint64)))
    return m[:n.value]


def plan_order(graph_p, n_parameter_head=0):
    """(status, column block ids in elimination order, residual-block index per row) from the host planner."""
    L = lib()
    nc, nr = i32(), i32()
    st = L.swgn_plan_order(graph_p, n_parameter_head, C.byref(nc), None, C.byref(nr), None)
    if st != 0:
        return st, None, None
    cols, rows = np.zeros(nc.value, np.int32), np.zeros(nr.value, np.int32)
    st = L.swgn_plan_order(graph_p, n_parameter_head, C.byref(nc), _ip(cols), C.byref(nr), _ip(rows))
    return st, cols, rows


def release_cached_memory():
    """Bytes of cached device / pinned slabs handed back to the driver (swgn_release_cached_memory)."""
    return int(lib().swgn_release_cached_memory())


def default_options():
    o = Options()
    lib().swgn_default_options(C.byref(o))
    return o


def _check(st, what):
    if st != 0:
        raise RuntimeError("%s failed: status %d: %s" % (what, st, lib().swgn_last_error().decode()))


class Batch:
    """swgn_batch: n independent windows resident on one device."""

    def __init__(self, graph_ptrs, options):
        L = lib()
        n = len(graph_ptrs)
        arr = (P(Graph) * n)(*graph_ptrs)
        self.h = C.c_void_p()
        self.n = n
        self.options = options
        self._graphs = arr
        _check(L.swgn_batch_create(C.byref(options), n, arr, C.byref(self.h)), "swgn_batch_create")

    def update_inputs(self, graph_ptrs=None):
        """Re-upload factor constants and initial states (same structure); returns H2D bytes."""
        arr = self._graphs if graph_ptrs is None else (P(Graph) * self.n)(*graph_ptrs)
        nb = i64()
        _check(lib().swgn_batch_update_inputs(self.h, arr, C.byref(nb)), "swgn_batch_update_inputs")
        return nb.value

    def prefetch_inputs(self, graph_ptrs=None):
        """Pack and upload the next step's inputs into the shadow block (may run on another thread during solve())."""
        arr = self._graphs if graph_ptrs is None else (P(Graph) * self.n)(*graph_ptrs)
        nb = i64()
        _check(lib().swgn_batch_prefetch_inputs(self.h, arr, C.byref(nb)), "swgn_batch_prefetch_inputs")
        return nb.value

    def commit_inputs(self):
        _check(lib().swgn_batch_commit_inputs(self.h), "swgn_batch_commit_inputs")

    def solve(self, summaries=None):
        sm = summaries if summaries is not None else (Summary * self.n)()
        _check(lib().swgn_batch_solve(self.h, sm), "swgn_batch_solve")
        return sm

    def timing(self):
        t, s = f64(), f64()
        nl, kl = i32(), i32()
        _check(lib().swgn_batch_last_timing(self.h, C.byref(t), C.byref(s), C.byref(nl), C.byref(kl)),
               "swgn_batch_last_timing")
        return t.value, s.value, nl.value, kl.value

    def set_state(self, w, x):
        x = np.ascontiguousarray(x, np.float64)
        _check(lib().swgn_batch_set_state(self.h, w, _dp(x)), "swgn_batch_set_state")

    def get_state(self, w, n_state):
        x = np.zeros(n_state)
        _check(lib().swgn_batch_get_state(self.h, w, _dp(x)), "swgn_batch_get_state")
        return x

    def schur_bytes(self, w):
        return lib().swgn_batch_schur_bytes(self.h, w)

    def states_size(self):
        return lib().swgn_batch_states_size(self.h)

    def get_states(self, out=None):
        if out is None:
            out = np.zeros(self.states_size())
        _check(lib().swgn_batch_get_states(self.h, _dp(out)), "swgn_batch_get_states")
        return out

    def set_states(self, x):
        x = np.ascontiguousarray(x, np.float64)
        assert x.size == self.states_size()
        _check(lib().swgn_batch_set_states(self.h, _dp(x)), "swgn_batch_set_states")

    def get_reduced(self, w):
        n = i32()
        _check(lib().swgn_batch_get_reduced(self.h, w, None, None, C.byref(n)), "get_reduced")
        S = np.zeros((n.value, n.value))
        r = np.zeros(n.value)
        _check(lib().swgn_batch_get_reduced(self.h, w, _dp(S), _dp(r), C.byref(n)), "get_reduced")
        return S, r

    def get_cholesky(self, w):
        n = i32()
        _check(lib().swgn_batch_get_cholesky(self.h, w, None, C.byref(n)), "get_cholesky")
        Lm = np.zeros((n.value, n.value))
        _check(lib().swgn_batch_get_cholesky(self.h, w, _dp(Lm), C.byref(n)), "get_cholesky")
        return Lm

    def tail_information(self, w, n_tail):
        A = np.zeros((n_tail, n_tail))
        _check(lib().swgn_batch_get_tail_information(self.h, w, n_tail, _dp(A)), "tail_information")
        return A

    def head_marginal(self, w, n_tail):
        """UpdateSchur: (A, b) of the trailing n_tail rows after an export-mode solve."""
        A = np.zeros((n_tail, n_tail))
        bv = np.zeros(n_tail)
        _check(lib().swgn_batch_get_head_marginal(self.h, w, n_tail, _dp(A), _dp(bv)), "head_marginal")
        return A, bv

    def marginal_prior(self, w, n_tail):
        """UpdateSchur + setmarginalizeinfo: (J0, r0, A, b) of the trailing n_tail rows after an export-mode solve."""
        J0, A = np.zeros((n_tail, n_tail)), np.zeros((n_tail, n_tail))
        r0, bv = np.zeros(n_tail), np.zeros(n_tail)
        _check(lib().swgn_batch_get_marginal_prior(self.h, w, n_tail, _dp(J0), _dp(r0), _dp(A), _dp(bv)), "marginal_prior")
        return J0, r0, A, bv

    def marginal_priors(self, n_tail):
        """The prior of every window in two launches: list of (J0, r0); n_tail[w] = 0 skips window w (None, None)."""
        nt = np.ascontiguousarray(n_tail, np.int32)
        j_off = np.concatenate([[0], np.cumsum(nt.astype(np.int64) ** 2)]).astype(np.int64)
        r_off = np.concatenate([[0], np.cumsum(nt.astype(np.int64))]).astype(np.int64)
        J, r = np.zeros(max(int(j_off[-1]), 1)), np.zeros(max(int(r_off[-1]), 1))
        _check(lib().swgn_batch_get_marginal_priors(self.h, _ip(nt), j_off.ctypes.data_as(P(i64)), r_off.ctypes.data_as(P(i64)), _dp(J), _dp(r)),
               "marginal_priors")
        return [(J[j_off[w]:j_off[w + 1]].reshape(nt[w], nt[w]).copy(), r[r_off[w]:r_off[w + 1]].copy()) if nt[w] else (None, None)
                for w in range(len(nt))]

    def chain_frames(self, w):
        """Hidden GNSS-frame states of window w's IMUGNSSFactor chains, (n_frames, 16)."""
        n = i32()
        _check(lib().swgn_batch_get_chain_frames(self.h, w, C.byref(n), None), "get_chain_frames")
        out = np.zeros((max(n.value, 1), 16))
        _check(lib().swgn_batch_get_chain_frames(self.h, w, C.byref(n), _dp(out)), "get_chain_frames")
        return out[:n.value]

    def columns(self, w):
        n = i32()
        _check(lib().swgn_batch_get_columns(self.h, w, C.byref(n), None, None, None), "get_columns")
        b, o, s = (np.zeros(n.value, np.int32) for _ in range(3))
        _check(lib().swgn_batch_get_columns(self.h, w, C.byref(n), _ip(b), _ip(o), _ip(s)), "get_columns")
        return b, o, s

    def rows(self, w):
        n = i32()
        _check(lib().swgn_batch_get_rows(self.h, w, C.byref(n), None, None), "get_rows")
        f, o = (np.zeros(n.value, np.int32) for _ in range(2))
        _check(lib().swgn_batch_get_rows(self.h, w, C.byref(n), _ip(f), _ip(o)), "get_rows")
        return f, o

    def evaluate(self, w, n_res, n_cols):
        cost = f64()
        r = np.zeros(n_res)
        g = np.zeros(n_cols)
        _check(lib().swgn_batch_evaluate(self.h, w, C.byref(cost), _dp(r), _dp(g)), "evaluate")
        return cost.value, r, g

    def dense_jacobian(self, w, n_res, n_cols):
        J = np.zeros((n_res, n_cols))
        _check(lib().swgn_batch_get_dense_jacobian(self.h, w, _dp(J)), "dense_jacobian")
        return J

    def linear_solve(self, w, D, n_cols):
        x = np.zeros(n_cols)
        Dp = None
        if D is not None:
            D = np.ascontiguousarray(D, np.float64)
            Dp = _dp(D)
        _check(lib().swgn_batch_linear_solve(self.h, w, Dp, _dp(x)), "linear_solve")
        return x

    @staticmethod
    def pack_epochs(epochs):
        """Per-window (epoch_begin, obs_amb, obs_sysfreq) lists -> the CSR of CSRs swgn_batch_ambiguity_fix takes."""
        win, eb, oa, sf = [0], [np.zeros(1, np.int64)], [], []
        base = 0
        for (e, a, f) in epochs:
            e = np.asarray(e, np.int64)
            eb.append(e[1:] + base)
            base += int(e[-1])
            oa.append(np.asarray(a, np.int32))
            sf.append(np.asarray(f, np.int32))
            win.append(win[-1] + len(e) - 1)
        pad = np.zeros(1, np.int32)
        return (np.array(win, np.int32), np.concatenate(eb).astype(np.int32), np.concatenate(oa + [pad]), np.concatenate(sf + [pad]))

    def ambiguity_fix_all(self, n_tail, epochs, last_fix=None):
        """LambdaSearch decision of every window in two launches.  epochs: per window (epoch_begin, obs_amb, obs_sysfreq)
        as returned by SynthWindow.ambiguity_epochs(), or the result of pack_epochs().  Returns (results, dd_pairs
        (n, n_tail, 2), F (n, 2, n_tail))."""
        win, eb, oa, sf = epochs if (len(epochs) == 4 and isinstance(epochs[0], np.ndarray) and epochs[0].dtype == np.int32
                                     and len(epochs[0]) == self.n + 1) else self.pack_epochs(epochs)
        res = (FixResult * self.n)()
        pairs = np.zeros((self.n, n_tail, 2), np.int32)
        F = np.zeros((self.n, 2, n_tail))
        lf = None if last_fix is None else np.ascontiguousarray(last_fix, np.int32)
        _check(lib().swgn_batch_ambiguity_fix(self.h, n_tail, _ip(win), _ip(eb), _ip(oa), _ip(sf), None if lf is None else _ip(lf), res,
                                              _ip(pairs), _dp(F)), "swgn_batch_ambiguity_fix")
        return res, pairs, F

    def close(self):
        if self.h:
            lib().swgn_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def preintegrate_batch(sample_begin, samples, bias, noise, device=0):
    """IMU pre-integration of n factors on the device -> (records (n, SWGN_IMU_STRIDE), info (n,))."""
    sample_begin = np.ascontiguousarray(sample_begin, np.int32)
    n = len(sample_begin) - 1
    samples = np.ascontiguousarray(samples, np.float64)
    bias = np.ascontiguousarray(bias, np.float64)
    noise = np.ascontiguousarray(noise, np.float64)
    rec = np.zeros((n, 474))
    info = np.zeros(n, np.int32)
    _check(lib().swgn_preintegrate_batch(device, n, _ip(sample_begin), _dp(samples), _dp(bias), _dp(noise), _dp(rec), _ip(info)),
           "swgn_preintegrate_batch")
    return rec, info


def lambda_batch(ns, m, a, Q, device=0):
    ns = np.ascontiguousarray(ns, np.int32)
    a = np.ascontiguousarray(a, np.float64)
    Q = np.ascontiguousarray(Q, np.float64)
    F = np.zeros(int(ns.sum()) * m)
    s = np.zeros(len(ns) * m)
    info = np.zeros(len(ns), np.int32)
    _check(lib().swgn_lambda_batch(device, len(ns), _ip(ns), m, _dp(a), _dp(Q), _dp(F), _dp(s),
                                   _ip(info)), "swgn_lambda_batch")
    return F, s, info


def ambiguity_fix(A, y, epoch_begin, obs_amb, obs_sysfreq, last_fix=0, device=0):
    n = len(y)
    A = np.ascontiguousarray(A, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    eb = np.ascontiguousarray(epoch_begin, np.int32)
    oa = np.ascontiguousarray(obs_amb, np.int32)
    sf = np.ascontiguousarray(obs_sysfreq, np.int32)
    pairs = np.zeros(2 * max(n, 1) * max(len(eb) - 1, 1), np.int32)
    F = np.zeros(2 * max(n, 1) * max(len(eb) - 1, 1))
    res = FixResult()
    _check(lib().swgn_ambiguity_fix(device, n, _dp(A), _dp(y), len(eb) - 1, _ip(eb), _ip(oa), _ip(sf),
                                    last_fix, _ip(pairs), _dp(F), C.byref(res)), "swgn_ambiguity_fix")
    nb = res.n_dd
    return pairs[:2 * nb].reshape(nb, 2), F[:2 * nb].reshape(2, nb).T.copy(), res


class MarginalizeOutput(C.Structure):
    _fields_ = [("cap_keep", i32), ("cap_n", i32), ("n_keep", i32), ("n", i32), ("m", i32),
                ("keep_block", P(i32)), ("keep_idx", P(i32)), ("J0", P(f64)), ("r0", P(f64))]


def marginalize(graph_ps, drops, device=0):
    """swgn_marginalize: MarginalizationInfo::marginalize for arbitrary drop sets.  graph_ps: pointers to Graph; drops: one uint8
    array per graph.  Returns [(keep_block, keep_idx, J0, r0, m)]."""
    n = len(graph_ps)
    L = lib()
    L.swgn_marginalize.argtypes = [i32, i32, P(P(Graph)), P(P(C.c_uint8)), P(MarginalizeOutput)]
    outs = (MarginalizeOutput * n)()
    keep = []
    dr = []
    for w in range(n):
        g = graph_ps[w].contents
        cap_n = sum(6 if g.block_manifold[b] == 1 else g.block_size[b] for b in range(g.n_blocks))
        kb, ki, J, r = np.zeros(g.n_blocks, np.int32), np.zeros(g.n_blocks, np.int32), np.zeros(cap_n * cap_n), np.zeros(cap_n)
        keep.append((kb, ki, J, r))
        outs[w].cap_keep, outs[w].cap_n = g.n_blocks, cap_n
        outs[w].keep_block, outs[w].keep_idx, outs[w].J0, outs[w].r0 = _ip(kb), _ip(ki), _dp(J), _dp(r)
        dr.append(np.ascontiguousarray(drops[w], np.uint8))
    garr = (P(Graph) * n)(*graph_ps)
    darr = (P(C.c_uint8) * n)(*[d.ctypes.data_as(P(C.c_uint8)) for d in dr])
    _check(L.swgn_marginalize(device, n, garr, darr, outs), "swgn_marginalize")
    res = []
    for w in range(n):
        kb, ki, J, r = keep[w]
        nn, nk = outs[w].n, outs[w].n_keep
        res.append((kb[:nk].copy(), ki[:nk].copy(), J[:nn * nn].reshape(nn, nn).copy(), r[:nn].copy(), outs[w].m))
    return res
