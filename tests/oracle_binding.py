"""ctypes binding of oracle/liboracle.so and oracle/_ref/libref_gnss.so -- TEST INFRASTRUCTURE.
Imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
import swgn  # noqa: E402

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double
P = C.POINTER
_o = None
_r = None


def _dp(a):
    return a.ctypes.data_as(P(f64))


def _ip(a):
    return a.ctypes.data_as(P(i32))


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def oracle():
    global _o
    if _o is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [P(swgn.Graph), P(swgn.Options)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_dims.argtypes = [C.c_void_p, P(i32)]
        L.oracle_columns.argtypes = [C.c_void_p, P(i32), P(i32), P(i32)]
        L.oracle_rows.argtypes = [C.c_void_p, P(i32), P(i32)]
        L.oracle_evaluate.argtypes = [C.c_void_p, P(f64), P(f64), P(f64), P(f64)]
        L.oracle_linear_solve.argtypes = [C.c_void_p, P(f64), P(f64), P(f64), P(f64)]
        L.oracle_minimize.argtypes = [C.c_void_p, P(swgn.Summary)]
        L.oracle_get_state.argtypes = [C.c_void_p, P(f64)]
        L.oracle_set_state.argtypes = [C.c_void_p, P(f64)]
        L.oracle_num_iteration_records.argtypes = [C.c_void_p]
        L.oracle_preintegrate_batch.restype = C.c_int
        L.oracle_preintegrate_batch.argtypes = [i32, P(i32), P(f64), P(f64), P(f64), P(f64)]
        L.oracle_evaluate_cost.restype = C.c_int
        L.oracle_evaluate_cost.argtypes = [C.c_void_p, P(f64), P(f64)]
        L.oracle_chain_frames.restype = C.c_int
        L.oracle_chain_frames.argtypes = [C.c_void_p, P(f64)]
        L.oracle_iteration_records.argtypes = [C.c_void_p, P(f64), P(f64), P(i32)]
        L.oracle_get_exports.argtypes = [C.c_void_p, P(f64), P(f64), P(f64)]
        L.oracle_tail_information.argtypes = [P(f64), i32, i32, P(f64)]
        L.oracle_update_schur.argtypes = [P(f64), P(f64), i32, i32, P(f64), P(f64)]
        L.oracle_schur_raw.argtypes = [i32, P(i32), i32, P(i32), P(i32), P(i32), P(f64), P(f64),
                                       P(f64), i32, P(f64), P(f64), P(f64)]
        L.oracle_lambda.argtypes = [i32, i32, P(f64), P(f64), P(f64), P(f64)]
        L.oracle_matinv.argtypes = [P(f64), i32]
        L.oracle_invert_psd.argtypes = [P(f64), i32, P(f64)]
        L.oracle_ambiguity_fix.argtypes = [i32, P(f64), P(f64), i32, P(i32), P(i32), P(i32), i32,
                                           P(i32), P(f64), P(swgn.FixResult)]
        L.oracle_distance.restype = f64
        L.oracle_distance.argtypes = [P(f64), P(f64), P(f64)]
        L.oracle_velocity_distance.restype = f64
        L.oracle_velocity_distance.argtypes = [P(f64)] * 5
        L.oracle_varerr2.restype = f64
        L.oracle_varerr2.argtypes = [f64, f64, f64]
        L.oracle_pose_plus.argtypes = [P(f64), P(f64), P(f64)]
        L.oracle_cauchy.argtypes = [f64, f64, P(f64)]
        L.oracle_factor_eval.argtypes = [i32, i32, P(f64), P(f64), P(f64), P(f64), P(f64)]
        L.oracle_solve_batch_timed.restype = f64
        L.oracle_solve_batch_timed.argtypes = [i32, P(P(swgn.Graph)), P(swgn.Options), i32, P(i64),
                                               P(f64), i64]
        _o = L
    return _o


def ref():
    """The reference's own lambda.cpp / common_function.cpp compiled by oracle/build_ref.sh.
    Returns None when it has not been built (it cannot be built on the GPU box)."""
    global _r
    if _r is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libref_gnss.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_lambda.argtypes = [i32, i32, P(f64), P(f64), P(f64), P(f64)]
        L.ref_matinv.argtypes = [P(f64), i32]
        L.ref_distance.restype = f64
        L.ref_distance.argtypes = [P(f64), P(f64), P(f64)]
        L.ref_velecitydistance.restype = f64
        L.ref_velecitydistance.argtypes = [P(f64)] * 5
        L.ref_dot.restype = f64
        L.ref_dot.argtypes = [P(f64), P(f64), i32]
        _r = L
    return _r


class OracleSolver:
    def __init__(self, graph_p, options):
        L = oracle()
        self.h = L.oracle_create(graph_p, C.byref(options))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        d = (i32 * 8)()
        L.oracle_dims(self.h, d)
        (self.n_res, self.n_cols, self.n_col_blocks, self.n_e_blocks, self.n_row_blocks, self.n_f,
         self.n_e, self.n_state) = list(d)

    def columns(self):
        b, o, s = (np.zeros(self.n_col_blocks, np.int32) for _ in range(3))
        oracle().oracle_columns(self.h, _ip(b), _ip(o), _ip(s))
        return b, o, s

    def rows(self):
        f, o = (np.zeros(self.n_row_blocks, np.int32) for _ in range(2))
        oracle().oracle_rows(self.h, _ip(f), _ip(o))
        return f, o

    def evaluate(self, jac=True):
        cost = f64()
        r = np.zeros(self.n_res)
        g = np.zeros(self.n_cols)
        J = np.zeros((self.n_res, self.n_cols)) if jac else None
        st = oracle().oracle_evaluate(self.h, C.byref(cost), _dp(r), _dp(g), _dp(J) if jac else None)
        assert st == 0
        return cost.value, r, g, J

    def evaluate_cost(self):
        """Cost-only evaluation (no Jacobians requested from the cost functions)."""
        cost = f64()
        r = np.zeros(self.n_res)
        st = oracle().oracle_evaluate_cost(self.h, C.byref(cost), _dp(r))
        assert st == 0
        return cost.value, r

    def linear_solve(self, D=None):
        x = np.zeros(self.n_cols)
        S = np.zeros((self.n_f, self.n_f))
        rhs = np.zeros(self.n_f)
        Dp = None
        if D is not None:
            D = np.ascontiguousarray(D, np.float64)
            Dp = _dp(D)
        st = oracle().oracle_linear_solve(self.h, Dp, _dp(x), _dp(S), _dp(rhs))
        return st, x, S, rhs

    def minimize(self):
        sm = swgn.Summary()
        st = oracle().oracle_minimize(self.h, C.byref(sm))
        return st, sm

    def state(self):
        x = np.zeros(self.n_state)
        oracle().oracle_get_state(self.h, _dp(x))
        return x

    def set_state(self, x):
        x = np.ascontiguousarray(x, np.float64)
        oracle().oracle_set_state(self.h, _dp(x))

    def chain_frames(self):
        """Current hidden GNSS-frame states of the IMUGNSSFactor chains, (n_frames, 16)."""
        n = oracle().oracle_chain_frames(self.h, None)
        out = np.zeros((max(n, 1), 16))
        oracle().oracle_chain_frames(self.h, _dp(out))
        return out[:n]

    def iteration_records(self):
        n = oracle().oracle_num_iteration_records(self.h)
        c, r = np.zeros(n), np.zeros(n)
        s = np.zeros(n, np.int32)
        oracle().oracle_iteration_records(self.h, _dp(c), _dp(r), _ip(s))
        return c, r, s

    def exports(self):
        S = np.zeros((self.n_f, self.n_f))
        r = np.zeros(self.n_f)
        Lm = np.zeros((self.n_f, self.n_f))
        n = oracle().oracle_get_exports(self.h, _dp(S), _dp(r), _dp(Lm))
        return n, S, r, Lm

    def __del__(self):
        try:
            if self.h:
                oracle().oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass


def tail_information(Lm, n_tail):
    Lm = np.ascontiguousarray(Lm)
    A = np.zeros((n_tail, n_tail))
    oracle().oracle_tail_information(_dp(Lm), Lm.shape[0], n_tail, _dp(A))
    return A


def preintegrate_batch(sample_begin, samples, bias, noise):
    sample_begin = np.ascontiguousarray(sample_begin, np.int32)
    n = len(sample_begin) - 1
    samples = np.ascontiguousarray(samples, np.float64)
    bias = np.ascontiguousarray(bias, np.float64)
    noise = np.ascontiguousarray(noise, np.float64)
    rec = np.zeros((n, 474))
    bad = oracle().oracle_preintegrate_batch(n, _ip(sample_begin), _dp(samples), _dp(bias), _dp(noise), _dp(rec))
    return rec, bad


def prior_sqrt(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    n = A.shape[0]
    J0, r0 = np.zeros((n, n)), np.zeros(n)
    L = oracle()
    L.oracle_prior_sqrt.argtypes = [P(f64), P(f64), i32, P(f64), P(f64)]
    L.oracle_prior_sqrt(_dp(A), _dp(b), n, _dp(J0), _dp(r0))
    return J0, r0


def update_schur(S, r, n_tail):
    S = np.ascontiguousarray(S)
    r = np.ascontiguousarray(r)
    A = np.zeros((n_tail, n_tail))
    b = np.zeros(n_tail)
    oracle().oracle_update_schur(_dp(S), _dp(r), S.shape[0], n_tail, _dp(A), _dp(b))
    return A, b


def lambda_search(a, Q, m=2, which="oracle"):
    """Q is passed column-major as the reference expects (symmetric in practice)."""
    n = len(a)
    a = np.ascontiguousarray(a, np.float64)
    Qc = np.asfortranarray(Q, np.float64)
    F = np.zeros(n * m)
    s = np.zeros(m)
    if which == "oracle":
        info = oracle().oracle_lambda(n, m, _dp(a), Qc.ctypes.data_as(P(f64)), _dp(F), _dp(s))
    else:
        info = ref().ref_lambda(n, m, _dp(a), Qc.ctypes.data_as(P(f64)), _dp(F), _dp(s))
    return info, F.reshape(m, n).T.copy(), s


def ambiguity_fix(A, y, epoch_begin, obs_amb, obs_sysfreq, last_fix=0):
    n = len(y)
    A = np.ascontiguousarray(A, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    eb = np.ascontiguousarray(epoch_begin, np.int32)
    oa = np.ascontiguousarray(obs_amb, np.int32)
    sf = np.ascontiguousarray(obs_sysfreq, np.int32)
    pairs = np.zeros(2 * max(n, 1) * max(len(eb) - 1, 1), np.int32)
    F = np.zeros(2 * max(n, 1) * max(len(eb) - 1, 1))
    res = swgn.FixResult()
    oracle().oracle_ambiguity_fix(n, _dp(A), _dp(y), len(eb) - 1, _ip(eb), _ip(oa), _ip(sf), last_fix,
                                  _ip(pairs), _dp(F), C.byref(res))
    nb = res.n_dd
    return pairs[:2 * nb].reshape(nb, 2), F[:2 * nb].reshape(2, nb).T.copy(), res


# ---- per-epoch GNSS preprocessing (oracle/oracle_gnss_epoch.cpp) ---------------------------------------------------
class OracleGnssTracker:
    def __init__(self, cfg):
        import swgn_gnss as G
        self.G = G
        L = oracle()
        L.oracle_gnss_tracker_new.restype = C.c_void_p
        L.oracle_gnss_tracker_new.argtypes = [P(G.Config)]
        L.oracle_gnss_tracker_free.argtypes = [C.c_void_p]
        L.oracle_gnss_tracker_count.argtypes = [C.c_void_p, i32]
        L.oracle_gnss_tracker_get.argtypes = [C.c_void_p, i32, i32, P(G.Ambiguity)]
        L.oracle_gnss_tracker_set_value.argtypes = [C.c_void_p, i32, i32, f64]
        L.oracle_gnss_preprocess.argtypes = [C.c_void_p, P(G.Epoch), P(G.Frame), P(G.Output)]
        self.h = L.oracle_gnss_tracker_new(C.byref(cfg))

    def count(self, family):
        return oracle().oracle_gnss_tracker_count(self.h, family)

    def get(self, family, handle):
        a = self.G.Ambiguity()
        assert oracle().oracle_gnss_tracker_get(self.h, family, handle, C.byref(a)) == 0
        return a

    def set_value(self, family, handle, v):
        oracle().oracle_gnss_tracker_set_value(self.h, family, handle, v)

    def preprocess(self, epoch, frame, out=None):
        out = out or self.G.OutputBuffers()
        rc = oracle().oracle_gnss_preprocess(self.h, C.byref(epoch), C.byref(frame), C.byref(out.c))
        assert rc == 0, rc
        return out

    def __del__(self):
        try:
            oracle().oracle_gnss_tracker_free(self.h)
        except Exception:
            pass


def update_azel(globalxyz, epoch):
    import swgn_gnss as G
    L = oracle()
    L.oracle_update_azel.argtypes = [P(f64), P(G.Epoch)]
    L.oracle_update_azel.restype = None
    x = np.ascontiguousarray(globalxyz, np.float64)
    L.oracle_update_azel(_dp(x), C.byref(epoch))
