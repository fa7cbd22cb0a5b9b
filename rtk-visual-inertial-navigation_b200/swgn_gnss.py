"""ctypes binding of include/swgn_gnss.h (per-epoch GNSS linearisation).  Harness-side only, like
swgn.py: the structs mirror the C ABI; the product entry points are reached through libswgn.so and fail loudly when it
is missing (there is no CPU path for the numerical parts)."""
import ctypes as C

import numpy as np

from swgn import Summary, _check, lib

i32, f64, u8 = C.c_int32, C.c_double, C.c_uint8
P = C.POINTER
NFREQ, MAXOBS, MAXSAT, NCLK = 2, 64, 107, 13
AMB_RTK, AMB_SPP, AMB_PCORR = 0, 1, 2
KEEP_POSE, KEEP_SPEED_BIAS, KEEP_BLACK, KEEP_AMB_RTK, KEEP_AMB_SPP, KEEP_AMB_PCORR = range(6)


class Obs(C.Structure):
    _fields_ = [
        ("sat", u8), ("sys", u8), ("svh", u8), ("pad0_", u8),
        ("rtk_slip_count", u8 * NFREQ), ("spp_slip_count", u8 * NFREQ), ("half_flag", u8 * NFREQ), ("pad1_", u8 * 6),
        ("spp_p", f64 * NFREQ), ("spp_l", f64 * NFREQ), ("spp_d", f64 * NFREQ),
        ("spp_lstd", f64 * NFREQ), ("spp_pstd", f64 * NFREQ), ("spp_dstd", f64 * NFREQ),
        ("rtk_p", f64 * NFREQ), ("rtk_l", f64 * NFREQ), ("rtk_pstd", f64 * NFREQ), ("rtk_lstd", f64 * NFREQ),
        ("spp_p0", f64 * NFREQ),
        ("sat_pos", f64 * 3), ("sat_vel", f64 * 3),
        ("el", f64),
        ("sat_var", f64), ("ion_var", f64), ("trop_var", f64),
        ("rtk_n", i32 * NFREQ), ("spp_n", i32 * NFREQ), ("pcorr_n", i32 * NFREQ),
    ]


class Epoch(C.Structure):
    _fields_ = [("n_obs", i32), ("pad_", i32), ("ros_time", f64), ("base_xyz", f64 * 3), ("br_time_diff", f64), ("obs", P(Obs))]


class Config(C.Structure):
    _fields_ = [
        ("use_imu", i32), ("use_rtk", i32), ("use_rtd", i32), ("use_spp_phase", i32), ("use_spp_correction", i32),
        ("use_doppler", i32), ("phase_all_reset_count", i32), ("estimate_pcorrection_period", i32),
        ("azelmin", f64), ("lams", (f64 * NFREQ) * 3), ("ambiguity_timeout", f64), ("slip_fraction_rtk", f64),
        ("init_max_iterations", i32), ("init_constant_after", i32), ("init_radius", f64), ("device", i32), ("pad_", i32),
    ]


class Frame(C.Structure):
    _fields_ = [("pose", f64 * 7), ("speed_bias", f64 * 9), ("gnss_dt", f64 * NCLK), ("blackvalue", f64),
                ("nonlinear", i32), ("rover_count", i32), ("epochs_since_start", i32), ("not_fix_count", i32)]


class Ambiguity(C.Structure):
    _fields_ = [("value", f64), ("last_update_time", f64), ("continue_count", i32),
                ("slip_count", u8), ("half_flag", u8), ("sys", u8), ("f", u8), ("sat", i32), ("alive", i32)]


class Output(C.Structure):
    _fields_ = [("cap_keep", i32), ("cap_n", i32), ("n_keep", i32), ("n", i32),
                ("keep_kind", P(i32)), ("keep_handle", P(i32)), ("keep_idx", P(i32)),
                ("x0", P(f64)), ("J0", P(f64)), ("r0", P(f64)),
                ("n_new", i32 * 3), ("n_slip_rtk", i32), ("n_slip_spp", i32), ("n_factors", i32), ("init_summary", Summary)]


def _ip(a):
    return a.ctypes.data_as(P(i32))


def _dp(a):
    return a.ctypes.data_as(P(f64))


class OutputBuffers:
    """Caller-owned buffers of one swgn_gnss_output."""

    def __init__(self, cap_keep=3 + 3 * NFREQ * MAXOBS, cap_n=16 + 3 * NFREQ * MAXOBS):
        self.keep_kind = np.zeros(cap_keep, np.int32)
        self.keep_handle = np.zeros(cap_keep, np.int32)
        self.keep_idx = np.zeros(cap_keep, np.int32)
        self.x0 = np.zeros(cap_n + 3)
        self.J0 = np.zeros(cap_n * cap_n)
        self.r0 = np.zeros(cap_n)
        self.c = Output()
        self.c.cap_keep, self.c.cap_n = cap_keep, cap_n
        self.c.keep_kind, self.c.keep_handle, self.c.keep_idx = _ip(self.keep_kind), _ip(self.keep_handle), _ip(self.keep_idx)
        self.c.x0, self.c.J0, self.c.r0 = _dp(self.x0), _dp(self.J0), _dp(self.r0)

    def prior(self):
        """(keep list [(kind, handle, first column)], x0, J0 (n, n), r0 (n,))"""
        n, nk = self.c.n, self.c.n_keep
        keep = [(int(self.keep_kind[k]), int(self.keep_handle[k]), int(self.keep_idx[k])) for k in range(nk)]
        nx = sum(7 if k[0] == KEEP_POSE else 9 if k[0] == KEEP_SPEED_BIAS else 1 for k in keep)
        return keep, self.x0[:nx].copy(), self.J0[:n * n].reshape(n, n).copy(), self.r0[:n].copy()


def default_config():
    c = Config()
    L = lib()
    L.swgn_gnss_config_default.argtypes = [P(Config)]
    L.swgn_gnss_config_default.restype = None
    L.swgn_gnss_config_default(C.byref(c))
    return c


def _proto():
    L = lib()
    if getattr(L, "_gnss_proto", False):
        return L
    L.swgn_gnss_tracker_create.argtypes = [P(Config), P(C.c_void_p)]
    L.swgn_gnss_tracker_destroy.argtypes = [C.c_void_p]
    L.swgn_gnss_tracker_destroy.restype = None
    L.swgn_gnss_tracker_count.argtypes = [C.c_void_p, i32]
    L.swgn_gnss_tracker_get.argtypes = [C.c_void_p, i32, i32, P(Ambiguity)]
    L.swgn_gnss_tracker_set_value.argtypes = [C.c_void_p, i32, i32, f64]
    L.swgn_gnss_tracker_erase.argtypes = [C.c_void_p, i32, i32]
    L.swgn_gnss_preprocess.argtypes = [i32, P(C.c_void_p), P(P(Epoch)), P(Frame), P(Output)]
    L.swgn_gnss_gate_residuals.argtypes = [i32, P(f64), P(f64), i32]
    L.swgn_gnss_epoch_records.argtypes = [C.c_void_p, P(Epoch), P(Frame), P(i32), P(i32), P(i32), P(f64), P(i32), P(i32), P(i32),
                                          P(i32), P(i32)]
    L._gnss_proto = True
    return L


class Tracker:
    """One receiver's ambiguity lists (swgn_gnss_tracker)."""

    def __init__(self, cfg):
        self.h = C.c_void_p()
        _check(_proto().swgn_gnss_tracker_create(C.byref(cfg), C.byref(self.h)), "swgn_gnss_tracker_create")

    def count(self, family):
        return _proto().swgn_gnss_tracker_count(self.h, family)

    def get(self, family, handle):
        a = Ambiguity()
        _check(_proto().swgn_gnss_tracker_get(self.h, family, handle, C.byref(a)), "swgn_gnss_tracker_get")
        return a

    def set_value(self, family, handle, value):
        _check(_proto().swgn_gnss_tracker_set_value(self.h, family, handle, value), "swgn_gnss_tracker_set_value")

    def erase(self, family, handle):
        _check(_proto().swgn_gnss_tracker_erase(self.h, family, handle), "swgn_gnss_tracker_erase")

    def records(self, epoch, frame):
        """AddGnssResidual of a preprocessed epoch: (kind, blocks (n, 3), data (n, 16), clk slots, [(family, handle)])."""
        cap = 5 * max(epoch.n_obs, 1)
        kind, blocks, data = np.zeros(cap, np.int32), np.zeros(3 * cap, np.int32), np.zeros(16 * cap)
        clk, fam, han = np.zeros(NCLK, np.int32), np.zeros(3 * NFREQ * MAXOBS, np.int32), np.zeros(3 * NFREQ * MAXOBS, np.int32)
        n, nclk, namb = i32(), i32(), i32()
        _check(_proto().swgn_gnss_epoch_records(self.h, C.byref(epoch), C.byref(frame), C.byref(n), _ip(kind), _ip(blocks), _dp(data),
                                                C.byref(nclk), _ip(clk), C.byref(namb), _ip(fam), _ip(han)), "swgn_gnss_epoch_records")
        n, nclk, namb = n.value, nclk.value, namb.value
        return kind[:n], blocks[:3 * n].reshape(n, 3), data[:16 * n].reshape(n, 16), clk[:nclk], list(zip(fam[:namb], han[:namb]))

    def close(self):
        if self.h:
            _proto().swgn_gnss_tracker_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def preprocess(trackers, epochs, frames, outputs=None):
    """swgn_gnss_preprocess for len(trackers) receivers; epochs: list of Epoch, frames: list of Frame (both modified in
    place).  Returns the list of OutputBuffers."""
    n = len(trackers)
    outputs = outputs or [OutputBuffers() for _ in range(n)]
    th = (C.c_void_p * n)(*[t.h for t in trackers])
    ep = (P(Epoch) * n)(*[C.pointer(e) for e in epochs])
    fr = (Frame * n)(*frames)
    oc = (Output * n)(*[o.c for o in outputs])
    _check(_proto().swgn_gnss_preprocess(n, th, ep, fr, oc), "swgn_gnss_preprocess")
    for i in range(n):
        C.memmove(C.byref(frames[i]), C.byref(fr[i]), C.sizeof(Frame))
        C.memmove(C.byref(outputs[i].c), C.byref(oc[i]), C.sizeof(Output))
    return outputs


def gate_residuals(rec, device=0):
    rec = np.ascontiguousarray(rec, np.float64).reshape(-1, 16)
    out = np.zeros((len(rec), 3))
    _check(_proto().swgn_gnss_gate_residuals(len(rec), _dp(rec), _dp(out), device), "swgn_gnss_gate_residuals")
    return out


class FixedIntegerJob(C.Structure):
    """swgn_fixed_integer_job (include/swgn.h)"""
    _fields_ = [("n_keep", i32), ("n", i32), ("keep_size", P(i32)), ("keep_idx", P(i32)), ("x0", P(f64)), ("J0", P(f64)), ("r0", P(f64)),
                ("x", P(f64)), ("n_dd", i32), ("dd_keep", P(i32)), ("F", P(f64)), ("dd_sysfreq", P(i32)), ("istd", f64),
                ("J0_out", P(f64)), ("r0_out", P(f64))]


class FixedIntegerArrays:
    """numpy storage of one job; .c is the struct pointing into it."""

    def __init__(self, keep_size, keep_idx, x0, J0, r0, x, dd_keep, F, dd_sysfreq, istd=1 / 0.03):
        self.keep_size = np.ascontiguousarray(keep_size, np.int32)
        self.keep_idx = np.ascontiguousarray(keep_idx, np.int32)
        self.x0, self.J0, self.r0 = (np.ascontiguousarray(a, np.float64) for a in (x0, J0, r0))
        self.x = np.ascontiguousarray(x, np.float64)
        self.dd_keep = np.ascontiguousarray(dd_keep, np.int32).reshape(-1)
        self.F = np.ascontiguousarray(F, np.float64)
        self.dd_sysfreq = np.ascontiguousarray(dd_sysfreq, np.int32)
        n = len(self.r0)
        self.J0_out, self.r0_out = np.zeros((n, n)), np.zeros(n)
        c = FixedIntegerJob()
        c.n_keep, c.n, c.n_dd, c.istd = len(self.keep_size), n, len(self.F), istd
        c.keep_size, c.keep_idx, c.dd_keep, c.dd_sysfreq = _ip(self.keep_size), _ip(self.keep_idx), _ip(self.dd_keep), _ip(self.dd_sysfreq)
        c.x0, c.J0, c.r0, c.x, c.F = _dp(self.x0), _dp(self.J0), _dp(self.r0), _dp(self.x), _dp(self.F)
        c.J0_out, c.r0_out = _dp(self.J0_out), _dp(self.r0_out)
        self.c = c


def fixed_integer_prior(jobs, device=0):
    """swgn_fixed_integer_prior for a list of FixedIntegerArrays; results in job.J0_out / job.r0_out."""
    L = lib()
    L.swgn_fixed_integer_prior.argtypes = [i32, i32, P(FixedIntegerJob)]
    arr = (FixedIntegerJob * len(jobs))(*[j.c for j in jobs])
    _check(L.swgn_fixed_integer_prior(device, len(jobs), arr), "swgn_fixed_integer_prior")
