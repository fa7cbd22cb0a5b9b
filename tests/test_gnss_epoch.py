"""Per-epoch GNSS linearisation (include/swgn_gnss.h): GnssPreprocess / AddGnssResidual /
MarginalizationInfo::marginalize (RVI/swf/swf_gnss.cpp:265-587, swf_core.cpp:87-205, marginalization_factor.cpp:260-377).

CPU part: the oracle restatement against (a) the reference's own update_azel compiled into oracle/_ref, (b) an independent
numpy assembly of the epoch's normal equations and Schur complement, (c) the bookkeeping expectations of a scripted
scenario (announced slips, an unannounced jump, an elevation mask, an unhealthy satellite, a data gap).
GPU part: the product (device elevations / gating residuals, export-mode pass, LEVENBERG_MARQUARDT pass through the C ABI)
against the oracle, epoch by epoch over the same scenario, for several receivers in one call."""
import ctypes as C
import os

import numpy as np
import pytest

import gnss_scenario as S
import oracle_binding as ob
import swgn_gnss as G

CLIGHT = 299792458.0
_libm = C.CDLL("libm.so.6")
_libm.sinf.restype = C.c_float
_libm.sinf.argtypes = [C.c_float]


def run_oracle(sc, cfg, n_epochs, hook=None):
    T = ob.OracleGnssTracker(cfg)
    dt, black = np.zeros(G.NCLK), 0.02
    log = []
    for k in range(n_epochs):
        e, obs, f = sc.epoch(k)
        for c in range(G.NCLK):
            f.gnss_dt[c] = dt[c]
        f.blackvalue = black
        if hook:
            hook(k, e, f)
        raw = S.copy_epoch(e, obs)
        f_in = S.copy_frame(f)
        out = T.preprocess(e, f)
        dt, black = np.array(f.gnss_dt[:]), f.blackvalue
        log.append(dict(epoch=e, obs=obs, frame=f, frame_in=f_in, raw=raw, out=out))
    return T, log


def test_update_azel_matches_the_reference_build():
    L = ob.ref()
    if L is None:
        pytest.skip("oracle/_ref not built")
    L.ref_update_azel.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_ubyte), C.POINTER(C.c_double)]
    cfg = G.default_config()
    sc = S.Scenario(3, cfg=cfg)
    for k in (0, 5):
        e, obs, f = sc.epoch(k)
        xyz = np.array([f.pose[c] + e.base_xyz[c] for c in range(3)])
        sat = np.array([[obs[i].sat_pos[c] for c in range(3)] for i in range(e.n_obs)]).ravel()
        svh = np.array([obs[i].svh for i in range(e.n_obs)], np.uint8)
        el = np.full(e.n_obs, -7.0)
        for i in range(e.n_obs):
            obs[i].el = -7.0
        L.ref_update_azel(xyz.ctypes.data_as(C.POINTER(C.c_double)), e.n_obs, sat.ctypes.data_as(C.POINTER(C.c_double)),
                          svh.ctypes.data_as(C.POINTER(C.c_ubyte)), el.ctypes.data_as(C.POINTER(C.c_double)))
        ob.update_azel(xyz, e)
        mine = np.array([obs[i].el for i in range(e.n_obs)])
        assert np.array_equal(mine, el)          # same libm, same operation order: bit-exact
        assert (el[svh != 0] == -7.0).all()       # unhealthy satellites are skipped
        up = el[svh == 0]
        assert up.min() > np.radians(11.0) and up.max() < np.radians(86.0)


def numpy_normal_equations(cfg, e, obs, f):
    """Independent assembly of A = sum J'J, b = sum J'r of the epoch's factors (RTK carrier, RTK pseudorange, Doppler,
    InitialBlackFactor; the default configuration) at (pose, speed-bias, black, N = 0, clocks), in the order
    clocks (slot order) | pose 6 | speed-bias 9 | black | ambiguities (observation order), and its Schur complement."""
    lams = [[cfg.lams[s][q] for q in range(2)] for s in range(3)]
    pos = np.array(f.pose[:3]) + np.array(e.base_xyz[:])
    rows = []   # (dict column -> value, residual)
    amb = []
    for i in range(e.n_obs):
        o = obs[i]
        if o.rtk_n[0] >= 0 and o.el >= cfg.azelmin:
            amb.append(o.rtk_n[0])

    def weight(el, var):
        b = CLIGHT * 5e-12 * e.br_time_diff
        s = float(_libm.sinf(C.c_float(el)))   # the reference's single-precision sinf, gnss_factor.cpp:100
        return 1.0 / np.sqrt(var / s / s + b * b)
    for i in range(e.n_obs):
        o = obs[i]
        sat = np.array(o.sat_pos[:])
        d = pos - sat
        rho = np.linalg.norm(d)
        u = d / rho
        r1 = rho + S.OMGE * (sat[0] * pos[1] - sat[1] * pos[0]) / CLIGHT
        lam = lams[o.sys][0]
        if o.el < cfg.azelmin:
            continue
        if o.rtk_n[0] >= 0:
            w = weight(o.el, (o.rtk_lstd[0] * lam) ** 2)
            rows.append(({("p", 0): w * u[0], ("p", 1): w * u[1], ("p", 2): w * u[2], ("n", o.rtk_n[0]): -w * lam, ("c", o.sys * 2): w},
                         w * (r1 - 0.0 * lam - o.rtk_l[0] * lam + f.gnss_dt[o.sys * 2])))
        if o.rtk_p[0] != 0 and o.svh == 0 and o.rtk_pstd[0] <= 2:
            w = weight(o.el, o.rtk_pstd[0] ** 2)
            rows.append(({("p", 0): w * u[0], ("p", 1): w * u[1], ("p", 2): w * u[2], ("c", o.sys * 2): w},
                         w * (r1 - o.rtk_p[0] + f.gnss_dt[o.sys * 2])))
    for i in range(e.n_obs):
        o = obs[i]
        if o.spp_d[0] == 0 or o.svh != 0 or o.spp_dstd[0] > 2 or o.el < cfg.azelmin:
            continue
        lam = lams[o.sys][0]
        w = np.sin(o.el) ** 2 / (o.spp_dstd[0] * lam)
        sat, vs = np.array(o.sat_pos[:]), np.array(o.sat_vel[:])
        vr = np.array(f.speed_bias[:3])
        d = pos - sat
        rho = np.linalg.norm(d)
        u = d / rho
        rate = S.range_rate(pos, sat, vr, vs)
        jp = w * (np.eye(3) - np.outer(u, u)) @ (vr - vs) / rho
        rows.append(({("v", 0): w * u[0], ("v", 1): w * u[1], ("v", 2): w * u[2], ("c", 12): w, ("p", 0): jp[0], ("p", 1): jp[1], ("p", 2): jp[2]},
                     w * (rate + f.gnss_dt[12] + o.spp_d[0] * lam)))
    rows.append(({("b", 0): 1.0}, f.blackvalue * 1.0))
    clk = sorted({k[1] for r, _ in rows for k in r if k[0] == "c"})
    col = {}
    for s in clk:
        col[("c", s)] = len(col)
    m = len(col)
    for c in range(6):
        col[("p", c)] = m + c
    for c in range(9):
        col[("v", c)] = m + 6 + c
    col[("b", 0)] = m + 15
    for a in amb:
        col[("n", a)] = len(col) if ("n", a) not in col else col[("n", a)]
    N = m + 16 + len(amb)
    J = np.zeros((len(rows), N))
    r = np.zeros(len(rows))
    for k, (jr, rr) in enumerate(rows):
        for key, v in jr.items():
            J[k, col[key]] = v
        r[k] = rr
    A, b = J.T @ J, J.T @ r
    Amm, Amr, Arr = A[:m, :m], A[:m, m:], A[m:, m:]
    As = Arr - Amr.T @ np.linalg.solve(Amm, Amr)
    bs = b[m:] - Amr.T @ np.linalg.solve(Amm, b[:m])
    return As, bs, amb, len(rows)


def test_oracle_prior_is_the_schur_complement_of_the_epoch():
    cfg = G.default_config()
    sc = S.Scenario(1, cfg=cfg)
    T, log = run_oracle(sc, cfg, 3)
    for rec in log:
        out, e, obs = rec["out"], rec["epoch"], rec["obs"]
        keep, x0, J0, r0 = out.prior()
        As, bs, amb, nrows = numpy_normal_equations(cfg, e, obs, rec["frame_in"])
        assert out.c.n_factors == nrows
        assert [k[0] for k in keep[:3]] == [G.KEEP_POSE, G.KEEP_SPEED_BIAS, G.KEEP_BLACK]
        assert [k[1] for k in keep[3:]] == amb and all(k[0] == G.KEEP_AMB_RTK for k in keep[3:])
        assert [k[2] for k in keep] == [0, 6, 15] + list(range(16, 16 + len(amb)))
        scale = np.abs(As).max()
        assert np.abs(J0.T @ J0 - As).max() < 1e-9 * scale
        assert np.abs(J0.T @ r0 - bs).max() < 1e-9 * np.abs(bs).max()
        # linearisation point: pose, speed-bias, black as given, ambiguities at zero (PhaseBiasSaveAndReset)
        assert np.array_equal(x0[:7], np.array(rec["frame_in"].pose[:]))
        assert np.array_equal(x0[7:16], np.array(rec["frame_in"].speed_bias[:]))
        assert (x0[17:] == 0).all()
        # rank: rotation (3) and the IMU biases (6) are unobserved by GNSS
        assert np.linalg.matrix_rank(J0, tol=1e-6) == J0.shape[0] - 9


def test_oracle_bookkeeping_on_the_scripted_scenario():
    cfg = G.default_config()
    sc = S.Scenario(0, cfg=cfg)
    T, log = run_oracle(sc, cfg, 14)
    new = [tuple(r["out"].c.n_new[:]) for r in log]
    slips = [r["out"].c.n_slip_rtk for r in log]
    # epoch 0: 20 satellites, one below the mask, one unhealthy -> 18 ambiguities of each phase family (the SPP list is
    # maintained whether or not USE_SPP_PHASE is set, as in the reference)
    assert new[0] == (18, 18, 0) and log[0]["epoch"].n_obs == 20
    assert new[1] == new[2] == new[3] == (0, 0, 0)
    assert new[4] == (1, 0, 0) and slips[4] == 0            # announced slip: new RTK ambiguity, no gate hit
    assert slips[6] == 1 and new[6] == (1, 1, 0)            # unannounced 3-cycle jump: caught by the median gate; resets SPP too
    assert new[8][0] >= 1                                   # second announced slip
    assert new[9] == (18, 18, 0)                            # 13 s gap > ambiguity_timeout: everything starts again
    assert all(n == (0, 0, 0) for n in new[10:])
    # counters: an ambiguity tracked since epoch 9 has been counted 5 times at the end
    a = T.get(G.AMB_RTK, T.count(G.AMB_RTK) - 1)
    assert a.continue_count == 5 and a.last_update_time == log[-1]["epoch"].ros_time
    # masked satellite: phase measurements zeroed, no handle; unhealthy satellite: untouched, no handle
    e, obs = log[0]["epoch"], log[0]["obs"]
    low = [i for i in range(e.n_obs) if obs[i].svh == 0 and obs[i].el < cfg.azelmin]
    assert len(low) == 1 and obs[low[0]].rtk_l[0] == 0 and obs[low[0]].rtk_n[0] == -1
    bad = [i for i in range(e.n_obs) if obs[i].svh]
    assert len(bad) == 1 and obs[bad[0]].rtk_l[0] != 0 and obs[bad[0]].rtk_n[0] == -1
    # the initialisation solve puts every residual near zero: float ambiguities absorb the range within the
    # pseudorange noise, clocks within a metre of the truth
    t_last = log[-1]["epoch"].ros_time - 1000.0
    for s in (0, 2, 4):
        assert abs(log[-1]["frame"].gnss_dt[s] - (sc.clk[s] + sc.clk_rate[s] * t_last)) < 1.0
    for r in log:
        s = r["out"].c.init_summary
        assert s.termination_type in (0, 1) and s.final_cost < 50 and s.num_iterations <= 2


def test_reset_of_all_phase_biases_and_constant_old_ambiguities():
    cfg = G.default_config()
    sc = S.Scenario(2, cfg=cfg)

    def hook(k, e, f):
        if k == 12:
            f.not_fix_count = cfg.phase_all_reset_count + 1
    T, log = run_oracle(sc, cfg, 14, hook)
    assert log[12]["out"].c.n_new[0] == 18          # not_fix_count above Phase_ALL_RESET_COUNT: every RTK ambiguity restarts
    # an ambiguity older than init_constant_after epochs is not moved by the initialisation solve
    cfg2 = G.default_config()
    cfg2.init_constant_after = 2
    sc2 = S.Scenario(2, cfg=cfg2)
    T2 = ob.OracleGnssTracker(cfg2)
    dt = np.zeros(G.NCLK)
    before = None
    for k in range(5):
        e, obs, f = sc2.epoch(k)
        for c in range(G.NCLK):
            f.gnss_dt[c] = dt[c]
        if k == 4:
            before = T2.get(G.AMB_RTK, 0).value
        T2.preprocess(e, f)
        dt = np.array(f.gnss_dt[:])
    assert T2.get(G.AMB_RTK, 0).continue_count == 5 and T2.get(G.AMB_RTK, 0).value == before


# ---- GPU: product against oracle ------------------------------------------------------------------------------------
def _compare_epoch(cfg, To, Tp, eo, ep, fo, fp, oo, op):
    # bookkeeping: identical decisions
    assert tuple(oo.c.n_new[:]) == tuple(op.c.n_new[:])
    assert (oo.c.n_slip_rtk, oo.c.n_slip_spp, oo.c.n_factors) == (op.c.n_slip_rtk, op.c.n_slip_spp, op.c.n_factors)
    assert eo.n_obs == ep.n_obs
    for i in range(eo.n_obs):
        a, b = eo.obs[i], ep.obs[i]
        assert (a.rtk_n[0], a.spp_n[0], a.pcorr_n[0]) == (b.rtk_n[0], b.spp_n[0], b.pcorr_n[0])
        assert abs(a.el - b.el) < 1e-12
        assert (a.rtk_l[0], a.spp_l[0], a.spp_p0[0], a.spp_p[0]) == (b.rtk_l[0], b.spp_l[0], b.spp_p0[0], b.spp_p[0])
    for fam in range(3):
        assert To.count(fam) == Tp.count(fam)
        for h in range(To.count(fam)):
            a, b = To.get(fam, h), Tp.get(fam, h)
            assert (a.continue_count, a.slip_count, a.half_flag, a.sys, a.f, a.sat, a.last_update_time) == \
                   (b.continue_count, b.slip_count, b.half_flag, b.sys, b.f, b.sat, b.last_update_time)
            assert abs(a.value - b.value) < 1e-6 * max(1.0, abs(a.value)), (fam, h, a.value, b.value)
    # the epoch's prior: same keep blocks, same information (J0 itself is defined up to an orthogonal factor)
    ko, xo, Jo, ro = oo.prior()
    kp, xp, Jp, rp = op.prior()
    assert ko == kp and np.array_equal(xo, xp)
    Ao, Ap = Jo.T @ Jo, Jp.T @ Jp
    assert np.abs(Ao - Ap).max() < 1e-9 * np.abs(Ao).max()
    bo, bp = Jo.T @ ro, Jp.T @ rp
    assert np.abs(bo - bp).max() < 1e-8 * np.abs(bo).max()
    assert abs(ro @ ro - rp @ rp) < 1e-8 * max(1.0, ro @ ro)
    # clocks / blackvalue after the initialisation solve
    for c in range(G.NCLK):
        assert abs(fo.gnss_dt[c] - fp.gnss_dt[c]) < 1e-6 * max(1.0, abs(fo.gnss_dt[c]))
    assert abs(fo.blackvalue - fp.blackvalue) < 1e-9
    so, sp = oo.c.init_summary, op.c.init_summary
    assert so.termination_type == sp.termination_type and so.num_iterations == sp.num_iterations
    assert abs(so.initial_cost - sp.initial_cost) < 1e-9 * so.initial_cost
    assert abs(so.final_cost - sp.final_cost) < 1e-6 * max(1.0, so.final_cost)


def _variant(name):
    cfg = G.default_config()
    if name == "spp":       # rover-only: SPP pseudorange / carrier phase + pseudorange corrections, no base station
        cfg.use_rtk = cfg.use_rtd = 0
        cfg.use_spp_phase = cfg.use_spp_correction = 1
        cfg.estimate_pcorrection_period = 3
    return cfg


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["rtk", "spp"])
def test_gpu_preprocess_matches_the_oracle_over_a_scenario(variant):
    """Three receivers (different seeds) through one swgn_gnss_preprocess call per epoch, 14 epochs."""
    cfg = _variant(variant)
    seeds = (0, 5, 9)
    sco = [S.Scenario(s, cfg=cfg) for s in seeds]
    scp = [S.Scenario(s, cfg=cfg) for s in seeds]
    To = [ob.OracleGnssTracker(cfg) for _ in seeds]
    Tp = [G.Tracker(cfg) for _ in seeds]
    dt = [np.zeros(G.NCLK) for _ in seeds]
    black = [0.02 for _ in seeds]
    seen_slip = 0
    for k in range(14):
        eo, fo, ep, fp, keep = [], [], [], [], []
        for r in range(len(seeds)):
            e1, o1, f1 = sco[r].epoch(k)
            e2, o2, f2 = scp[r].epoch(k)
            for f in (f1, f2):
                for c in range(G.NCLK):
                    f.gnss_dt[c] = dt[r][c]
                f.blackvalue = black[r]
                if k == 12 and r == 1:
                    f.not_fix_count = cfg.phase_all_reset_count + 1
            eo.append(e1), fo.append(f1), ep.append(e2), fp.append(f2)
            keep += [o1, o2]
        oo = [To[r].preprocess(eo[r], fo[r]) for r in range(len(seeds))]
        op = G.preprocess(Tp, ep, fp)
        for r in range(len(seeds)):
            _compare_epoch(cfg, To[r], Tp[r], eo[r], ep[r], fo[r], fp[r], oo[r], op[r])
            seen_slip += oo[r].c.n_slip_rtk + oo[r].c.n_slip_spp
            # both sides continue from the ORACLE's estimates so that rounding differences do not accumulate into
            # a different decision several epochs later
            dt[r], black[r] = np.array(fo[r].gnss_dt[:]), fo[r].blackvalue
            for fam in range(3):
                for h in range(To[r].count(fam)):
                    Tp[r].set_value(fam, h, To[r].get(fam, h).value)
    assert seen_slip >= 3


@pytest.mark.gpu
def test_gpu_gate_residuals_and_records():
    cfg = G.default_config()
    sc = S.Scenario(4, cfg=cfg)
    e, obs, f = sc.epoch(0)
    rec = np.zeros((e.n_obs, 16))
    for i in range(e.n_obs):
        rec[i, 0:3] = obs[i].sat_pos[:]
        rec[i, 3:6] = np.array(f.pose[:3]) + np.array(e.base_xyz[:])
        rec[i, 9] = cfg.lams[obs[i].sys][0]
        rec[i, 10:13] = obs[i].rtk_l[0], 17.0 + i, 3.5
        rec[i, 13:16] = obs[i].spp_l[0], -4.0 - i, -1.25
    out = G.gate_residuals(rec)
    xyz = rec[0, 3:6].copy()
    ob.update_azel(xyz, e)
    for i in range(e.n_obs):
        if obs[i].svh == 0:
            assert abs(out[i, 0] - obs[i].el) < 1e-13
        rho = S.sagnac_range(xyz, np.array(obs[i].sat_pos[:]))
        lam = rec[i, 9]
        assert abs(out[i, 1] - (rho - rec[i, 11] * lam - rec[i, 10] * lam + rec[i, 12])) < 1e-6
        assert abs(out[i, 2] - (rho - rec[i, 14] * lam - rec[i, 13] * lam + rec[i, 15])) < 1e-6
    # the packer after one preprocessing: kinds and block references as AddGnssResidual orders them
    T = G.Tracker(cfg)
    G.preprocess([T], [e], [f])
    kind, blocks, data, clk, amb = T.records(e, f)
    n_rtk = sum(1 for i in range(e.n_obs) if obs[i].rtk_n[0] >= 0 and obs[i].el >= cfg.azelmin)
    assert list(kind[:n_rtk]) == [2] * n_rtk and list(kind[n_rtk:2 * n_rtk]) == [3] * n_rtk and list(kind[2 * n_rtk:]) == [4] * n_rtk
    assert list(clk) == [0, 2, 4, 12] and len(amb) == n_rtk
    assert (blocks[:n_rtk, 0] == 0).all() and (blocks[2 * n_rtk:, 0] == 1).all() and (blocks[2 * n_rtk:, 2] == 0).all()


@pytest.mark.gpu
def test_gpu_epoch_without_usable_observations():
    cfg = G.default_config()
    sc = S.Scenario(6, cfg=cfg)
    e, obs, f = sc.epoch(0)
    for i in range(e.n_obs):
        obs[i].svh = 1
    T = G.Tracker(cfg)
    out = G.preprocess([T], [e], [f])[0]
    keep, x0, J0, r0 = out.prior()
    assert keep == [(G.KEEP_BLACK, -1, 0)] and J0.shape == (1, 1) and J0[0, 0] == 1.0 and r0[0] == f.blackvalue
    assert T.count(G.AMB_RTK) == 0


_REFDEMO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libswgn_refdemo.so")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(_REFDEMO), reason="oracle/_ref not built")
def test_gpu_preprocess_through_the_reference_wire_struct():
    """The epoch goes through a real mea_t (RVI/gnss/include/common_function.h:115-123, compiled from the reference) and the
    maintainer-side binding shim/reference_gnss_binding.h: same results as the direct call."""
    L = C.CDLL(_REFDEMO)
    L.swgn_refdemo_gnss_preprocess.argtypes = [C.c_void_p, C.POINTER(G.Epoch), C.POINTER(G.Frame), C.POINTER(G.Output), C.POINTER(C.c_int)]
    cfg = G.default_config()
    sa, sb = S.Scenario(7, cfg=cfg), S.Scenario(7, cfg=cfg)
    Ta, Tb = G.Tracker(cfg), G.Tracker(cfg)
    nbytes = C.c_int()
    for k in range(3):
        ea, oa, fa = sa.epoch(k)
        eb, ob_, fb = sb.epoch(k)
        out_a = G.preprocess([Ta], [ea], [fa])[0]
        out_b = G.OutputBuffers()
        assert L.swgn_refdemo_gnss_preprocess(Tb.h, C.byref(eb), C.byref(fb), C.byref(out_b.c), C.byref(nbytes)) == 0
        ka, xa, Ja, ra = out_a.prior()
        kb, xb, Jb, rb = out_b.prior()
        assert ka == kb and np.array_equal(xa, xb) and np.array_equal(Ja, Jb) and np.array_equal(ra, rb)
        assert [oa[i].el for i in range(ea.n_obs)] == [ob_[i].el for i in range(eb.n_obs)]
        assert [oa[i].rtk_n[0] for i in range(ea.n_obs)] == [ob_[i].rtk_n[0] for i in range(eb.n_obs)]
        assert fa.gnss_dt[:] == fb.gnss_dt[:]
    assert nbytes.value > 64 * 300   # sizeof(mea_t): MAXOBS ObsMea records plus the header


def test_oracle_prior_matches_the_reference_marginalization_info():
    """The reference's own MarginalizationInfo (RVI/factor/marginalization_factor.cpp:58-400, compiled into oracle/_ref) over
    the reference's own factor objects (RTKCarrierPhaseFactor, RTKPseudorangeFactor, SppDopplerFactor, InitialBlackFactor with
    the constructor arguments AddGnssResidual passes): addResidualBlockInfo x n, marginalize(true, true),
    getParameterBlocks().  Its linearised factor carries the same information as the oracle's restatement; the keep blocks
    come in the reference's unordered_map order and are permuted."""
    L = ob.ref()
    if L is None or not hasattr(L, "ref_marginalize"):
        pytest.skip("oracle/_ref not built")
    i32, f64, P = C.c_int32, C.c_double, C.POINTER
    L.ref_marginalize.argtypes = [C.c_int, P(i32), P(i32), P(f64), C.c_int, P(i32), P(i32), P(i32), P(f64), P(i32), P(i32), P(i32), P(i32),
                                  P(i32), P(f64), P(f64)]
    O = ob.oracle()
    O.oracle_gnss_epoch_factors.argtypes = [C.c_void_p, P(G.Epoch), P(G.Frame), C.c_int, P(i32), P(i32), P(f64), P(f64), P(i32)]
    for variant in ("rtk", "spp"):
        cfg = _variant(variant)
        sc = S.Scenario(11, cfg=cfg)
        T, log = run_oracle(sc, cfg, 5)
        for rec in log[1:]:
            e, f_in, out = rec["epoch"], rec["frame_in"], rec["out"]
            cap = 6 * e.n_obs + 1
            kind, off, recs = np.zeros(cap, np.int32), np.zeros(3 * cap, np.int32), np.zeros(16 * cap)
            store, ns = np.zeros(30 + 6 * e.n_obs), i32()
            nf = O.oracle_gnss_epoch_factors(T.h, C.byref(e), C.byref(f_in), cap, kind.ctypes.data_as(P(i32)), off.ctypes.data_as(P(i32)),
                                             recs.ctypes.data_as(P(f64)), store.ctypes.data_as(P(f64)), C.byref(ns))
            assert nf == out.c.n_factors
            offs = sorted({int(o) for o in off[:3 * nf] if o >= 0})
            bidx = {o: k for k, o in enumerate(offs)}
            size = np.array([7 if o == 0 else 9 if o == 7 else 1 for o in offs], np.int32)
            drop = np.array([1 if 17 <= o < 30 else 0 for o in offs], np.int32)
            boff = np.array(offs, np.int32)
            blocks = np.array([bidx[int(o)] if o >= 0 else -1 for o in off[:3 * nf]], np.int32)
            n, m, nk = i32(), i32(), i32()
            kb, ki = np.zeros(len(offs), np.int32), np.zeros(len(offs), np.int32)
            J, r = np.zeros(out.c.n ** 2), np.zeros(out.c.n)
            rc = L.ref_marginalize(nf, kind.ctypes.data_as(P(i32)), blocks.ctypes.data_as(P(i32)), recs.ctypes.data_as(P(f64)), len(offs),
                                   size.ctypes.data_as(P(i32)), drop.ctypes.data_as(P(i32)), boff.ctypes.data_as(P(i32)),
                                   store.ctypes.data_as(P(f64)), C.byref(n), C.byref(m), C.byref(nk), kb.ctypes.data_as(P(i32)),
                                   ki.ctypes.data_as(P(i32)), J.ctypes.data_as(P(f64)), r.ctypes.data_as(P(f64)))
            assert rc == 0
            keep, x0, J0, r0 = out.prior()
            assert n.value == out.c.n and nk.value == out.c.n_keep and m.value == int(drop.sum())
            # column permutation: reference keep order -> oracle keep order (both identify blocks by their storage offset)
            my_off = []
            for kd, h, idx in keep:
                if kd == G.KEEP_POSE:
                    my_off.append(0)
                elif kd == G.KEEP_SPEED_BIAS:
                    my_off.append(7)
                elif kd == G.KEEP_BLACK:
                    my_off.append(16)
            amb_offs = [o for o in offs if o >= 30]
            my_off += amb_offs   # the oracle's ambiguity order is storage order
            assert len(my_off) == len(keep)
            tang = lambda o: 6 if o == 0 else 9 if o == 7 else 1
            ref_cols = {}
            for k in range(nk.value):
                o = offs[kb[k]]
                ref_cols[o] = list(range(ki[k], ki[k] + tang(o)))
            perm = [c for o in my_off for c in ref_cols[o]]
            Jr = J.reshape(n.value, n.value)[:, perm]
            A_ref, A_or = Jr.T @ Jr, J0.T @ J0
            assert np.abs(A_ref - A_or).max() < 1e-9 * np.abs(A_or).max()
            b_ref, b_or = Jr.T @ r, J0.T @ r0
            assert np.abs(b_ref - b_or).max() < 1e-8 * np.abs(b_or).max()
            assert abs(r @ r - r0 @ r0) < 1e-8 * max(1.0, r0 @ r0)


# ---- the prior rebuild after FIX_CONTINUE_THRESHOLD accepted fixes (RVI/swf/swf_lambda.cpp:249-355) --------------------
def _fixed_integer_job(seed, zero_ambiguities=True):
    """A job on a real epoch prior: keep = pose, speed-bias, blackvalue, 18 ambiguities; double differences inside every
    system against its first ambiguity; x = a state near the prior's linearisation point."""
    cfg = G.default_config()
    sc = S.Scenario(seed, cfg=cfg)
    T, log = run_oracle(sc, cfg, 2)
    out, e, obs = log[1]["out"], log[1]["epoch"], log[1]["obs"]
    keep, x0, J0, r0 = out.prior()
    size = [7 if k[0] == G.KEEP_POSE else 9 if k[0] == G.KEEP_SPEED_BIAS else 1 for k in keep]
    idx = [k[2] for k in keep]
    rng = np.random.default_rng(seed)
    x = x0 + 0.01 * rng.normal(size=len(x0))
    x[3:7] /= np.linalg.norm(x[3:7])
    amb_keep = [i for i, k in enumerate(keep) if k[0] == G.KEEP_AMB_RTK]
    sys_of = {k[1]: T.get(G.AMB_RTK, k[1]).sys for k in keep if k[0] == G.KEEP_AMB_RTK}
    if zero_ambiguities:
        x[16:] = 0.0   # PhaseBiasSaveAndReset
    else:
        x[16:] = rng.normal(size=len(x) - 16) * 3
    dd, F, sf = [], [], []
    for s in range(3):
        members = [i for i in amb_keep if sys_of[keep[i][1]] == s]
        for a in members[1:]:
            dd.append((a, members[0]))
            F.append(float(rng.integers(-40, 40)))
            sf.append(2 * s)
    return G.FixedIntegerArrays(size, idx, x0, J0, r0, x, dd, F, sf)


def _oracle_fixed_integer(job):
    O = ob.oracle()
    O.oracle_fixed_integer_prior.argtypes = [C.POINTER(G.FixedIntegerJob)]
    assert O.oracle_fixed_integer_prior(C.byref(job.c)) == 0
    return job.J0_out.copy(), job.r0_out.copy()


def test_fixed_integer_prior_oracle_matches_the_reference_classes():
    """Oracle restatement of swf_lambda.cpp:249-355 against the reference's own MarginalizationInfo / MarginalizationFactor /
    FixedIntegerFactor executed here (oracle/ref_marg_shim.cpp), and against the closed form: the new information is the old
    one plus the double-difference constraints w^2 (e_p - e_n)(e_p - e_n)' with the dummies eliminated."""
    L = ob.ref()
    if L is None or not hasattr(L, "ref_fixed_integer_prior"):
        pytest.skip("oracle/_ref not built")
    i32, f64, P = C.c_int32, C.c_double, C.POINTER
    L.ref_fixed_integer_prior.argtypes = [C.c_int, C.c_int, P(i32), P(i32), P(f64), P(f64), P(f64), P(f64), C.c_int, P(i32), P(f64), P(i32),
                                          f64, P(i32), P(i32), P(f64), P(f64)]
    for seed, zero in ((21, True), (22, False)):
        job = _fixed_integer_job(seed, zero)
        Jo, ro = _oracle_fixed_integer(job)
        n, nk = job.c.n, job.c.n_keep
        order, col = np.zeros(nk, np.int32), np.zeros(nk, np.int32)
        Jr, rr = np.zeros((n, n)), np.zeros(n)
        rc = L.ref_fixed_integer_prior(nk, n, job.c.keep_size, job.c.keep_idx, job.c.x0, job.c.J0, job.c.r0, job.c.x, job.c.n_dd, job.c.dd_keep,
                                       job.c.F, job.c.dd_sysfreq, job.c.istd, order.ctypes.data_as(P(i32)), col.ctypes.data_as(P(i32)),
                                       Jr.ctypes.data_as(P(f64)), rr.ctypes.data_as(P(f64)))
        assert rc == 0
        tang = [6 if s == 7 else int(s) for s in job.keep_size]
        ref_cols = {int(order[k]): list(range(col[k], col[k] + tang[order[k]])) for k in range(nk)}
        perm = [c for q in range(nk) for c in ref_cols[q]]
        Jr = Jr[:, perm]
        Ao, Ar = Jo.T @ Jo, Jr.T @ Jr
        assert np.abs(Ao - Ar).max() < 1e-9 * np.abs(Ao).max()
        assert np.abs(Jo.T @ ro - Jr.T @ rr).max() < 1e-8 * max(1.0, np.abs(Jo.T @ ro).max())
        # closed form of the information: per system the dummy couples its members like a star; eliminating it leaves
        # w^2 (I - 11'/k) on the k members (the reference ambiguity included) in the shifted variables
        A_old = job.J0.reshape(n, n).T @ job.J0.reshape(n, n)
        w2 = job.c.istd ** 2
        add = np.zeros((n, n))
        for s in sorted(set(job.dd_sysfreq)):
            pairs = [tuple(job.dd_keep[2 * d:2 * d + 2]) for d in range(job.c.n_dd) if job.dd_sysfreq[d] == s]
            members = [job.keep_idx[pairs[0][1]]] + [job.keep_idx[p] for p, _ in pairs]
            k = len(members)
            add[np.ix_(members, members)] += w2 * (np.eye(k) - np.ones((k, k)) / k)
        assert np.abs(Ao - (A_old + add)).max() < 1e-8 * np.abs(Ao).max()


@pytest.mark.gpu
def test_gpu_fixed_integer_prior_matches_the_oracle():
    jobs = [_fixed_integer_job(31, True), _fixed_integer_job(32, False), _fixed_integer_job(33, True)]
    want = [_oracle_fixed_integer(j) for j in jobs]
    for j in jobs:
        j.J0_out[:] = 0
        j.r0_out[:] = 0
    G.fixed_integer_prior(jobs)
    for j, (Jo, ro) in zip(jobs, want):
        Ag, Ao = j.J0_out.T @ j.J0_out, Jo.T @ Jo
        assert np.abs(Ag - Ao).max() < 1e-9 * np.abs(Ao).max()
        bg, bo = j.J0_out.T @ j.r0_out, Jo.T @ ro
        assert np.abs(bg - bo).max() < 1e-8 * max(1.0, np.abs(bo).max())
        assert abs(j.r0_out @ j.r0_out - ro @ ro) < 1e-8 * max(1.0, ro @ ro)


def test_chain_frame_from_the_epoch_prior_matches_the_reference_add_marg_info():
    """swgn_gnss_chain_frame (host-side scatter, no device) against IMUGNSSBase::AddMargInfo of the reference
    (gnss_imu_factor.cpp:245-352) executed on real MarginalizationInfo objects: three consecutive epoch priors of one
    receiver -- ambiguities shared between the epochs, one replaced after a slip, blackvalue treated as a phase bias like
    the reference does -- become three hidden frames of one chain."""
    L = ob.ref()
    if L is None or not hasattr(L, "ref_add_marg_info"):
        pytest.skip("oracle/_ref not built")
    i32, f64, P = C.c_int32, C.c_double, C.POINTER
    L.ref_add_marg_info.argtypes = [C.c_int, P(i32), P(i32), P(i32), P(i32), P(f64), P(i32), P(f64), P(f64), C.c_int, P(i32), P(i32),
                                    P(f64), P(f64), P(f64), P(f64), P(f64)]
    lib = G._proto()
    lib.swgn_gnss_chain_frame.argtypes = [P(G.Output), P(f64), P(f64), i32, P(i32), P(f64), P(f64), P(f64)]
    cfg = G.default_config()
    sc = S.Scenario(0, cfg=cfg)
    T, log = run_oracle(sc, cfg, 6)
    epochs = log[3:6]     # epoch 4 carries an announced slip: a new ambiguity joins, the old one stays in the chain
    # identities: 0..2 pose of epoch e, 3..5 speed-bias of epoch e, 6 blackvalue, 7.. ambiguities by handle
    ids, sizes, idxs, x0s, ns, As, bs, nkeep = [], [], [], [], [], [], [], []
    for e, rec in enumerate(epochs):
        keep, x0, J0, r0 = rec["out"].prior()
        for kd, h, col in keep:
            ids.append(e if kd == G.KEEP_POSE else 3 + e if kd == G.KEEP_SPEED_BIAS else 6 if kd == G.KEEP_BLACK else 7 + h)
            sizes.append(7 if kd == G.KEEP_POSE else 9 if kd == G.KEEP_SPEED_BIAS else 1)
            idxs.append(col)
        x0 = x0.copy()
        x0[16] = 0.0   # AddMargInfo asserts that every size-1 block was linearised at 0 (the estimator's blackvalue is 0 there)
        x0s.append(x0)
        ns.append(J0.shape[0])
        As.append((J0.T @ J0).ravel())
        bs.append(J0.T @ r0)
        nkeep.append(len(keep))
    max_ids = max(ids) + 1
    arr = lambda v, t: np.ascontiguousarray(np.concatenate([np.atleast_1d(x) for x in v]), t)
    ids_a, sizes_a, idxs_a = arr(ids, np.int32), arr(sizes, np.int32), arr(idxs, np.int32)
    k_ref = i32()
    slot = np.zeros(max_ids, np.int32)
    kcap = max_ids
    pH, pr, pN, NN, Nr = np.zeros(3 * 225), np.zeros(45), np.zeros(3 * 15 * kcap), np.zeros(kcap * kcap), np.zeros(kcap)
    rc = L.ref_add_marg_info(3, arr(nkeep, np.int32).ctypes.data_as(P(i32)), sizes_a.ctypes.data_as(P(i32)), idxs_a.ctypes.data_as(P(i32)),
                             ids_a.ctypes.data_as(P(i32)), ob._dp(arr(x0s, np.float64)), arr(ns, np.int32).ctypes.data_as(P(i32)),
                             ob._dp(arr(As, np.float64)), ob._dp(arr(bs, np.float64)), max_ids, C.byref(k_ref),
                             slot.ctypes.data_as(P(i32)), ob._dp(pH), ob._dp(pr), ob._dp(pN), ob._dp(NN), ob._dp(Nr))
    assert rc == 0
    k = k_ref.value
    # the caller's slot rule = the reference's: size-1 keep blocks in order of first appearance
    order = []
    for i, s in zip(ids, sizes):
        if s == 1 and i not in order:
            order.append(i)
    assert k == len(order) and all(slot[i] == q for q, i in enumerate(order))
    assert k > len([i for i in ids[:nkeep[0]] if i >= 6])          # the slip added a phase bias to the chain
    chain_N = np.zeros(k * k + k)
    o = 0
    for e, rec in enumerate(epochs):
        out = rec["out"]
        out.x0[16] = 0.0
        keep_slot = np.array([order.index(i) if s == 1 else -1 for i, s in zip(ids[o:o + nkeep[e]], sizes[o:o + nkeep[e]])], np.int32)
        o += nkeep[e]
        frame, frame_N = np.zeros(274), np.zeros(15 * k)
        pose = np.array(rec["frame_in"].pose[:])
        sb = np.array(rec["frame_in"].speed_bias[:])
        st = lib.swgn_gnss_chain_frame(C.byref(out.c), ob._dp(pose), ob._dp(sb), k, keep_slot.ctypes.data_as(P(i32)), ob._dp(frame),
                                       ob._dp(frame_N), ob._dp(chain_N))
        assert st == 0
        H, rhs = frame[48:273], frame[32:47]
        scale = np.abs(pH[225 * e:225 * e + 225]).max()
        assert np.abs(H - pH[225 * e:225 * e + 225]).max() < 1e-12 * scale
        assert np.abs(rhs - pr[15 * e:15 * e + 15]).max() < 1e-12 * max(1.0, np.abs(pr).max())
        assert np.abs(frame_N - pN[e * 15 * k:(e + 1) * 15 * k]).max() < 1e-12 * scale
        assert np.array_equal(frame[0:7], pose) and np.array_equal(frame[16:23], pose) and np.array_equal(frame[23:32], sb)
    assert np.abs(chain_N[:k * k] - NN[:k * k]).max() < 1e-12 * np.abs(NN).max()
    assert np.abs(chain_N[k * k:] - Nr[:k]).max() < 1e-12 * max(1.0, np.abs(Nr).max())


_REF_EST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_estimator.so")


class RefEstimator:
    """The reference's own SWFOptimization::GnssPreprocess (swf_gnss.cpp + swf_core.cpp compiled unmodified into
    oracle/_ref/libref_estimator.so), running on this repository's ceres:: shim and the device."""

    def __init__(self, cfg):
        L = C.CDLL(_REF_EST)
        i32, f64, P = C.c_int32, C.c_double, C.POINTER
        L.ref_est_create.restype = C.c_void_p
        L.ref_est_create.argtypes = [P(G.Config)]
        L.ref_est_destroy.argtypes = [C.c_void_p]
        L.ref_est_gnss_preprocess.argtypes = [C.c_void_p, P(G.Config), P(G.Epoch), P(G.Frame), C.c_int, C.c_int, P(i32), P(i32), P(i32), P(i32),
                                              P(i32), P(i32), P(f64), P(f64), P(f64)]
        L.ref_est_ambiguities.argtypes = [C.c_void_p, C.c_int, C.c_int, P(i32), P(i32), P(f64), P(i32), P(i32), P(f64)]
        L.ref_est_set_ambiguity.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, f64]
        self.L, self.cfg = L, cfg
        self.h = L.ref_est_create(C.byref(cfg))

    def preprocess(self, epoch, frame):
        i32, f64, P = C.c_int32, C.c_double, C.POINTER
        cap_k, cap_n = 3 + 6 * G.MAXOBS, 16 + 6 * G.MAXOBS
        n, nk = i32(), i32()
        kind, s2f, pos, idx = (np.zeros(cap_k, np.int32) for _ in range(4))
        x0, J0, r0 = np.zeros(cap_n + 3), np.zeros(cap_n * cap_n), np.zeros(cap_n)
        rc = self.L.ref_est_gnss_preprocess(self.h, C.byref(self.cfg), C.byref(epoch), C.byref(frame), cap_k, cap_n, C.byref(n), C.byref(nk),
                                            kind.ctypes.data_as(P(i32)), s2f.ctypes.data_as(P(i32)), pos.ctypes.data_as(P(i32)),
                                            idx.ctypes.data_as(P(i32)), ob._dp(x0), ob._dp(J0), ob._dp(r0))
        assert rc == 0, rc
        n, nk = n.value, nk.value
        keep = [(int(kind[k]), int(s2f[k]), int(pos[k]), int(idx[k])) for k in range(nk)]
        return keep, x0, J0[:n * n].reshape(n, n).copy(), r0[:n].copy()

    def ambiguities(self, fam):
        i32, P = C.c_int32, C.POINTER
        cap = 4096
        s2f, pos, cc, sl = (np.zeros(cap, np.int32) for _ in range(4))
        val, lut = np.zeros(cap), np.zeros(cap)
        n = self.L.ref_est_ambiguities(self.h, fam, cap, s2f.ctypes.data_as(P(i32)), pos.ctypes.data_as(P(i32)), ob._dp(val),
                                       cc.ctypes.data_as(P(i32)), sl.ctypes.data_as(P(i32)), ob._dp(lut))
        return {(int(s2f[k]), int(pos[k])): (float(val[k]), int(cc[k]), int(sl[k]), float(lut[k])) for k in range(n)}

    def set_ambiguity(self, fam, s2f, pos, v):
        self.L.ref_est_set_ambiguity(self.h, fam, s2f, pos, v)

    def close(self):
        self.L.ref_est_destroy(self.h)


def _product_ambiguities(T, fam):
    """{(sat * 2 + f, position in that list): (handle, value, continue_count, slip_count, last_update_time)}"""
    out, seen = {}, {}
    for h in range(T.count(fam)):
        a = T.get(fam, h)
        key = a.sat * 2 + a.f
        out[(key, seen.get(key, 0))] = (h, a.value, a.continue_count, a.slip_count, a.last_update_time)
        seen[key] = seen.get(key, 0) + 1
    return out


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(_REF_EST), reason="oracle/_ref not built")
@pytest.mark.parametrize("variant", ["rtk", "spp"])
def test_gpu_preprocess_matches_the_reference_estimator_code(variant):
    """The reference's GnssPreprocess ITSELF (unmodified swf_gnss.cpp / swf_core.cpp / marginalization_factor.cpp / gnss_factor.cpp,
    its ceres::Solve calls going through the shim to the device -- with no linear_solver_ordering, as the reference leaves it)
    against swgn_gnss_preprocess, epoch by epoch over the scripted scenario: same ambiguity bookkeeping (new / reset / counted /
    timed-out entries per satellite list), same elevations and masks, the same information in the epoch's prior, the same
    clocks and ambiguity values after the initialisation solve."""
    cfg = _variant(variant)
    cfg.estimate_pcorrection_period = 500   # a compile-time constant of the reference (parameters.h:27)
    scr, scp = S.Scenario(13, cfg=cfg, unhealthy_has_phase=False), S.Scenario(13, cfg=cfg, unhealthy_has_phase=False)
    R = RefEstimator(cfg)
    T = G.Tracker(cfg)
    dt, black = np.zeros(G.NCLK), 0.0    # AddMargInfo-style consumers need blackvalue linearised at 0; the estimator starts it at 0
    slips = 0
    for k in range(14):
        er, obr, fr = scr.epoch(k)
        ep, obp, fp = scp.epoch(k)
        for f in (fr, fp):
            for c in range(G.NCLK):
                f.gnss_dt[c] = dt[c]
            f.blackvalue = black
            if k == 12:
                f.not_fix_count = cfg.phase_all_reset_count + 1
        keep_r, x0_r, J_r, r_r = R.preprocess(er, fr)
        out = G.preprocess([T], [ep], [fp])[0]
        keep_p, x0_p, J_p, r_p = out.prior()
        slips += out.c.n_slip_rtk + out.c.n_slip_spp
        # observations
        for i in range(er.n_obs):
            assert abs(obr[i].el - obp[i].el) < 1e-12
            assert (obr[i].rtk_l[0], obr[i].spp_l[0], obr[i].spp_p0[0], obr[i].spp_p[0]) == (obp[i].rtk_l[0], obp[i].spp_l[0], obp[i].spp_p0[0], obp[i].spp_p[0])
        # bookkeeping, list by list
        for fam in range(3):
            ar, ap = R.ambiguities(fam), _product_ambiguities(T, fam)
            assert set(ar) == set(ap), (k, fam)
            for key, (val, cc, sl, lut) in ar.items():
                h, pval, pcc, psl, plut = ap[key]
                if fam == G.AMB_PCORR:   # the reference never sets SLIP_COUNT of a pseudorange-correction entry (swf_gnss.cpp:477-487: stack garbage)
                    sl = psl
                assert (cc, sl, lut) == (pcc, psl, plut), (k, fam, key)
                assert abs(val - pval) < 1e-6 * max(1.0, abs(val)), (k, fam, key, val, pval)
        # the epoch's prior: same keep blocks (the reference's order is its unordered_map's), same information
        handle_of = {}
        for fam in range(3):
            for key, v in _product_ambiguities(T, fam).items():
                handle_of[(fam, key)] = v[0]
        cols_p = {}
        for kd, h, col in keep_p:
            cols_p[(kd, h)] = col
        perm, tang = [], {G.KEEP_POSE: 6, G.KEEP_SPEED_BIAS: 9}
        assert len(keep_r) == len(keep_p)
        ref_cols = {}
        for kd, s2f, pos, col in keep_r:
            h = -1 if kd < G.KEEP_AMB_RTK else handle_of[(kd - G.KEEP_AMB_RTK, (s2f, pos))]
            ref_cols[(kd, h)] = col
        for kd, h, col in keep_p:
            t = tang.get(kd, 1)
            perm += list(range(ref_cols[(kd, h)], ref_cols[(kd, h)] + t))
        Jr = J_r[:, perm]
        A_r, A_p = Jr.T @ Jr, J_p.T @ J_p
        assert np.abs(A_r - A_p).max() < 1e-9 * np.abs(A_r).max()
        b_r, b_p = Jr.T @ r_r, J_p.T @ r_p
        assert np.abs(b_r - b_p).max() < 1e-8 * max(1.0, np.abs(b_r).max())
        # state after the initialisation solve
        for c in range(G.NCLK):
            assert abs(fr.gnss_dt[c] - fp.gnss_dt[c]) < 1e-6 * max(1.0, abs(fr.gnss_dt[c])), (k, c)
        assert abs(fr.blackvalue - fp.blackvalue) < 1e-9
        # continue both from the reference's estimates
        dt, black = np.array(fr.gnss_dt[:]), fr.blackvalue
        for fam in range(3):
            ap = _product_ambiguities(T, fam)
            for key, (val, cc, sl, lut) in R.ambiguities(fam).items():
                T.set_value(fam, ap[key][0], val)
    assert slips >= 1 and T.count(G.AMB_RTK) > 30
    R.close()


@pytest.mark.skipif(not os.path.exists(_REF_EST), reason="oracle/_ref not built")
@pytest.mark.parametrize("strong", [True, False])
def test_oracle_lambda_search_matches_the_reference_estimator_code(strong):
    """SWFOptimization::LambdaSearch ITSELF (unmodified swf_lambda.cpp with FindReferenceSatellites, lambda.cpp and the prior
    rebuild of :249-355 through the reference's MarginalizationInfo) executed on an estimator holding three preprocessed
    epochs, against the oracle's restatement of the decision (reference satellites, double-difference gate, lambda(),
    ratio tests) and of the rebuild.  strong: an informative ambiguity matrix -> the fix is accepted and the prior rebuilt;
    weak: the ratio test fails, not_fix_count advances, the prior stays."""
    cfg = G.default_config()
    sc = S.Scenario(17, cfg=cfg, unhealthy_has_phase=False, half_flag=11)
    R = RefEstimator(cfg)
    i32, f64, P = C.c_int32, C.c_double, C.POINTER
    R.L.ref_est_lambda_search.argtypes = [C.c_void_p, C.c_int, C.c_int, P(i32), P(i32), P(f64), P(f64), P(f64), P(i32), P(i32), P(i32), P(f64), P(f64)]
    epochs = []
    dt = np.zeros(G.NCLK)
    for k in range(3):
        e, obs, f = sc.epoch(k)
        f.blackvalue = 0.0
        f.nonlinear = 0         # no residual gating: the ambiguities keep their first entries whether or not a device is present
        for c in range(G.NCLK):
            f.gnss_dt[c] = dt[c]
        R.preprocess(e, f)      # (its initialisation solve needs the device; without one it reports FAILURE and changes nothing)
        dt = np.array(f.gnss_dt[:])
        epochs.append((e, obs))
    amb = sorted(R.ambiguities(G.AMB_RTK))           # [(sat * 2 + f, position in the list)]
    n = len(amb)
    assert n == 18
    row = {key: r for r, key in enumerate(amb)}
    rng = np.random.default_rng(3)
    sys_of = {}
    for e, obs in epochs:
        for i in range(e.n_obs):
            sys_of[obs[i].sat * 2] = obs[i].sys
    ints = rng.integers(-60, 60, n).astype(float)
    y = ints + np.array([0.37, -0.81, 0.12])[[sys_of[k[0]] for k in amb]] + rng.normal(0, 0.01 if strong else 0.3, n)
    for (s2f, pos), v in zip(amb, y):
        R.set_ambiguity(G.AMB_RTK, s2f, pos, float(v))
    B = rng.normal(size=(n, n))
    A = (2500.0 if strong else 3.0) * np.eye(n) + (40.0 if strong else 0.05) * (B @ B.T)
    J0 = np.linalg.cholesky(A).T.copy()
    r0 = rng.normal(size=n) * 0.1
    flags = np.zeros(6, np.int32)
    keep_row, keep_col = np.zeros(n, np.int32), np.zeros(n, np.int32)
    Jn, rn = np.zeros((n, n)), np.zeros(n)
    s2f_a = np.array([k[0] for k in amb], np.int32)
    pos_a = np.array([k[1] for k in amb], np.int32)
    rc = R.L.ref_est_lambda_search(R.h, 3, n, s2f_a.ctypes.data_as(P(i32)), pos_a.ctypes.data_as(P(i32)), ob._dp(np.ascontiguousarray(A)),
                                   ob._dp(J0), ob._dp(r0), flags.ctypes.data_as(P(i32)), keep_row.ctypes.data_as(P(i32)),
                                   keep_col.ctypes.data_as(P(i32)), ob._dp(Jn), ob._dp(rn))
    assert rc == 0
    rtk_fix, fix, last_fix, not_fix_count, n_fix_solutions, rebuilt = [int(v) for v in flags]
    # ---- the oracle's decision on the same inputs
    eb, oa, sf = [0], [], []
    for e, obs in epochs:
        for i in range(e.n_obs):
            key = (obs[i].sat * 2, 0)
            if obs[i].svh == 0 and obs[i].rtk_l[0] != 0 and key in row:
                oa.append(row[key])
                sf.append(obs[i].sys * 2)
        eb.append(len(oa))
    pairs, F, res = ob.ambiguity_fix(A, y, eb, oa, sf, last_fix=0)
    assert res.status == 0 and res.n_dd >= 4
    assert bool(res.search_ok) == bool(rtk_fix) == bool(rebuilt) == strong
    assert not_fix_count == (0 if strong else 1) and n_fix_solutions == (1 if strong else 0)
    if not strong:
        R.close()
        return
    # ---- the rebuilt prior: last_marg_info + FixedIntegerFactors of the newest epoch's double differences, dummies dropped
    newest = set(oa[eb[2]:eb[3]])
    last_count = 0
    while last_count < len(pairs) and int(pairs[last_count][0]) in newest:
        last_count += 1
    dd = [(int(a), int(b)) for a, b in pairs[:last_count]]
    Fr = np.round(F[:last_count, 0])
    assert np.array_equal(Fr, np.round(ints[[a for a, _ in dd]] - ints[[b for _, b in dd]]))   # the integers that were planted
    sfd = [sys_of[amb[a][0]] * 2 for a, _ in dd]
    job = G.FixedIntegerArrays([1] * n, list(range(n)), np.zeros(n), J0, r0, np.zeros(n), dd, Fr, sfd)
    Jo, ro = _oracle_fixed_integer(job)
    perm = [0] * n
    for k in range(n):
        perm[keep_row[k]] = keep_col[k]
    Jr = Jn[:, perm]
    A_ref, A_or = Jr.T @ Jr, Jo.T @ Jo
    assert np.abs(A_ref - A_or).max() < 1e-9 * np.abs(A_ref).max()
    assert np.abs(Jr.T @ rn - Jo.T @ ro).max() < 1e-8 * max(1.0, np.abs(Jr.T @ rn).max())
    R.close()
