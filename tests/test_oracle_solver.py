"""End-to-end behaviour of the oracle on the synthetic windows of SURVEY.md 8d: convergence,
export mode (M2), is_use masks (M3), ordering errors, read-backs (A9) and the fix decision (A10)."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import swgn


def test_cfg1_one_iteration_and_convergence():
    """configs[0]: 5 KF x 50 LM VI window, one Gauss-Newton iteration on the CPU (plumbing)."""
    w = swgn.SynthWindow(1, 0)
    opt = w.options()
    opt.max_num_iterations = 1
    o = ob.OracleSolver(w.graph_p, opt)
    st, sm = o.minimize()
    assert st == 0 and sm.num_iterations == 1
    assert sm.final_cost < sm.initial_cost
    opt.max_num_iterations = 30
    o = ob.OracleSolver(w.graph_p, opt)
    st, sm = o.minimize()
    assert st == 0 and sm.termination_type == swgn_const("CONVERGENCE")
    c, r, s = o.iteration_records()
    assert np.all(np.diff(c[s == 1]) <= 1e-9)  # monotone over accepted steps


def swgn_const(name):
    return {"CONVERGENCE": 0, "NO_CONVERGENCE": 1, "FAILURE": 2}[name]


def test_cfg2_dimensions_follow_survey_formula():
    w = swgn.SynthWindow(2, 0)
    o = ob.OracleSolver(w.graph_p, w.options())
    # n_F = 14*9 (odd speed-bias) + 9 (sb0) + 30*6 (poses) + 20 (N) + 10 (drift) + 1 (blackvalue)
    assert o.n_f == 346
    # n_E = 3*300 + 15*9 + 30 clock slots + blackvalue2
    assert o.n_e == 900 + 135 + 30 + 1
    assert o.n_e_blocks == 300 + 15 + 30 + 1
    st, sm = o.minimize()
    assert st == 0 and sm.num_iterations == 8 and sm.final_cost < 1e-5 * sm.initial_cost


def test_solution_approaches_truth_with_more_iterations():
    w = swgn.SynthWindow(2, 3)
    opt = w.options()
    opt.max_num_iterations = 40
    o = ob.OracleSolver(w.graph_p, opt)
    st, sm = o.minimize()
    assert st == 0
    x, t, x0 = o.state(), w.truth(), w.state0()
    offs = w.block_offsets()
    F = w.n_frames
    pos_err = max(np.abs(x[offs[f]:offs[f] + 3] - t[offs[f]:offs[f] + 3]).max() for f in range(F))
    pos_err0 = max(np.abs(x0[offs[f]:offs[f] + 3] - t[offs[f]:offs[f] + 3]).max() for f in range(F))
    # absolute accuracy is bounded by the prior on frame 0 (sigma 5 cm) and the float ambiguities
    assert pos_err < 0.1 and pos_err < 0.6 * pos_err0, (pos_err, pos_err0)


def test_export_mode_leaves_state_untouched_and_exports_reduced_system():
    w = swgn.SynthWindow(2, 1)
    opt = w.options()
    opt.is_optimize = 0
    opt.max_num_iterations = 1
    o = ob.OracleSolver(w.graph_p, opt)
    x0 = o.state()
    st, sm = o.minimize()
    assert sm.num_iterations == 1 and sm.num_successful_steps == 1 and sm.num_unsuccessful_steps == 1
    assert np.array_equal(o.state(), x0)
    n, S, r, L = o.exports()
    assert n == o.n_f
    # the exported system is J'J-Schur with the dogleg diagonal sqrt(1e-12)*diag (M4)
    cost, res, g, J = o.evaluate()
    d = np.sqrt(np.clip((J * J).sum(0), 1e-6, 1e32)) * np.sqrt(1e-12)
    H = J.T @ J + np.diag(d * d)
    ne = o.n_e
    Sref = H[ne:, ne:] - H[ne:, :ne] @ np.linalg.solve(H[:ne, :ne], H[:ne, ne:])
    Sfull = np.triu(S) + np.triu(S, 1).T
    assert np.linalg.norm(Sfull - Sref) / np.linalg.norm(Sref) < 1e-9
    rref = (J.T @ res)[ne:] - H[ne:, :ne] @ np.linalg.solve(H[:ne, :ne], (J.T @ res)[:ne])
    assert np.linalg.norm(r - rref) / np.linalg.norm(rref) < 1e-8
    # UpdateSchur (swf_gnss.cpp:25-61) on the exported system == dense marginal of the head block
    nt = w.n_amb
    A, b = ob.update_schur(S, r, nt)
    m = n - nt
    Aref = Sref[m:, m:] - Sref[m:, :m] @ np.linalg.pinv(Sref[:m, :m], rcond=0, hermitian=True) @ Sref[:m, m:]
    assert np.linalg.norm(A - Aref) / np.linalg.norm(Aref) < 1e-6


def test_cholesky_export_and_tail_information():
    w = swgn.SynthWindow(2, 2)
    o = ob.OracleSolver(w.graph_p, w.options())
    st, sm = o.minimize()
    n, S, r, L = o.exports()
    Sfull = np.triu(S) + np.triu(S, 1).T
    assert np.linalg.norm(L @ L.T - Sfull) / np.linalg.norm(Sfull) < 1e-12
    nt = w.n_amb
    A = ob.tail_information(L, nt)
    # information of the trailing block = inverse of the trailing block of S^-1
    Aref = np.linalg.inv(np.linalg.inv(Sfull)[n - nt:, n - nt:])
    assert np.linalg.norm(A - Aref) / np.linalg.norm(Aref) < 1e-6


def test_is_use_mask_moves_cost_into_fixed_cost():
    w = swgn.SynthWindow(1, 4)
    g = w.graph
    nfac = g.n_proj + g.n_imu + g.n_gnss + g.n_prior + g.n_unit
    o_all = ob.OracleSolver(w.graph_p, w.options())
    c_all = o_all.evaluate(jac=False)[0]
    mask = np.ones(nfac, np.uint8)
    mask[:g.n_proj // 2] = 0  # drop the first half of the projection factors
    # every landmark must keep a residual: keep factors whose landmark would otherwise vanish
    lm_seen = {}
    for i in range(g.n_proj):
        lm_seen.setdefault(g.proj_blocks[3 * i + 2], []).append(i)
    for lm, idx in lm_seen.items():
        if all(mask[i] == 0 for i in idx):
            mask[idx[0]] = 1
    g.is_use = mask.ctypes.data_as(C.POINTER(C.c_uint8))
    try:
        o = ob.OracleSolver(w.graph_p, w.options())
        assert o.n_res < o_all.n_res
        st, sm = o.minimize()
        c_used = o_all and sm.initial_cost
        assert abs(c_used - c_all) <= 1e-9 * c_all  # initial cost includes fixed_cost
        assert sm.fixed_cost > 0
    finally:
        g.is_use = None


def test_dependent_first_group_is_rejected():
    w = swgn.SynthWindow(1, 0)
    g = w.graph
    groups = np.ctypeslib.as_array(g.block_group, shape=(g.n_blocks,))
    saved = groups.copy()
    try:
        groups[0] = 0  # pose 0 into the e-group: shares residuals with landmarks
        with pytest.raises(RuntimeError):
            ob.OracleSolver(w.graph_p, w.options())
    finally:
        groups[:] = saved


def test_ambiguity_fix_on_synthetic_covariance():
    """Decision pipeline on a well-conditioned float solution: D rows, ratio test, integers."""
    rng = np.random.default_rng(11)
    n = 20
    zt = rng.integers(-50, 50, size=n).astype(float)
    B = rng.normal(size=(n, n))
    Qy = (B @ B.T + n * np.eye(n)) * 1e-4
    A = np.linalg.inv(Qy)
    y = zt + rng.multivariate_normal(np.zeros(n), Qy)
    sysfreq = np.array([0] * 8 + [2] * 7 + [4] * 5, np.int32)
    eb = np.array([0, n, 2 * n], np.int32)
    oa = np.concatenate([np.arange(n), np.arange(n)]).astype(np.int32)
    sf = np.concatenate([sysfreq, sysfreq])
    pairs, F, res = ob.ambiguity_fix(A, y, eb, oa, sf)
    assert res.status == 0 and res.n_dd == n - 3  # one reference satellite per system
    assert res.search_ok == 1
    assert np.array_equal(F[:, 0], zt[pairs[:, 0]] - zt[pairs[:, 1]])
    for a_, b_ in pairs:
        assert sysfreq[a_] == sysfreq[b_]
    # too few ambiguities: no search (swf_lambda.cpp:96-99)
    p2, F2, r2 = ob.ambiguity_fix(A[:5, :5], y[:5], np.array([0, 5], np.int32), np.arange(5, dtype=np.int32), sysfreq[:5])
    assert r2.status == 1
