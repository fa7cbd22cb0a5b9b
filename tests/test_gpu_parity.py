"""Parity of the CUDA path (through the C ABI of include/swgn.h) against the CPU oracle and the
reference's golden fixtures.  Tolerances: the reference is not bit-reproducible with itself
(pointer-ordered chunks, mutex-ordered S updates: SURVEY.md 7 'parity definition'), so floating
point stages are compared at stated fp64 tolerances; structure, iteration counts and the integer
ambiguity decision are compared exactly."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import swgn
from linear_graph import LinearGraph

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

# stated fp64 tolerances
TOL_RES = 1e-11      # residual vector, relative 2-norm
TOL_JAC = 1e-12      # Jacobian, relative Frobenius norm
TOL_S = 1e-12        # reduced system
TOL_STATE = 1e-6     # state vector after the full solve: max |dx| / max(1, |x|), any block
TOL_COST = 1e-6      # final cost, relative
# ... and per kind of parameter block, ~10x what the B200 shows on the BASELINE window after 8 iterations (the landmarks carry
# the largest difference: 5.9e-8; tools/gpu_stage_check.py prints the table).  cond(J'J + D^2) of these windows is 1e16..1e17,
# so the differences are amplified rounding of an ill-conditioned solve, not algorithmic ones.
TOL_STATE_KIND = {"position": 1e-7, "quaternion": 1e-8, "speed": 1e-7, "acc bias": 2e-7, "gyro bias": 1e-8, "landmark": 6e-7,
                  "ambiguity": 1e-7, "scalar": 2e-7}


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def state_err(x, xo):
    return float(np.max(np.abs(x - xo) / np.maximum(1.0, np.abs(xo))))


@pytest.mark.parametrize("which,wid", [(1, 0), (1, 5), (2, 0), (2, 1), (2, 9)])
def test_structure_and_evaluation(which, wid):
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    o = ob.OracleSolver(w.graph_p, opt)
    b = swgn.Batch([w.graph_p], opt)
    assert all(np.array_equal(x, y) for x, y in zip(b.columns(0), o.columns()))
    assert all(np.array_equal(x, y) for x, y in zip(b.rows(0), o.rows()))
    cost, r, g = b.evaluate(0, o.n_res, o.n_cols)
    ocost, orr, og, oJ = o.evaluate()
    J = b.dense_jacobian(0, o.n_res, o.n_cols)
    assert abs(cost - ocost) <= 1e-12 * ocost
    assert rel(r, orr) < TOL_RES
    assert rel(J, oJ) < TOL_JAC
    assert rel(g, og) < 1e-11
    b.close()


@pytest.mark.parametrize("which,wid", [(1, 0), (2, 0), (2, 4)])
def test_linear_solve_stage(which, wid):
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    o = ob.OracleSolver(w.graph_p, opt)
    b = swgn.Batch([w.graph_p], opt)
    rng = np.random.default_rng(wid)
    D = rng.uniform(0.5, 1.5, o.n_cols) * 1e-2
    x = b.linear_solve(0, D, o.n_cols)
    st, ox, oS, orhs = o.linear_solve(D)
    assert st == 0
    S, rhs = b.get_reduced(0)
    # the per-linear-solve gate (1e-9 on identical linearisation) is applied where it is meaningful:
    # the reduced system the eliminator produces, and the backward error of the solution
    assert rel(np.triu(S), np.triu(oS)) < TOL_S
    assert rel(rhs, orhs) < 1e-12
    _, r, g, J = o.evaluate()
    H = J.T @ J + np.diag(D * D)
    nH = np.linalg.norm(H, 2)

    def backward(v):
        return float(np.linalg.norm(H @ v - g) / (nH * np.linalg.norm(v) + np.linalg.norm(g)))
    assert backward(x) < 1e-16 and backward(ox) < 1e-16
    # forward error: cond(J'J + D^2) is 1e15..1e17 here (cond * eps ~ 1..40), the two solutions agree far better than that bound
    assert rel(x, ox) < 1e-7
    b.close()


@pytest.mark.parametrize("which,wid", [(1, 0), (1, 3), (2, 0), (2, 2), (2, 11)])
def test_full_solve_matches_oracle(which, wid):
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    b = swgn.Batch([w.graph_p], opt)
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    o = ob.OracleSolver(w.graph_p, opt)
    st, osm = o.minimize()
    assert st == 0
    assert sm.num_iterations == osm.num_iterations
    assert sm.num_successful_steps == osm.num_successful_steps
    assert sm.num_unsuccessful_steps == osm.num_unsuccessful_steps
    assert sm.termination_type == osm.termination_type
    assert sm.num_linear_solves == osm.num_linear_solves
    assert (sm.n_e, sm.n_f, sm.n_residuals) == (osm.n_e, osm.n_f, osm.n_residuals)
    assert abs(sm.initial_cost - osm.initial_cost) <= 1e-11 * osm.initial_cost
    assert abs(sm.final_cost - osm.final_cost) <= TOL_COST * osm.final_cost
    assert state_err(x, o.state()) < TOL_STATE
    for kind, err in swgn.state_error_by_kind(w, x, o.state()).items():
        assert err < TOL_STATE_KIND[kind], (kind, err)
    b.close()


VARIANTS = [swgn.SYNTH_SPP, swgn.SYNTH_FIXED_INTEGER, swgn.SYNTH_FREE_EXTRINSIC,
            swgn.SYNTH_SPP | swgn.SYNTH_FIXED_INTEGER | swgn.SYNTH_FREE_EXTRINSIC]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("shape", [dict(n_keyframes=8, n_landmarks=60, n_gnss_epochs=4, n_sats=10), {}])
def test_factor_kinds_the_default_window_does_not_hold(variant, shape):
    """SppPseudorangeFactor, SppCarrierPhaseFactor, FixedIntegerFactor (gnss_factor.cpp:9-96) and projection factors with a
    free camera extrinsic (projection_factor.cpp:50-57, ESTIMATE_EXTRINSIC): evaluation, one linear solve, full solve."""
    w = swgn.SynthWindow(2, 3, variant=variant, **shape)
    opt = w.options()
    kinds = set(np.ctypeslib.as_array(w.graph.gnss_kind, shape=(w.graph.n_gnss,)).tolist())
    if variant & swgn.SYNTH_SPP:
        assert {0, 1} <= kinds and not ({2, 3} & kinds)
    if variant & swgn.SYNTH_FIXED_INTEGER:
        assert 5 in kinds
    o = ob.OracleSolver(w.graph_p, opt)
    b = swgn.Batch([w.graph_p], opt)
    assert all(np.array_equal(x, y) for x, y in zip(b.columns(0), o.columns()))
    assert all(np.array_equal(x, y) for x, y in zip(b.rows(0), o.rows()))
    if variant & swgn.SYNTH_FREE_EXTRINSIC:
        ext_block = 2 * w.n_frames  # the generator's block layout: poses, speed-biases, extrinsic
        assert ext_block in set(b.columns(0)[0].tolist())
    cost, r, g = b.evaluate(0, o.n_res, o.n_cols)
    ocost, orr, og, oJ = o.evaluate()
    J = b.dense_jacobian(0, o.n_res, o.n_cols)
    assert abs(cost - ocost) <= 1e-12 * ocost
    assert rel(r, orr) < TOL_RES and rel(J, oJ) < TOL_JAC and rel(g, og) < 1e-11
    D = np.random.default_rng(variant).uniform(0.5, 1.5, o.n_cols) * 1e-2
    x = b.linear_solve(0, D, o.n_cols)
    st, ox, oS, orhs = o.linear_solve(D)
    S, rhs = b.get_reduced(0)
    assert rel(np.triu(S), np.triu(oS)) < TOL_S and rel(rhs, orhs) < 1e-12 and rel(x, ox) < 1e-6
    sm = b.solve()[0]
    st, osm = o.minimize()
    assert (sm.num_iterations, sm.num_successful_steps, sm.termination_type, sm.num_linear_solves) == \
        (osm.num_iterations, osm.num_successful_steps, osm.termination_type, osm.num_linear_solves)
    assert abs(sm.final_cost - osm.final_cost) <= TOL_COST * osm.final_cost
    assert state_err(b.get_state(0, w.n_state), o.state()) < TOL_STATE
    b.close()


@pytest.mark.parametrize("which,wid,jacobi,iters", [(1, 0, 0, 8), (1, 4, 1, 8), (2, 0, 0, 8), (2, 6, 1, 8), (2, 1, 1, 20)])
def test_levenberg_marquardt_solve_matches_oracle(which, wid, jacobi, iters):
    """Ceres' default strategy (levenberg_marquardt_strategy.cc:66-165) with and without jacobi_scaling on the synthetic
    windows: same accept / reject sequence and termination as the oracle, state within the dogleg tolerances."""
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    opt.trust_region_strategy = 1
    opt.jacobi_scaling = jacobi
    opt.max_num_iterations = iters
    b = swgn.Batch([w.graph_p], opt)
    sm = b.solve()[0]
    o = ob.OracleSolver(w.graph_p, opt)
    st, osm = o.minimize()
    assert st == 0
    assert (sm.num_iterations, sm.num_successful_steps, sm.num_unsuccessful_steps, sm.termination_type, sm.num_linear_solves) == \
        (osm.num_iterations, osm.num_successful_steps, osm.num_unsuccessful_steps, osm.termination_type, osm.num_linear_solves)
    assert abs(sm.final_cost - osm.final_cost) <= TOL_COST * osm.final_cost
    assert state_err(b.get_state(0, w.n_state), o.state()) < TOL_STATE
    b.close()


def test_converged_solve_and_early_termination():
    """More iterations than needed: function tolerance stops both paths at the same iteration."""
    w = swgn.SynthWindow(1, 2)
    opt = w.options()
    opt.max_num_iterations = 40
    b = swgn.Batch([w.graph_p], opt)
    sm = b.solve()[0]
    o = ob.OracleSolver(w.graph_p, opt)
    st, osm = o.minimize()
    assert osm.termination_type == 0 and sm.termination_type == 0
    assert sm.num_iterations == osm.num_iterations < 40
    assert state_err(b.get_state(0, w.n_state), o.state()) < TOL_STATE
    b.close()


def test_golden_ceres_problems_through_the_abi():
    """CERES linear_least_squares_problems.cc problems 2 and 4 (hand-computed S, r, S\\r, A\\b)."""
    with open(os.path.join(HERE, "golden", "ceres_llsq_problems.json")) as f:
        fx = json.load(f)
    opt = swgn.default_options()
    # problem 2, D = 0: the numbers of the reference's comment (:135-178)
    p = fx["problem2"]
    lg = LinearGraph(p)
    b = swgn.Batch([lg.graph_p], opt)
    x = b.linear_solve(0, None, lg.n_cols)
    S, r = b.get_reduced(0)
    g = p["golden"]
    Sfull = np.triu(S) + np.triu(S, 1).T
    np.testing.assert_allclose(Sfull, np.array(g["S"]), atol=6e-5)
    r_doc = np.array(g["r"])
    r_doc[2] = 4.0323  # typo in the reference's comment, see test_oracle_golden.py
    np.testing.assert_allclose(r, r_doc, atol=6e-5)
    np.testing.assert_allclose(x, np.array(g["A_solve_b"]), atol=6e-5)
    b.close()
    # problems 2 and 4 with their D against the dense construction of schur_eliminator_test.cc
    for name in ("problem2", "problem4"):
        p = fx[name]
        lg = LinearGraph(p)
        b = swgn.Batch([lg.graph_p], opt)
        D = np.array(p["D"], float)
        x = b.linear_solve(0, D, lg.n_cols)
        S, r = b.get_reduced(0)
        from test_oracle_golden import schur_reference
        S_ref, r_ref, x_ref = schur_reference(p, True)
        Sfull = np.triu(S) + np.triu(S, 1).T
        assert rel(Sfull, S_ref) < 1e-14
        assert rel(r, r_ref) < 1e-14
        assert rel(x, x_ref) < 1e-13
        b.close()


def test_export_mode():
    """is_optimize = 0 with a parameter head: one evaluation + one Eliminate, S and r exported,
    the state left untouched (CERES schur_complement_solver.cc:172-188)."""
    w = swgn.SynthWindow(2, 1)
    opt = w.options()
    opt.is_optimize = 0
    opt.max_num_iterations = 1
    b = swgn.Batch([w.graph_p], opt)
    x0 = b.get_state(0, w.n_state)
    sm = b.solve()[0]
    assert np.array_equal(b.get_state(0, w.n_state), x0)
    o = ob.OracleSolver(w.graph_p, opt)
    st, osm = o.minimize()
    assert (sm.num_iterations, sm.num_successful_steps, sm.num_unsuccessful_steps, sm.termination_type) == \
        (osm.num_iterations, osm.num_successful_steps, osm.num_unsuccessful_steps, osm.termination_type)
    S, r = b.get_reduced(0)
    n, oS, orr, _ = o.exports()
    assert S.shape[0] == n
    assert rel(np.triu(S), np.triu(oS)) < TOL_S
    assert rel(r, orr) < 1e-9
    b.close()


@pytest.mark.parametrize("kf,lm", [(10, 100), (10, 1000), (40, 100), (40, 1000)])
def test_window_size_sweep_corners_match_oracle(kf, lm):
    """BASELINE configs[4] / SURVEY 8d cfg5: the corners of the keyframes x landmarks sweep (GNSS epochs =
    KF / 2) solved on the device and by the oracle."""
    w = swgn.SynthWindow(2, 1, n_keyframes=kf, n_landmarks=lm, n_gnss_epochs=kf // 2)
    opt = w.options()
    b = swgn.Batch([w.graph_p], opt)
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    o = ob.OracleSolver(w.graph_p, opt)
    st, osm = o.minimize()
    assert st == 0
    assert (sm.num_iterations, sm.num_successful_steps, sm.num_unsuccessful_steps, sm.termination_type) == \
        (osm.num_iterations, osm.num_successful_steps, osm.num_unsuccessful_steps, osm.termination_type)
    assert abs(sm.final_cost - osm.final_cost) <= TOL_COST * osm.final_cost
    assert state_err(x, o.state()) < TOL_STATE
    assert sm.n_f == o.n_f and sm.n_e == o.n_e
    b.close()


@pytest.mark.parametrize("which,wid,n_head", [(2, 1, None), (1, 0, 2), (2, 3, 5)])
def test_update_schur_on_the_device(which, wid, n_head):
    """UpdateSchur (RVI/swf/swf_gnss.cpp:25-61) after the export-mode solve: (A, b) of the head blocks by
    the eigen pseudo-inverse Schur reduction, against the oracle's restatement, and the defining
    property: A z_n = b is the head part of the solution of the full reduced system S z = r."""
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    if n_head is not None:
        opt.n_parameter_head = n_head
    opt.is_optimize = 0
    opt.max_num_iterations = 1
    b = swgn.Batch([w.graph_p], opt)
    b.solve()
    S, r = b.get_reduced(0)
    st, d = swgn.plan_probe(w.graph_p, opt.n_parameter_head)
    cb, co, cs = b.columns(0)
    n_tail = int(cs[len(cs) - opt.n_parameter_head:].sum())
    A, bv = b.head_marginal(0, n_tail)
    Sf = np.triu(S) + np.triu(S, 1).T
    Ao, bo = ob.update_schur(Sf, r, n_tail)
    # A = A_nn - A_nm A_mm^+ A_mn cancels the leading digits of A_nn (cfg1: |A_nn| ~ 1e9, |A| ~ 1e4), so two
    # correct evaluations agree to ~1e-16 * |A_nn| / |A|
    assert rel(A, Ao) < 1e-7 and rel(bv, bo) < 1e-6
    assert np.allclose(A, A.T, rtol=0, atol=1e-9 * np.abs(A).max())
    z = np.linalg.solve(Sf, r)
    assert rel(np.linalg.solve(A, bv), z[-n_tail:]) < 1e-6
    b.close()


@pytest.mark.parametrize("which,wid,n_head", [(2, 1, None), (1, 0, 2)])
def test_marginal_prior_on_the_device(which, wid, n_head):
    """Export-mode solve -> UpdateSchur -> setmarginalizeinfo (marginalization_factor.cpp:449-475): the prior
    factor (J0, r0) of the head blocks.  J0 = sqrt(S) V' is defined up to the sign / order of the eigenvectors,
    so J0'J0 = A and J0'r0 = b are checked, against the device's own (A, b) and against the oracle's factor."""
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    if n_head is not None:
        opt.n_parameter_head = n_head
    opt.is_optimize = 0
    opt.max_num_iterations = 1
    b = swgn.Batch([w.graph_p], opt)
    b.solve()
    cb, co, cs = b.columns(0)
    n_tail = int(cs[len(cs) - opt.n_parameter_head:].sum())
    J0, r0, A, bv = b.marginal_prior(0, n_tail)
    A2, b2 = b.head_marginal(0, n_tail)
    assert np.array_equal(A, A2) and np.array_equal(bv, b2)
    scale = np.abs(A).max()
    Au = np.triu(A) + np.triu(A, 1).T  # SelfAdjointEigenSolver reads one triangle; A itself is symmetric to rounding only
    assert np.abs(J0.T @ J0 - Au).max() < 1e-9 * scale
    assert np.abs(J0.T @ r0 - bv).max() < 1e-8 * max(1.0, np.abs(bv).max())
    oJ, orr = ob.prior_sqrt(Au, bv)
    assert np.abs(oJ.T @ oJ - J0.T @ J0).max() < 1e-9 * scale
    assert abs(orr @ orr - r0 @ r0) < 1e-8 * (orr @ orr)
    # the prior drives the same solution as the information form: argmin |r0 + J0 d|^2 = -A^-1 b
    d = np.linalg.lstsq(J0, -r0, rcond=None)[0]
    assert rel(d, -np.linalg.solve(A, bv)) < 1e-6
    b.close()


def test_marginal_priors_of_a_whole_batch_equal_the_per_window_calls():
    ws = [swgn.SynthWindow(2, 1), swgn.SynthWindow(2, 3), swgn.SynthWindow(2, 5)]
    opt = ws[0].options()
    opt.is_optimize = 0
    opt.max_num_iterations = 1
    b = swgn.Batch([w.graph_p for w in ws], opt)
    b.solve()
    nt = []
    for k in range(3):
        cb, co, cs = b.columns(k)
        nt.append(int(cs[len(cs) - opt.n_parameter_head:].sum()))
    nt[1] = 0  # skipped window
    got = b.marginal_priors(nt)
    assert got[1] == (None, None)
    for k in (0, 2):
        J0, r0, A, bv = b.marginal_prior(k, nt[k])
        assert np.array_equal(got[k][0], J0) and np.array_equal(got[k][1], r0)
    b.close()


def test_cholesky_export_and_tail_information():
    w = swgn.SynthWindow(2, 2)
    opt = w.options()
    b = swgn.Batch([w.graph_p], opt)
    b.solve()
    o = ob.OracleSolver(w.graph_p, opt)
    o.minimize()
    n, oS, orr, Lo = o.exports()
    Lg = b.get_cholesky(0)
    assert np.all(np.triu(Lg, 1) == 0)
    assert rel(Lg, Lo) < 1e-8
    A = b.tail_information(0, w.n_amb)
    Ao = ob.tail_information(Lo, w.n_amb)
    assert rel(A, Ao) < 1e-8
    assert np.allclose(A, A.T, rtol=0, atol=1e-9 * np.abs(A).max())
    b.close()


def test_is_use_mask():
    w = swgn.SynthWindow(1, 4)
    g = w.graph
    nfac = g.n_proj + g.n_imu + g.n_gnss + g.n_prior + g.n_unit
    mask = np.ones(nfac, np.uint8)
    seen = set()
    for i in range(g.n_proj):  # keep the first observation of every landmark, drop every other one
        lm = g.proj_blocks[3 * i + 2]
        if lm in seen and i % 2 == 0:
            mask[i] = 0
        seen.add(lm)
    g.is_use = mask.ctypes.data_as(C.POINTER(C.c_uint8))
    try:
        opt = w.options()
        b = swgn.Batch([w.graph_p], opt)
        sm = b.solve()[0]
        o = ob.OracleSolver(w.graph_p, opt)
        st, osm = o.minimize()
        assert sm.n_residuals == osm.n_residuals < 2 * g.n_proj + 15 * g.n_imu + 16
        assert abs(sm.fixed_cost - osm.fixed_cost) <= 1e-11 * osm.fixed_cost and sm.fixed_cost > 0
        assert abs(sm.final_cost - osm.final_cost) <= TOL_COST * osm.final_cost
        # half of the observations are masked out: landmarks seen once or twice are barely
        # constrained, so rounding-level differences reach the state at a larger factor
        assert state_err(b.get_state(0, w.n_state), o.state()) < 50 * TOL_STATE
        b.close()
    finally:
        g.is_use = None


def test_batch_equals_individual_windows():
    """Windows are independent: solving them in one batch gives bit-identical states to solving
    each alone (the kernels are deterministic), and every one matches the oracle."""
    ws = [swgn.SynthWindow(2, i) for i in range(6)] + [swgn.SynthWindow(1, i) for i in range(3)]
    opt = ws[0].options()
    opt.n_parameter_head = 0  # the VI-only windows have no ambiguity blocks to hold back
    b = swgn.Batch([w.graph_p for w in ws], opt)
    sms = b.solve()
    packed = b.get_states()
    off = 0
    for i, w in enumerate(ws):
        x = b.get_state(i, w.n_state)
        assert np.array_equal(x, packed[off:off + w.n_state])
        off += w.n_state
        if i in (0, 5, 7):
            o1 = w.options()
            o1.n_parameter_head = opt.n_parameter_head
            b1 = swgn.Batch([w.graph_p], opt)
            sm1 = b1.solve()[0]
            assert np.array_equal(b1.get_state(0, w.n_state), x)
            assert sm1.final_cost == sms[i].final_cost
            b1.close()
    for i in (1, 8):
        o = ob.OracleSolver(ws[i].graph_p, opt)
        o.minimize()
        assert state_err(b.get_state(i, ws[i].n_state), o.state()) < TOL_STATE
    # restarting from re-uploaded initial states reproduces the result
    b.set_states(np.concatenate([w.state0() for w in ws]))
    b.solve()
    assert np.array_equal(b.get_states(), packed)
    b.close()


def test_lambda_batch_bit_exact():
    """cfg4 (K8): ragged batch of lambda() problems, integer candidates and squared norms
    bit-identical to the oracle (itself pinned bit-exactly on the reference's lambda.cpp)."""
    rng = np.random.default_rng(2026)
    ns, As, Qs, exp = [], [], [], []
    for k in range(200):
        n = int(rng.integers(2, 31))
        B = rng.normal(size=(n, n + 2))
        u = rng.normal(size=(n, 1))
        Q = (B @ B.T + 40.0 * (u @ u.T)) * 10.0 ** rng.integers(-3, 1) + 1e-4 * np.eye(n)
        Q = 0.5 * (Q + Q.T)
        a = rng.uniform(-30, 30, size=n)
        ns.append(n)
        As.append(a)
        Qs.append(np.asfortranarray(Q).ravel(order="F"))
        exp.append(ob.lambda_search(a, Q, 2, "oracle"))
    F, s, info = swgn.lambda_batch(ns, 2, np.concatenate(As), np.concatenate(Qs))
    o = 0
    for k, n in enumerate(ns):
        ei, eF, es = exp[k]
        assert info[k] == ei
        if ei == 0:
            Fk = F[2 * o:2 * o + 2 * n].reshape(2, n).T
            assert np.array_equal(Fk, eF)
            assert np.array_equal(s[2 * k:2 * k + 2], es)
        o += n
    # failure path: non positive-definite covariance
    F, s, info = swgn.lambda_batch([3], 2, np.array([0.1, 0.2, 0.3]), (-np.eye(3)).ravel())
    assert info[0] == -1


@pytest.mark.parametrize("wid", [0, 1, 2, 3])
def test_ambiguity_fix_decision_bit_exact(wid):
    """cfg4: covariance recovery + LAMBDA ambiguity fix on the solved BASELINE window; the decision
    (double-difference rows, integer vectors, squared norms, ratio test) is bit-exact given the
    same A, y; with the GPU's own A, y the integers and the decision are identical."""
    w = swgn.SynthWindow(2, wid)
    opt = w.options()
    b = swgn.Batch([w.graph_p], opt)
    b.solve()
    o = ob.OracleSolver(w.graph_p, opt)
    o.minimize()
    nt = w.n_amb
    n, oS, orr, Lo = o.exports()
    Ao = ob.tail_information(Lo, nt)
    offs = w.block_offsets()
    xo, xg = o.state(), b.get_state(0, w.n_state)
    yo = np.array([xo[offs[w.first_amb_block + k]] for k in range(nt)])
    yg = np.array([xg[offs[w.first_amb_block + k]] for k in range(nt)])
    eb, oa, sf = w.ambiguity_epochs()
    for last_fix in (0, 1):
        po, Fo, ro = ob.ambiguity_fix(Ao, yo, eb, oa, sf, last_fix)
        pg, Fg, rg = swgn.ambiguity_fix(Ao, yo, eb, oa, sf, last_fix)
        assert (rg.status, rg.n_dd, rg.search_ok, rg.n_different) == (ro.status, ro.n_dd, ro.search_ok, ro.n_different)
        assert np.array_equal(pg, po) and np.array_equal(Fg, Fo)
        assert list(rg.s) == list(ro.s) and rg.s0_partial == ro.s0_partial and rg.s1_partial == ro.s1_partial
    Ag = b.tail_information(0, nt)
    pg, Fg, rg = swgn.ambiguity_fix(Ag, yg, eb, oa, sf, 0)
    po, Fo, ro = ob.ambiguity_fix(Ao, yo, eb, oa, sf, 0)
    # The GPU state differs from the oracle's at the 1e-7 level (stated tolerance), which may flip
    # the near-tied choice of a reference satellite (swf_lambda.cpp:24-52); the decision itself --
    # ratio test and the fixed integer relations between ambiguities -- must be the same.  Compare
    # the integers in a reference-independent form: N_a - N_c for all a, c of one group.
    assert rg.search_ok == ro.search_ok and rg.n_dd == ro.n_dd

    def relations(pairs, F):
        groups = {}
        for (a, ref), f in zip(pairs, np.round(F[:, 0])):
            groups.setdefault(int(ref), {int(ref): 0.0})[int(a)] = float(f)
        out = {}
        for ref, g in groups.items():
            base = min(g)
            out[frozenset(g)] = {a: v - g[base] for a, v in g.items()}
        return out

    assert relations(pg, Fg) == relations(po, Fo)
    b.close()


def test_full_size_batch_properties():
    """At BASELINE batch shape (many 20-KF windows) the oracle is too slow to check every window;
    size-independent properties instead: every window reduces its cost by orders of magnitude, the
    exported Cholesky factor reproduces a symmetric positive-definite S, a few sampled windows
    match the oracle."""
    n = 96
    ws = [swgn.SynthWindow(2, 1000 + i) for i in range(n)]
    opt = ws[0].options()
    b = swgn.Batch([w.graph_p for w in ws], opt)
    sms = b.solve()
    for i in range(n):
        assert sms[i].termination_type in (0, 1)
        assert sms[i].final_cost < 1e-4 * sms[i].initial_cost
        assert sms[i].num_iterations <= 8
    for i in (0, 50, 95):
        L = b.get_cholesky(i)
        assert np.all(np.diag(L) > 0)
        o = ob.OracleSolver(ws[i].graph_p, opt)
        o.minimize()
        assert state_err(b.get_state(i, ws[i].n_state), o.state()) < TOL_STATE
    tot, schur, nl, kl = b.timing()
    assert tot > 0 and schur > 0 and nl >= 8 and kl > nl
    b.close()


def test_recycled_slabs_do_not_leak_state_between_batches():
    """swgn_batch_destroy parks the batch's device / pinned slabs in a process-wide cache and later creates
    reuse them (batch.cpp slab_alloc): a batch built on recycled memory must give bit-identical results to
    the one built on fresh memory, whatever the previous tenant left behind (other window, other size,
    export mode, chains)."""
    a, c = swgn.SynthWindow(2, 3), swgn.SynthWindow(1, 1)
    ch = swgn.SynthWindow(4, 0)

    def run(w, **kw):
        opt = w.options()
        for k, v in kw.items():
            setattr(opt, k, v)
        b = swgn.Batch([w.graph_p], opt)
        sm = b.solve()[0]
        x = b.get_state(0, w.n_state).copy()
        b.close()
        return x, sm.final_cost, sm.num_iterations

    first = run(a)
    for other, kw in ((c, {"n_parameter_head": 0}), (ch, {}), (a, {"is_optimize": 0}), (c, {"n_parameter_head": 0})):
        run(other, **kw)
        again = run(a)
        assert np.array_equal(again[0], first[0]) and again[1] == first[1] and again[2] == first[2]
    # the parked slabs can be handed back, and a batch built afterwards (fresh memory again) agrees as well
    assert swgn.release_cached_memory() > 0
    assert swgn.release_cached_memory() == 0
    again = run(a)
    assert np.array_equal(again[0], first[0]) and again[1] == first[1]


def test_window_without_retained_blocks_solves_like_the_oracle():
    """Edge of the reduced program: every f-block constant, so the reduced system is empty (n_f = 0) and the
    solve is Schur chunks + back-substitution only (problem 2 of the reference's llsq fixtures, e-blocks free)."""
    with open(os.path.join(HERE, "golden", "ceres_llsq_problems.json")) as f:
        p = json.load(f)["problem2"]
    lg = LinearGraph(p)
    lg.block_const[p["num_eliminate_blocks"]:] = 1
    opt = swgn.default_options()
    o = ob.OracleSolver(lg.graph_p, opt)
    assert o.n_f == 0 and o.n_e == 2
    _, osm = o.minimize()
    b = swgn.Batch([lg.graph_p], opt)
    sm = b.solve()[0]
    x = b.get_state(0, lg.n_cols)
    b.close()
    assert sm.termination_type == osm.termination_type and sm.num_iterations == osm.num_iterations
    assert abs(sm.final_cost - osm.final_cost) <= 1e-12 * osm.final_cost
    np.testing.assert_allclose(x, o.state(), rtol=0, atol=1e-12)


def test_update_inputs_rejects_a_graph_with_another_structure():
    """Same factor counts, different structure (two projection factors exchange their landmarks; a block turns constant; a
    NULL graph): swgn_batch_update_inputs must answer SWGN_ERR_INVALID instead of repacking against the old tables."""
    w = swgn.SynthWindow(1, 0)
    b = swgn.Batch([w.graph_p], w.options())
    assert b.update_inputs([w.graph_p]) > 0  # unchanged structure: accepted
    w2 = swgn.SynthWindow(1, 0)
    g = w2.graph
    k = next(i for i in range(1, g.n_proj) if g.proj_blocks[3 * i + 2] != g.proj_blocks[2])
    g.proj_blocks[2], g.proj_blocks[3 * k + 2] = g.proj_blocks[3 * k + 2], g.proj_blocks[2]
    with pytest.raises(RuntimeError, match="structure"):
        b.update_inputs([w2.graph_p])
    g.proj_blocks[2], g.proj_blocks[3 * k + 2] = g.proj_blocks[3 * k + 2], g.proj_blocks[2]
    assert b.update_inputs([w2.graph_p]) > 0
    g.block_const[0] = 1
    with pytest.raises(RuntimeError, match="structure"):
        b.update_inputs([w2.graph_p])
    g.block_const[0] = 0
    null = (C.POINTER(swgn.Graph) * 1)()
    nb = C.c_int64()
    assert swgn.lib().swgn_batch_update_inputs(b.h, null, C.byref(nb)) == 1  # SWGN_ERR_INVALID
    sm = b.solve()[0]  # the batch is still usable
    assert sm.termination_type in (0, 1)
    b.close()


def test_batched_ambiguity_fix_equals_the_per_window_calls():
    """swgn_batch_ambiguity_fix: tail information + float ambiguities + LambdaSearch decision of every window in two
    launches.  Bit-identical to the per-window entry points (swgn_batch_get_tail_information + swgn_ambiguity_fix) fed
    with the same device data, for both values of last_fix; a window without a factor reports status 3."""
    ws = [swgn.SynthWindow(2, wid) for wid in (0, 1, 2, 3, 7, 12)]
    opt = ws[0].options()
    b = swgn.Batch([w.graph_p for w in ws], opt)
    b.solve()
    nt = ws[0].n_amb
    epochs = [w.ambiguity_epochs() for w in ws]
    for last in (0, 1):
        res, pairs, F = b.ambiguity_fix_all(nt, epochs, [last] * len(ws))
        for i, w in enumerate(ws):
            offs = w.block_offsets()
            x = b.get_state(i, w.n_state)
            y = np.array([x[offs[w.first_amb_block + k]] for k in range(nt)])
            A = b.tail_information(i, nt)
            p1, F1, r1 = swgn.ambiguity_fix(A, y, *epochs[i], last_fix=last)
            r = res[i]
            assert (r.status, r.n_dd, r.search_ok, r.n_different) == (r1.status, r1.n_dd, r1.search_ok, r1.n_different)
            assert list(r.s) == list(r1.s) and r.s0_partial == r1.s0_partial and r.s1_partial == r1.s1_partial
            assert np.array_equal(pairs[i, :r.n_dd], p1)
            assert np.array_equal(F[i].ravel()[:2 * r.n_dd].reshape(2, r.n_dd).T, F1)
    assert any(res[i].status == 0 for i in range(len(ws)))
    b.close()
    # before any solve there is no Cholesky factor: every window answers status 3
    b = swgn.Batch([ws[0].graph_p], opt)
    res, _, _ = b.ambiguity_fix_all(nt, epochs[:1])
    assert res[0].status == 3
    b.close()


def test_prefetched_inputs_give_the_same_solve_as_a_fresh_batch():
    """swgn_batch_prefetch_inputs (from a second host thread, while a solve is running) + swgn_batch_commit_inputs: the next
    solve must be bit-identical to a fresh batch created from the new inputs; twice, so that both shadow blocks are used."""
    import threading
    ws = [swgn.SynthWindow(2, wid) for wid in (0, 1)]
    opt = ws[0].options()
    b = swgn.Batch([w.graph_p for w in ws], opt)
    b.solve()
    rng = np.random.default_rng(5)
    for step in range(3):
        nxt = [swgn.SynthWindow(2, wid) for wid in (0, 1)]  # same structure, new measurements and initial states
        for w in nxt:
            g = w.graph
            uv = np.ctypeslib.as_array(g.proj_uv, shape=(2 * g.n_proj,))
            uv += rng.normal(size=uv.shape) * 1e-4
            st = np.ctypeslib.as_array(g.state, shape=(g.n_state,))
            offs = w.block_offsets()
            for k in range(w.n_amb):
                st[offs[w.first_amb_block + k]] += 0.01 * rng.normal()
        t = threading.Thread(target=lambda: b.prefetch_inputs([w.graph_p for w in nxt]))
        t.start()
        b.solve()  # the current step keeps running on the live inputs
        t.join()
        b.commit_inputs()
        sm = b.solve()
        fresh = swgn.Batch([w.graph_p for w in nxt], opt)
        fsm = fresh.solve()
        for i, w in enumerate(nxt):
            assert sm[i].num_iterations == fsm[i].num_iterations and sm[i].final_cost == fsm[i].final_cost
            assert np.array_equal(b.get_state(i, w.n_state), fresh.get_state(i, w.n_state))
        fresh.close()
    with pytest.raises(RuntimeError, match="prefetch"):
        b.commit_inputs()
    b.close()


def test_host_evaluated_cost_functions_reproduce_the_device_factors():
    """swgn_graph.host_*: residual blocks the device has no kind for are evaluated by the caller's own cost function on the
    host at every evaluation point.  Here the IMU factors of a window are taken out of the device kind and handed over as
    host-evaluated blocks whose callback is the (pinned) CPU restatement of IMUFactor: evaluation, linear solve and the
    full solve must reproduce the all-device run (rows are ordered differently, so to rounding, not bit for bit)."""
    w = swgn.SynthWindow(1, 2)
    opt = w.options()
    ref = swgn.Batch([w.graph_p], opt)
    sm0 = ref.solve()[0]
    x0 = ref.get_state(0, w.n_state)
    ref.close()

    g = w.graph
    n_imu = g.n_imu
    glob = np.array(list(g.Pbg) + list(g.gravity) + list(g.proj_sqrt_info))
    recs = np.ctypeslib.as_array(g.imu_data, shape=(n_imu, 474)).copy()
    sizes = (7, 9, 7, 9)
    calls = {"n": 0, "jac": 0}
    f = ob.oracle().oracle_factor_eval
    f.argtypes = [C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 5

    def host_eval(user, factor, params, residuals, jacobians):
        calls["n"] += 1
        x = np.concatenate([np.ctypeslib.as_array(params[k], shape=(sizes[k],)) for k in range(4)])
        r = np.zeros(15)
        J = np.zeros(15 * 32)
        want_j = bool(jacobians)
        st = f(1, 0, glob.ctypes.data_as(C.POINTER(C.c_double)), recs[factor].ctypes.data_as(C.POINTER(C.c_double)),
               x.ctypes.data_as(C.POINTER(C.c_double)), r.ctypes.data_as(C.POINTER(C.c_double)),
               J.ctypes.data_as(C.POINTER(C.c_double)) if want_j else None)
        C.memmove(residuals, r.ctypes.data, 15 * 8)
        if want_j:
            calls["jac"] += 1
            o = 0
            for k in range(4):
                if jacobians[k]:
                    C.memmove(jacobians[k], J[o:o + 15 * sizes[k]].ctypes.data, 15 * sizes[k] * 8)
                o += 15 * sizes[k]
        return st

    cb = swgn.HOST_EVAL_FN(host_eval)
    g2 = swgn.Graph()
    C.memmove(C.byref(g2), C.byref(g), C.sizeof(swgn.Graph))
    g2.n_imu = 0
    nres = (C.c_int32 * n_imu)(*([15] * n_imu))
    begin = (C.c_int32 * (n_imu + 1))(*[4 * i for i in range(n_imu + 1)])
    g2.n_host = n_imu
    g2.host_nres = nres
    g2.host_blk_begin = begin
    g2.host_blocks = g.imu_blocks
    g2.host_eval = C.cast(cb, C.c_void_p)
    g2.host_user = None
    b = swgn.Batch([C.pointer(g2)], opt)
    o = ob.OracleSolver(w.graph_p, opt)
    cost, _, _ = b.evaluate(0, o.n_res, o.n_cols)
    ocost = o.evaluate()[0]
    assert abs(cost - ocost) <= 1e-12 * ocost
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    b.close()
    assert calls["n"] > 2 * n_imu and calls["jac"] > 0
    assert (sm.num_iterations, sm.num_successful_steps, sm.termination_type) == (sm0.num_iterations, sm0.num_successful_steps, sm0.termination_type)
    assert abs(sm.final_cost - sm0.final_cost) <= 1e-8 * sm0.final_cost
    assert state_err(x, x0) < 1e-7


def _plain_copy(w):
    """The window's graph with every block variable and no program order: the form swgn_marginalize takes (the reference's
    MarginalizationInfo knows neither constant blocks nor an ordering)."""
    g = swgn.Graph()
    C.memmove(C.byref(g), w.graph_p, C.sizeof(swgn.Graph))
    free = np.zeros(g.n_blocks, np.int32)
    g.block_const = free.ctypes.data_as(C.POINTER(C.c_int32))
    g.n_order = 0
    g.order = None
    g.is_use = None
    return g, free


@pytest.mark.parametrize("which,overrides,drop_nothing", [(1, {}, False), (2, dict(n_keyframes=6, n_landmarks=40, n_gnss_epochs=3, n_sats=8), False),
                                                          (1, dict(n_keyframes=3, n_landmarks=12), True)])
def test_marginalize_with_an_arbitrary_drop_set_matches_the_oracle(which, overrides, drop_nothing):
    """swgn_marginalize = MarginalizationInfo::marginalize + getParameterBlocks (marginalization_factor.cpp:260-400) as MargFrames
    uses it: drop the oldest frame's pose and speed-bias and the first landmarks, keep everything else.  The oracle side is
    the restated MarginalizationInfo that tests/test_gnss_epoch.py pins on the reference's own class."""
    ws = [swgn.SynthWindow(which, wid, **overrides) for wid in (0, 1)]
    copies = [_plain_copy(w) for w in ws]
    gps, drops = [], []
    for w, (g, _) in zip(ws, copies):
        drop = np.zeros(g.n_blocks, np.uint8)
        sizes = [g.block_size[b] for b in range(g.n_blocks)]
        if not drop_nothing:                          # (nothing dropped: the square root of the whole information, as
            drop[sizes.index(7)] = 1                  #  InitializeSqrtInfo builds the first prior, swf_core.cpp:479-543)
            drop[sizes.index(9)] = 1                  # first pose, first speed-bias
            lms = [b for b in range(g.n_blocks) if sizes[b] == 3][:8]
            drop[lms] = 1
        gps.append(C.pointer(g))
        drops.append(drop)
    got = swgn.marginalize(gps, drops)
    O = ob.oracle()
    i32, f64, P = C.c_int32, C.c_double, C.POINTER
    O.oracle_marginalize_graph.argtypes = [P(swgn.Graph), P(C.c_uint8), C.c_int, C.c_int, P(i32), P(i32), P(i32), P(i32), P(i32), P(f64), P(f64)]
    for (kb, ki, J0, r0, m), gp, drop in zip(got, gps, drops):
        n = J0.shape[0]
        nk, no, mo = i32(), i32(), i32()
        okb, oki, oJ, orr = np.zeros(len(drop), np.int32), np.zeros(len(drop), np.int32), np.zeros(n * n), np.zeros(n)
        rc = O.oracle_marginalize_graph(gp, drop.ctypes.data_as(P(C.c_uint8)), len(drop), n, C.byref(nk), C.byref(no), C.byref(mo),
                                        okb.ctypes.data_as(P(i32)), oki.ctypes.data_as(P(i32)), oJ.ctypes.data_as(P(f64)), orr.ctypes.data_as(P(f64)))
        assert rc == 0 and no.value == n and mo.value == m and nk.value == len(kb)
        assert np.array_equal(okb[:nk.value], kb) and np.array_equal(oki[:nk.value], ki)
        oJ = oJ.reshape(n, n)
        Ag, Ao = J0.T @ J0, oJ.T @ oJ
        assert np.abs(Ag - Ao).max() < 1e-9 * np.abs(Ao).max()
        bg, bo = J0.T @ r0, oJ.T @ orr
        assert np.abs(bg - bo).max() < 1e-7 * max(1.0, np.abs(bo).max())


def test_window_too_wide_for_shared_memory_is_reported_as_too_large():
    """150 keyframes -> a 1 575-row reduced system: the Cholesky panel plus its vectors exceed 227 KB of shared memory.  The
    library says so (SWGN_ERR_TOO_LARGE = 5 with the size in the message) instead of surfacing a bare CUDA 'invalid argument'."""
    w = swgn.SynthWindow(1, 0, n_keyframes=150, n_landmarks=60)
    opt = w.options()
    opt.n_parameter_head = 0
    with pytest.raises(RuntimeError, match="status 5.*shared memory"):
        swgn.Batch([w.graph_p], opt)
    # the library is still usable afterwards
    w2 = swgn.SynthWindow(1, 0)
    b = swgn.Batch([w2.graph_p], w2.options())
    assert b.solve()[0].termination_type in (0, 1)
    b.close()
