"""The ceres:: source-compatibility layer (include/ceres/*.h + shim/): host bookkeeping semantics
(CPU) and the application-style solve through ceres::Problem / ceres::Solve (GPU), compared with
the same window solved directly through the C ABI and with the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_binding as ob
import swgn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f64 = C.c_double


def demo():
    L = C.CDLL(os.path.join(ROOT, "rtk-visual-inertial-navigation_b200", "libswgn_ceres_demo.so"))
    L.swgn_ceres_selftest.restype = C.c_int
    L.swgn_ceres_demo_solve.restype = C.c_int
    L.swgn_ceres_demo_solve.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(f64), C.POINTER(f64), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(f64), C.POINTER(f64), C.c_char_p, C.c_int]
    return L


def test_problem_bookkeeping_semantics():
    assert demo().swgn_ceres_selftest() == 0


def run_demo(which, wid, export_mode, n_state, n_f):
    state = np.zeros(n_state)
    cost = np.zeros(2)
    steps = (C.c_int * 2)()
    hs = C.c_int()
    mat = np.zeros(max(n_f * n_f, 1))
    rhs = np.zeros(max(n_f, 1))
    msg = C.create_string_buffer(512)
    rc = demo().swgn_ceres_demo_solve(which, wid, export_mode, 0, state.ctypes.data_as(C.POINTER(f64)), cost.ctypes.data_as(C.POINTER(f64)),
                                      steps, C.byref(hs), mat.ctypes.data_as(C.POINTER(f64)), rhs.ctypes.data_as(C.POINTER(f64)), msg, 512)
    return rc, state, cost, list(steps), hs.value, mat, rhs, msg.value.decode()


def test_small_blas_header_on_the_reference_test_cases():
    """include/ceres/small_blas.h (called directly by the application's GNSS-IMU factor) against the cases of
    CERES/internal/ceres/small_blas_test.cc: every block placement, kOperation +1 / -1 / 0, fixed and dynamic."""
    L = demo()
    L.swgn_small_blas_selftest.restype = C.c_int
    assert L.swgn_small_blas_selftest() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("which,wid", [(1, 0), (2, 0), (2, 5)])
def test_solve_through_ceres_api_equals_c_abi_and_oracle(which, wid):
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    o = ob.OracleSolver(w.graph_p, opt)
    st, osm = o.minimize()
    rc, state, cost, steps, hs, mat, rhs, msg = run_demo(which, wid, 0, w.n_state, o.n_f)
    assert rc == osm.termination_type, msg
    b = swgn.Batch([w.graph_p], opt)
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    # same flat graph, same device code: bit-identical to the direct C-ABI solve
    assert np.array_equal(state, x)
    assert cost[1] == sm.final_cost and steps == [sm.num_successful_steps, sm.num_unsuccessful_steps]
    xo = o.state()
    assert np.max(np.abs(state - xo) / np.maximum(1, np.abs(xo))) < 1e-6
    if opt.n_parameter_head > 0:
        assert hs == o.n_f
        Lg = b.get_cholesky(0)
        assert np.array_equal(mat[:hs * hs].reshape(hs, hs), Lg)  # lhs_out2 mirrors the device factor
    b.close()


@pytest.mark.gpu
def test_export_mode_through_side_channel():
    """is_optimize = false + parameter_head: lhs_out / rhs_out / hs_row filled, user state untouched
    (the GlobalMarge / IntegerSolve call pattern, RVI/swf/swf_image.cpp:401-415, swf_gnss.cpp:135-162)."""
    w = swgn.SynthWindow(2, 1)
    opt = w.options()
    opt.is_optimize = 0
    opt.max_num_iterations = 1
    o = ob.OracleSolver(w.graph_p, opt)
    o.minimize()
    n, oS, orr, _ = o.exports()
    rc, state, cost, steps, hs, mat, rhs, msg = run_demo(2, 1, 1, w.n_state, o.n_f)
    assert rc >= 0, msg
    assert np.array_equal(state, w.state0())
    assert hs == n
    S = mat[:n * n].reshape(n, n)
    assert np.linalg.norm(np.triu(S) - np.triu(oS)) / np.linalg.norm(np.triu(oS)) < 1e-12
    assert np.linalg.norm(rhs[:n] - orr) / np.linalg.norm(orr) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("which,wid", [(4, 0), (3, 1)])
def test_imu_gnss_factor_through_ceres_api(which, wid):
    """IMUGNSSFactor objects (same public members as RVI/factor/gnss_imu_factor.h) added through
    ceres::Problem::AddResidualBlock: the shim's adapter turns them into chain records, the solve
    equals the direct C-ABI solve bit for bit, and the hidden GNSS frames are written back into
    the application arrays the factor points at."""
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    o = ob.OracleSolver(w.graph_p, opt)
    rc, state, cost, steps, hs, mat, rhs, msg = run_demo(which, wid, 0, w.n_state, o.n_f)
    assert rc >= 0, msg
    b = swgn.Batch([w.graph_p], opt)
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    hf = b.chain_frames(0)
    g = w.graph
    off = w.block_offsets()
    L = swgn.synth_lib()
    n = g.chain_frame_begin[g.n_chain]
    pb, sb = np.zeros(n, np.int32), np.zeros(n, np.int32)
    L.swgn_synth_chain_frame_blocks.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.swgn_synth_chain_frame_blocks(w.h, pb.ctypes.data_as(C.POINTER(C.c_int32)), sb.ctypes.data_as(C.POINTER(C.c_int32)))
    hidden = np.zeros(w.n_state, bool)
    for i in range(n):
        hidden[off[pb[i]]:off[pb[i]] + 7] = True
        hidden[off[sb[i]]:off[sb[i]] + 9] = True
        assert np.array_equal(state[off[pb[i]]:off[pb[i]] + 7], hf[i, :7])
        assert np.array_equal(state[off[sb[i]]:off[sb[i]] + 9], hf[i, 7:])
    assert not np.array_equal(hf, w.chain_frames0())
    assert np.array_equal(state[~hidden], x[~hidden])
    assert cost[1] == sm.final_cost and steps == [sm.num_successful_steps, sm.num_unsuccessful_steps]
    b.close()


def refdemo():
    """oracle/_ref/libswgn_refdemo.so: shim/ceres_shim_refdemo.cpp + the reference's own factor sources (oracle/build_ref.sh)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libswgn_refdemo.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.swgn_ceres_refdemo_solve.restype = C.c_int
    L.swgn_ceres_refdemo_solve.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(f64), C.POINTER(f64),
                                           C.POINTER(C.c_int), C.c_char_p, C.c_int]
    return L


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libswgn_refdemo.so")), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,wid,variant,strategy", [(1, 0, 0, 0), (2, 0, 0, 0), (2, 2, 7, 0), (2, 1, 0, 1), (1, 3, 4, 1)])
def test_reference_factor_classes_drop_in_through_the_shim(which, wid, variant, strategy):
    """A window built from the REFERENCE'S OWN projection_factor / IMUFactor + IntegrationBase / RTK*, Spp*, FixedInteger
    factor classes and PoseLocalParameterization (compiled unmodified from /root/reference), registered with
    shim/reference_adapters.h and solved by the shim's ceres::Solve on the device, with the reference's DOGLEG settings
    (strategy 0) and with Ceres' defaults LEVENBERG_MARQUARDT + jacobi_scaling (strategy 1).  The device's costs must be
    the costs the reference's own Evaluate() methods give on the CPU at the same states, and the state must be the one
    the C ABI returns for the same graph."""
    w = swgn.SynthWindow(which, wid, variant=variant)
    state = np.zeros(w.n_state)
    cost = np.zeros(4)
    steps = (C.c_int * 2)()
    msg = C.create_string_buffer(512)
    rc = refdemo().swgn_ceres_refdemo_solve(which, wid, variant, strategy, 0, 0, state.ctypes.data_as(C.POINTER(f64)),
                                            cost.ctypes.data_as(C.POINTER(f64)), steps, msg, 512)
    assert rc in (0, 1), msg.value.decode()
    dev_initial, dev_final, cpu_initial, cpu_final = cost
    assert abs(dev_initial - cpu_initial) <= 1e-11 * cpu_initial, (dev_initial, cpu_initial)
    assert abs(dev_final - cpu_final) <= 1e-9 * cpu_final, (dev_final, cpu_final)
    assert dev_final < 1e-3 * dev_initial
    opt = w.options()
    if strategy == 1:
        opt.trust_region_strategy = 1
        opt.jacobi_scaling = 1
    b = swgn.Batch([w.graph_p], opt)
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    b.close()
    assert (steps[0], steps[1]) == (sm.num_successful_steps, sm.num_unsuccessful_steps)
    assert float(np.max(np.abs(x - state) / np.maximum(1.0, np.abs(state)))) < 1e-9


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libswgn_refdemo.so")), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,wid,strategy", [(1, 1, 0), (2, 4, 0), (1, 2, 1)])
def test_reference_initialisation_factors_are_evaluated_on_the_host(which, wid, strategy):
    """Cost functions without a device adapter -- the reference's own InitialPoseFactor, InitialBiasFactor (initial_factor.cpp)
    and InitPose0Factor (pose0_factor.cpp), compiled unmodified -- go through the generic CostFunction contract: the shim
    calls their Evaluate() on the host at every evaluation point.  The device's costs must again be the reference's own."""
    w = swgn.SynthWindow(which, wid)
    state = np.zeros(w.n_state)
    cost = np.zeros(4)
    steps = (C.c_int * 2)()
    msg = C.create_string_buffer(512)
    rc = refdemo().swgn_ceres_refdemo_solve(which, wid, 0, strategy, 1, 0, state.ctypes.data_as(C.POINTER(f64)),
                                            cost.ctypes.data_as(C.POINTER(f64)), steps, msg, 512)
    assert rc in (0, 1), msg.value.decode()
    dev_initial, dev_final, cpu_initial, cpu_final = cost
    assert abs(dev_initial - cpu_initial) <= 1e-11 * cpu_initial, (dev_initial, cpu_initial)
    assert abs(dev_final - cpu_final) <= 1e-9 * cpu_final, (dev_final, cpu_final)
    assert dev_final < 1e-3 * dev_initial
    # the extra factors pull the window towards the truth they are anchored at: the result differs from the plain window's
    plain = np.zeros(w.n_state)
    refdemo().swgn_ceres_refdemo_solve(which, wid, 0, strategy, 0, 0, plain.ctypes.data_as(C.POINTER(f64)), cost.ctypes.data_as(C.POINTER(f64)), steps, msg, 512)
    assert np.abs(plain - state).max() > 1e-9


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libswgn_refdemo.so")), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,wid", [(4, 0), (4, 2), (3, 0)])
def test_reference_imugnss_factor_drops_in_through_the_shim(which, wid):
    """Composition A built from the reference's OWN IMUGNSSBase / IMUGNSSFactor objects (gnss_imu_factor.cpp compiled
    unmodified): the shim's adapter reads the object's public members into a chain record, the factor is evaluated on the
    device, and after ceres::Solve the hidden GNSS-frame states are back in the arrays gnss_poses[i] / gnss_speed_bias[i]
    point at.  Checked against (a) the reference class itself on the CPU -- its cost at the initial state, and at the
    returned state after a fresh elimination there -- and (b) the C ABI solve of the same graph."""
    L = refdemo()
    L.swgn_ceres_refdemo_chain_frames.argtypes = [C.POINTER(f64), C.c_int]
    w = swgn.SynthWindow(which, wid)
    state = np.zeros(w.n_state)
    cost = np.zeros(4)
    steps = (C.c_int * 2)()
    msg = C.create_string_buffer(512)
    rc = L.swgn_ceres_refdemo_solve(which, wid, 0, 0, 0, 0, state.ctypes.data_as(C.POINTER(f64)), cost.ctypes.data_as(C.POINTER(f64)), steps, msg, 512)
    assert rc in (0, 1), msg.value.decode()
    dev_initial, dev_final, cpu_initial, cpu_final = cost
    assert abs(dev_initial - cpu_initial) <= 1e-9 * cpu_initial, (dev_initial, cpu_initial)
    # the device's final cost is the linearised cost of the last accepted candidate; the reference class re-eliminates at
    # the returned states: equal up to the second-order term of the last step
    assert abs(dev_final - cpu_final) <= 2e-3 * cpu_final, (dev_final, cpu_final)
    assert dev_final < 1e-3 * dev_initial
    n = L.swgn_ceres_refdemo_chain_frames(None, 0)
    assert n == w.graph.chain_frame_begin[w.graph.n_chain]
    frames = np.zeros((n, 16))
    L.swgn_ceres_refdemo_chain_frames(frames.ctypes.data_as(C.POINTER(f64)), n)
    b = swgn.Batch([w.graph_p], w.options())
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    hf = b.chain_frames(0)
    b.close()
    assert (steps[0], steps[1]) == (sm.num_successful_steps, sm.num_unsuccessful_steps)
    assert float(np.max(np.abs(x - state) / np.maximum(1.0, np.abs(state)))) < 1e-9
    assert float(np.max(np.abs(hf - frames) / np.maximum(1.0, np.abs(frames)))) < 1e-9
    assert np.abs(frames - w.chain_frames0()).max() > 1e-6


_REF_EST = os.path.join(ROOT, "oracle", "_ref", "libref_estimator.so")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(_REF_EST), reason="oracle/_ref not built")
@pytest.mark.parametrize("wid", [0, 3])
def test_reference_update_schur_reads_the_shim_exports(wid):
    """SWFOptimization::UpdateSchur and UpdateSchurHessianOnly (RVI/swf/swf_gnss.cpp:25-94, compiled unmodified) applied to
    ceres::internal::{lhs_out, rhs_out, lhs_out2, hs_row, parameter_head} as the shim leaves them after a ceres::Solve of a
    cfg2 window built from the reference's factor classes: what the reference computes from those arrays is what the C ABI
    read-backs return for the same graph (swgn_batch_get_head_marginal, swgn_batch_get_tail_information)."""
    L = C.CDLL(_REF_EST)
    L.ref_est_update_schur.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(f64), C.POINTER(f64)]
    w = swgn.SynthWindow(2, wid)
    n_amb = w.n_amb
    n = C.c_int32()
    A, bv = np.zeros(n_amb * n_amb), np.zeros(n_amb)
    # export mode -> UpdateSchur
    assert L.ref_est_update_schur(2, wid, 0, n_amb, C.byref(n), A.ctypes.data_as(C.POINTER(f64)), bv.ctypes.data_as(C.POINTER(f64))) == 0
    assert n.value == n_amb
    opt = w.options()
    opt.is_optimize = 0
    opt.max_num_iterations = 1
    b = swgn.Batch([w.graph_p], opt)
    b.solve()
    A2, b2 = b.head_marginal(0, n_amb)
    b.close()
    A = A.reshape(n_amb, n_amb)
    # both sides reduce the SAME exported S (326 leading rows, condition 1e15+) with an eigen pseudo-inverse that drops
    # eigenvalues below an ABSOLUTE 1e-8: which of the eigenvalues near that threshold survive depends on the eigen-solver's
    # rounding (here: the stand-in Jacobi solver behind the reference code vs the device's), so (A, b) agree to 1e-4, not to
    # rounding (the device against the oracle's restatement, same algorithm: 1e-7, test_update_schur_on_the_device)
    assert np.abs(A - A2).max() < 1e-3 * np.abs(A2).max()
    assert np.abs(bv - b2).max() < 1e-3 * max(1.0, np.abs(b2).max())
    # optimising solve -> UpdateSchurHessianOnly
    A3 = np.zeros(n_amb * n_amb)
    assert L.ref_est_update_schur(2, wid, 1, n_amb, C.byref(n), A3.ctypes.data_as(C.POINTER(f64)), bv.ctypes.data_as(C.POINTER(f64))) == 0
    b = swgn.Batch([w.graph_p], w.options())
    b.solve()
    A4 = b.tail_information(0, n_amb)
    b.close()
    A3 = A3.reshape(n_amb, n_amb)
    assert np.abs(A3 - A4).max() < 1e-9 * np.abs(A4).max()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(_REF_EST), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,wid", [(4, 0), (4, 3), (3, 1)])
def test_reference_my_ordering_executed_on_a_composition_a_window(which, wid):
    """SWFOptimization::MyOrdering (RVI/swf/swf_gnss.cpp:629-783, unmodified) executed on a composition-A window whose parameter
    blocks live in an estimator's own storage (para_pose, para_speed_bias, para_ex_Pose, f_manager.feature, the RTK ambiguity
    lists, blackvalue2), called where MyOptimization calls it, the solve then running with the ordering it produced.  That
    ordering is the one the graph generator restates (same elimination set, same sequence of the remaining blocks), and the
    solve returns the states of the C-ABI solve of the flat graph."""
    L = C.CDLL(_REF_EST)
    L.ref_est_my_ordering.argtypes = [C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(f64), C.POINTER(C.c_int)]
    w = swgn.SynthWindow(which, wid)
    g = w.graph
    nb = C.c_int32()
    groups = np.full(g.n_blocks, -7, np.int32)
    state = np.zeros(w.n_state)
    steps = (C.c_int * 2)()
    rc = L.ref_est_my_ordering(which, wid, g.n_blocks, C.byref(nb), groups.ctypes.data_as(C.POINTER(C.c_int32)), state.ctypes.data_as(C.POINTER(f64)), steps)
    assert rc in (0, 1) and nb.value == g.n_blocks
    mine = np.array([g.block_group[b] for b in range(g.n_blocks)])
    const = np.array([g.block_const[b] for b in range(g.n_blocks)])
    touched = np.zeros(g.n_blocks, bool)
    for arr, n in ((g.proj_blocks, 3 * g.n_proj), (g.imu_blocks, 4 * g.n_imu), (g.prior_blocks, g.prior_blk_begin[g.n_prior]),
                   (g.chain_blocks, g.chain_blk_begin[g.n_chain]), (g.unit_block, g.n_unit)):
        for k in range(n):
            touched[arr[k]] = True
    live = touched & (const == 0)
    assert (groups[live] >= 0).all() and (groups[~live] == -1).all()
    assert set(np.where(groups == 0)[0]) == set(np.where(live & (mine == 0))[0])           # the elimination set
    rest = np.where(live & (groups > 0))[0]
    assert list(rest[np.argsort(groups[rest], kind="stable")]) == list(rest[np.argsort(mine[rest], kind="stable")])   # the sequence after it
    b = swgn.Batch([w.graph_p], w.options())
    sm = b.solve()[0]
    x = b.get_state(0, w.n_state)
    b.close()
    assert (steps[0], steps[1]) == (sm.num_successful_steps, sm.num_unsuccessful_steps)
    assert float(np.max(np.abs(x - state) / np.maximum(1.0, np.abs(state)))) < 1e-9


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(_REF_EST), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,wid", [(4, 0), (4, 5), (3, 2)])
def test_reference_add_all_residual_builds_and_solves_the_window(which, wid):
    """SWFOptimization::AddAllResidual(NormalMode) (RVI/swf/swf_core.cpp:209-415, unmodified) executed on an estimator
    filled with a composition-A window: the reference's own loop adds the marginalisation prior, the IMU links, the
    IMUGNSSFactor of every gap and the projection factors to a ceres::Problem (the shim's), sets the solver options, calls
    MyOrdering and ceres::Solve.  The states it leaves in para_pose / para_speed_bias / ptsInWorld / the ambiguity lists,
    and in the hidden GNSS frames of the IMUGNSSBase objects, are those of the C-ABI solve of the flat graph."""
    L = C.CDLL(_REF_EST)
    L.ref_est_add_all_residual.argtypes = [C.c_int, C.c_uint64, C.c_int, C.POINTER(f64), C.POINTER(f64), C.POINTER(C.c_int32)]
    w = swgn.SynthWindow(which, wid)
    g = w.graph
    state = np.zeros(w.n_state)
    n_hidden = g.chain_frame_begin[g.n_chain]
    frames = np.zeros((n_hidden, 16))
    nf = C.c_int32()
    rc = L.ref_est_add_all_residual(which, wid, w.n_state, state.ctypes.data_as(C.POINTER(f64)), frames.ctypes.data_as(C.POINTER(f64)), C.byref(nf))
    assert rc == 0 and nf.value == n_hidden
    assert np.abs(state - w.state0()).max() > 1e-6            # the solve ran
    opt = w.options()
    opt.max_num_iterations = 8                                 # what NormalMode sets (swf_core.cpp:397-401)
    b = swgn.Batch([w.graph_p], opt)
    b.solve()
    x = b.get_state(0, w.n_state)
    hf = b.chain_frames(0)
    b.close()
    assert float(np.max(np.abs(x - state) / np.maximum(1.0, np.abs(state)))) < 1e-9
    assert float(np.max(np.abs(hf - frames) / np.maximum(1.0, np.abs(frames)))) < 1e-9


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(_REF_EST), reason="oracle/_ref not built")
@pytest.mark.parametrize("which,wid", [(4, 0), (4, 1), (3, 0)])
def test_reference_marginalisation_of_the_oldest_frame_through_the_shim(which, wid):
    """SWFOptimization::AddAllResidual(MargeIncludeMode2) (RVI/swf/swf_core.cpp:209-468, unmodified), the marginalisation of the
    oldest image frame as MargFrames runs it (RVI/swf/swf.cpp:343-364), executed on an estimator filled with a composition-A
    window: the reference's code collects the factors that touch the frame's pose, speed-bias and the landmarks whose tracks start
    there, appends the blocks to keep to parameter_head, calls MyOrdering, ceres::Solve in export mode (the shim, on the device),
    UpdateSchur and MarginalizationInfo::setmarginalizeinfo.  The information (J0'J0, J0'r0) of the prior it ends with equals the
    one the C ABI returns for the same factors, the same drop set and the extrinsic free as in that mode
    (AddParameter2Problem(problem, true), swf_core.cpp:365)."""
    L = C.CDLL(_REF_EST)
    i32, P = C.c_int32, C.POINTER
    L.ref_est_marginalize_oldest.argtypes = [C.c_int, C.c_uint64, C.c_int, P(C.c_uint8), P(i32), P(i32), P(i32), P(i32), P(f64), P(f64)]
    w = swgn.SynthWindow(which, wid)
    g0 = w.graph
    nb, cap = g0.n_blocks, 1024
    drop = np.zeros(nb, np.uint8)
    n, nk = i32(), i32()
    kb, ki = np.zeros(nb, np.int32), np.zeros(nb, np.int32)
    J, r = np.zeros(cap * cap), np.zeros(cap)
    rc = L.ref_est_marginalize_oldest(which, wid, cap, drop.ctypes.data_as(P(C.c_uint8)), C.byref(n), C.byref(nk), kb.ctypes.data_as(P(i32)),
                                      ki.ctypes.data_as(P(i32)), J.ctypes.data_as(P(f64)), r.ctypes.data_as(P(f64)))
    assert rc == 0
    n, nk = n.value, nk.value
    kb, ki = kb[:nk], ki[:nk]
    Jr, rr = J[:n * n].reshape(n, n), r[:n]
    assert drop.sum() > 2 and nk > 3
    # ---- the same marginalisation through the C ABI
    g = swgn.Graph()
    C.memmove(C.byref(g), w.graph_p, C.sizeof(swgn.Graph))
    sizes = np.array([g.block_size[b] for b in range(nb)])
    F = int((sizes == 9).sum())
    touching = lambda blocks: any(drop[b] for b in blocks)
    pj = [i for i in range(g.n_proj) if touching(g.proj_blocks[3 * i:3 * i + 3])]
    im = [i for i in range(g.n_imu) if touching(g.imu_blocks[4 * i:4 * i + 4])]
    assert g.n_gnss == 0 and g.n_host == 0 and 0 < len(pj) < g.n_proj and len(im) == 1
    # only the factors that touch a dropped block (and the prior, the unit factor) enter the problem; the chains touch none
    proj_blocks = np.array([g.proj_blocks[3 * i + k] for i in pj for k in range(3)], np.int32)
    proj_uv = np.array([g.proj_uv[2 * i + k] for i in pj for k in range(2)])
    imu_blocks = np.array([g.imu_blocks[4 * i + k] for i in im for k in range(4)], np.int32)
    imu_data = np.concatenate([np.ctypeslib.as_array(g.imu_data, ((g.n_imu * 474),))[474 * i:474 * (i + 1)] for i in im]).copy()
    g.n_proj, g.proj_blocks, g.proj_uv = len(pj), proj_blocks.ctypes.data_as(P(i32)), proj_uv.ctypes.data_as(P(f64))
    g.n_imu, g.imu_blocks, g.imu_data = len(im), imu_blocks.ctypes.data_as(P(i32)), imu_data.ctypes.data_as(P(f64))
    g.n_chain = 0
    touched = np.zeros(nb, bool)
    touched[proj_blocks] = True
    touched[imu_blocks] = True
    touched[[g.prior_blocks[k] for k in range(g.prior_blk_begin[g.n_prior])]] = True
    konst = np.array([g.block_const[b] for b in range(nb)], np.int32)
    konst[2 * F] = 0                                   # the extrinsic is left variable in this mode
    keep = [b for b in range(nb) if touched[b] and not drop[b] and not konst[b] and b != nb - 1]
    assert sorted(keep) == sorted(kb.tolist())         # the reference kept the same blocks
    group = np.zeros(nb, np.int32)
    group[0], group[F] = 2, 1                          # pose and speed-bias of the frame; its landmarks stay in group 0
    for k, b in enumerate(keep):
        group[b] = 3 + k
    g.block_const = konst.ctypes.data_as(P(i32))
    g.block_group = group.ctypes.data_as(P(i32))
    g.is_use = None
    g.n_order, g.order = 0, None
    opt = w.options()
    opt.is_optimize, opt.max_num_iterations, opt.n_parameter_head = 0, 1, len(keep)
    b = swgn.Batch([C.pointer(g)], opt)
    b.solve()
    cb, co, cs = b.columns(0)
    assert list(cb[len(cb) - len(keep):]) == keep
    J0, r0, A, bv = b.marginal_prior(0, n)
    b.close()
    # swgn_marginalize on the same factors: constant and untouched blocks take no part, the keep blocks come in block order
    # (blackvalue2, which MyOrdering eliminates in group 0, is dropped explicitly)
    drop2 = drop.copy()
    drop2[nb - 1] = 1
    (mkb, mki, mJ, mr, mm), = swgn.marginalize([C.pointer(g)], [drop2])
    assert list(mkb) == keep and mJ.shape == (n, n)
    assert np.abs(mJ.T @ mJ - J0.T @ J0).max() < 1e-9 * np.abs(J0.T @ J0).max()
    assert np.abs(mJ.T @ mr - J0.T @ r0).max() < 1e-9 * max(1.0, np.abs(J0.T @ r0).max())
    # ---- same information, the reference's block order mapped onto the device's
    tang = {b_: (6 if sizes[b_] == 7 else int(sizes[b_])) for b_ in keep}
    dev_off, o = {}, 0
    for b_ in keep:
        dev_off[b_] = o
        o += tang[b_]
    assert o == n
    perm = np.concatenate([np.arange(dev_off[b_], dev_off[b_] + tang[b_]) for b_ in kb[np.argsort(ki)]])
    Ar, br = Jr.T @ Jr, Jr.T @ rr
    Ad, bd = (J0.T @ J0)[np.ix_(perm, perm)], (J0.T @ r0)[perm]
    assert np.abs(Ar - Ad).max() < 1e-9 * np.abs(Ad).max(), np.abs(Ar - Ad).max() / np.abs(Ad).max()   # measured 3e-12 .. 6e-12
    assert np.abs(br - bd).max() < 1e-9 * max(1.0, np.abs(bd).max()), np.abs(br - bd).max()                 # measured 1e-11 .. 7e-11
