"""The C-ABI library loads, exports every symbol include/swgn.h declares, fails loudly without a
device, and its host-side planner agrees with the oracle's preprocessing (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_binding as ob
import swgn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for header in ("swgn.h", "swgn_gnss.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(swgn_[a-z_0-9]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(os.path.join(ROOT, "rtk-visual-inertial-navigation_b200", "libswgn.so"))
    names = declared_symbols()
    assert len(names) >= 35 and "swgn_gnss_preprocess" in names
    for n in names:
        assert hasattr(L, n), n
    assert swgn.lib().swgn_version().decode().startswith("swgn")


def test_release_cached_memory_is_callable_without_a_device():
    """Nothing is cached before the first batch was destroyed: 0 bytes, no CUDA call, no error."""
    if swgn.lib().swgn_device_count() == 0:
        assert swgn.release_cached_memory() == 0


def test_default_options_are_the_reference_settings():
    o = swgn.default_options()
    assert o.max_num_iterations == 8 and o.max_num_consecutive_invalid_steps == 5
    assert o.initial_trust_region_radius == 1e4 and o.dogleg_min_mu == 1e-12
    assert o.function_tolerance == 1e-6 and o.gradient_tolerance == 1e-10 and o.parameter_tolerance == 1e-8
    assert o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32 and o.is_optimize == 1


@pytest.mark.parametrize("which,wid", [(1, 0), (2, 0), (2, 7)])
def test_planner_matches_oracle_preprocessing(which, wid):
    w = swgn.SynthWindow(which, wid)
    st, d = swgn.plan_probe(w.graph_p, w.options().n_parameter_head)
    assert st == 0
    o = ob.OracleSolver(w.graph_p, w.options())
    assert d["n_cols"] == o.n_col_blocks and d["n_ecols"] == o.n_e_blocks
    assert d["n_e"] == o.n_e and d["n_f"] == o.n_f and d["n_t"] == o.n_cols
    assert d["n_res"] == o.n_res and d["n_rows"] == o.n_row_blocks and d["n_chunks"] == o.n_e_blocks
    # same column and row order as the oracle's preprocessing (CPU view of what the gpu tests read back)
    st, pcols, prows = swgn.plan_order(w.graph_p, w.options().n_parameter_head)
    ocols, _, _ = o.columns()
    orows, _ = o.rows()
    assert st == 0 and np.array_equal(pcols, ocols) and np.array_equal(prows, orows)
    # SURVEY.md 8d byte formula, recomputed independently from the graph
    g = w.graph
    kinds = [g.gnss_kind[i] for i in range(g.n_gnss)]
    n_cp, n_pr, n_dop = kinds.count(2), kinds.count(3), kinds.count(4)
    npri = g.prior_n[0]
    n_black = g.n_unit  # 1 x (1 + 1) each
    doubles = (g.n_proj * (2 * 3 + 2 * 6 + 2) + g.n_imu * 15 * (6 + 9 + 6 + 9 + 1) + n_cp * (6 + 1 + 1 + 1) +
               n_pr * (6 + 1 + 1) + n_dop * (9 + 1 + 6 + 1) + npri * (npri + 1) + 2 * n_black + o.n_cols +
               o.n_f * (o.n_f + 1) // 2 + o.n_f + o.n_e)
    assert d["schur_bytes"] == 8 * doubles


def test_planner_rejects_dependent_first_group_and_missing_ordering():
    w = swgn.SynthWindow(1, 0)
    g = w.graph
    groups = np.ctypeslib.as_array(g.block_group, shape=(g.n_blocks,))
    saved = groups.copy()
    try:
        groups[0] = 0
        st, _ = swgn.plan_probe(w.graph_p)
        assert st == 4  # SWGN_ERR_ORDERING
        assert b"independent" in swgn.lib().swgn_last_error()
        groups[:] = saved
        groups[1] = -1
        st, _ = swgn.plan_probe(w.graph_p)
        assert st == 4
    finally:
        groups[:] = saved


@pytest.mark.skipif(swgn.lib().swgn_device_count() > 0, reason="a CUDA device is present")
def test_create_fails_loudly_without_a_device():
    w = swgn.SynthWindow(1, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        swgn.Batch([w.graph_p], w.options())


@pytest.mark.skipif(swgn.lib().swgn_device_count() > 0, reason="a CUDA device is present")
def test_epoch_and_marginalisation_entry_points_fail_loudly_without_a_device():
    """The host-side phases of include/swgn_gnss.h run anywhere; everything numerical needs the device and says so."""
    import gnss_scenario as S
    import swgn_gnss as G
    cfg = G.default_config()
    sc = S.Scenario(0, cfg=cfg)
    e, obs, f = sc.epoch(0)
    T = G.Tracker(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.preprocess([T], [e], [f])
    assert T.count(G.AMB_RTK) == 0          # nothing was decided without the device residuals
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.gate_residuals(np.zeros((2, 16)))
    w = swgn.SynthWindow(1, 0)
    g = swgn.Graph()
    C.memmove(C.byref(g), w.graph_p, C.sizeof(swgn.Graph))
    free = np.zeros(g.n_blocks, np.int32)
    g.block_const = free.ctypes.data_as(C.POINTER(C.c_int32))
    g.n_order, g.order, g.is_use = 0, None, None
    drop = np.zeros(g.n_blocks, np.uint8)
    drop[0] = 1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        swgn.marginalize([C.pointer(g)], [drop])
    # argument errors are reported before any device is touched
    job = G.FixedIntegerArrays([1, 1], [0, 1], [0.0, 0.0], np.eye(2), [0.0, 0.0], [0.0, 0.0], [], [], [])
    with pytest.raises(RuntimeError, match="no fixed double difference"):
        G.fixed_integer_prior([job])


@pytest.mark.parametrize("which,wid,kw", [(1, 0, {}), (2, 0, {}), (2, 5, {}), (3, 0, {}), (4, 1, {}),
                                           (2, 1, dict(n_keyframes=40, n_landmarks=100, n_gnss_epochs=20))])
def test_symbolic_cholesky_masks_cover_the_numeric_factor(which, wid, kw):
    """k_chol skips the tiles the planner's symbolic fill-in marks as zero (32-row panels x 16-column groups).
    Host-only check that the masks are conservative: every non-zero of the reduced system AND of its
    Cholesky factor U (computed densely from the oracle's S) lies in a live group of its panel."""
    w = swgn.SynthWindow(which, wid, **kw)
    opt = w.options()
    masks = swgn.plan_chol_masks(w.graph_p, opt.n_parameter_head)
    o = ob.OracleSolver(w.graph_p, opt)
    st, x, S, rhs = o.linear_solve(np.full(o.n_cols, 1e-3))
    assert st == 0
    n = o.n_f
    assert len(masks) == (n + 31) // 32
    Sf = np.triu(S) + np.triu(S, 1).T
    U = np.linalg.cholesky(Sf).T
    for name, M in (("S", np.triu(S)), ("U", U)):
        for k in range(len(masks)):
            rows = M[32 * k:min(n, 32 * k + 32)]
            cols = np.nonzero(np.abs(rows).sum(axis=0) > 0)[0]
            groups = set(int(c) // 16 for c in cols)
            live = set(g for g in range(64) if (int(masks[k]) >> g) & 1)
            assert groups <= live, (name, k, sorted(groups - live))
    # and the masks do skip something on the BASELINE window (otherwise the feature is dead weight)
    if which == 2 and not kw:
        assert sum(bin(int(m)).count("1") for m in masks) < sum((n + 1 + 15) // 16 - 2 * k for k in range(len(masks)))


@pytest.mark.parametrize("which,wid", [(1, 0), (2, 0), (2, 3), (3, 0), (4, 0)])
def test_gather_streams_are_well_formed(which, wid):
    """Host-only: the per-warp gather streams k_schur consumes -- ranges of every operand and store, one end
    flag per tile, the live terms are exactly the MMAs the planner counted, the warps are balanced."""
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    L = swgn.lib()
    L.swgn_plan_stream_check.argtypes = [C.POINTER(swgn.Graph), C.c_int32, C.POINTER(C.c_int64)]
    out = (C.c_int64 * 8)()
    assert L.swgn_plan_stream_check(w.graph_p, opt.n_parameter_head, out) == 0, L.swgn_last_error()
    stages, live, pad, tiles, smin, smax, e_stages, e_live = list(out)
    st, d = swgn.plan_probe(w.graph_p, opt.n_parameter_head)
    assert live + e_live == d["n_mma"]
    assert stages * 8 == live + pad and tiles >= d["n_scells"]
    assert smax - smin <= max(4, 0.05 * smax)      # longest-first dealing keeps the 8 warps level
    assert pad < 0.45 * live                        # padding to full stages stays a minority


def test_planner_on_the_edges_of_the_reduced_program():
    """RemoveFixedBlocks corner cases (CERES program_test.cc RemoveFixedBlocks*): nothing constant -> the whole
    program; every block constant -> empty reduced program, reported (the reference's preprocessor returns
    CONVERGENCE with the fixed cost there, trust_region_preprocessor.cc:380-384, solver.cc:418-432 -- a stated deviation of the shim);
    all f-blocks constant -> a valid plan with an empty reduced system; all e-blocks constant -> the first
    elimination group is empty, which Ceres answers by switching the linear solver: reported as unsupported."""
    import json
    from linear_graph import LinearGraph
    with open(os.path.join(ROOT, "tests", "golden", "ceres_llsq_problems.json")) as f:
        p = json.load(f)["problem2"]
    ne = p["num_eliminate_blocks"]
    st, info = swgn.plan_probe(LinearGraph(p).graph_p, 0)
    assert st == 0 and info["n_cols"] == len(p["col_sizes"]) and info["n_rows"] == len(p["row_sizes"])
    lg = LinearGraph(p)
    lg.block_const[:] = 1
    st, info = swgn.plan_probe(lg.graph_p, 0)
    assert st == 1 and b"empty reduced program" in swgn.lib().swgn_last_error()
    lg = LinearGraph(p)
    lg.block_const[ne:] = 1
    st, info = swgn.plan_probe(lg.graph_p, 0)
    assert st == 0 and info["n_f"] == 0 and info["n_e"] == sum(p["col_sizes"][:ne]) and info["n_scells"] == 0
    lg = LinearGraph(p)
    lg.block_const[:ne] = 1
    st, info = swgn.plan_probe(lg.graph_p, 0)
    assert st != 0 and b"first elimination group is empty" in swgn.lib().swgn_last_error()
