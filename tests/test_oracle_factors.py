"""The reference holds no test for its factors (parity unpinned by the reference);
the oracle's restatements are pinned here by analytic-vs-numeric Jacobian checks on the manifold
(x [+] delta with PoseLocalParameterization::Plus) and by closed-form properties."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import swgn

P = C.POINTER(C.c_double)


def dp(a):
    return a.ctypes.data_as(P)


def factor_eval(kind, kind2, globals_, record, params, sizes, nres, jac=True):
    r = np.zeros(nres)
    J = np.zeros(nres * sum(sizes)) if jac else None
    st = ob.oracle().oracle_factor_eval(kind, kind2, dp(globals_), dp(record), dp(params), dp(r), dp(J) if jac else None)
    assert st == 0
    Js = []
    if jac:
        o = 0
        for s in sizes:
            Js.append(J[o:o + nres * s].reshape(nres, s).copy())
            o += nres * s
    return r, Js


def plus(x, sizes, manif, delta):
    out, o, d = [], 0, 0
    for s, m in zip(sizes, manif):
        if m:
            y = np.zeros(7)
            xb = np.ascontiguousarray(x[o:o + 7])
            db = np.ascontiguousarray(delta[d:d + 6])
            ob.oracle().oracle_pose_plus(dp(xb), dp(db), dp(y))
            out.append(y)
            d += 6
        else:
            out.append(x[o:o + s] + delta[d:d + s])
            d += s
        o += s
    return np.concatenate(out)


def check_numeric(kind, kind2, globals_, record, params, sizes, manif, nres, h, rtol):
    r0, Js = factor_eval(kind, kind2, globals_, record, params, sizes, nres)
    local = [6 if m else s for s, m in zip(sizes, manif)]
    # analytic Jacobian in tangent space: global J times [I6; 0] for poses (7th column must be 0)
    Jl = []
    for J, m in zip(Js, manif):
        if m:
            assert np.all(J[:, 6] == 0.0)
            Jl.append(J[:, :6])
        else:
            Jl.append(J)
    Ja = np.hstack(Jl)
    n = sum(local)
    Jn = np.zeros((nres, n))
    for k in range(n):
        d = np.zeros(n)
        d[k] = h
        rp, _ = factor_eval(kind, kind2, globals_, record, plus(params, sizes, manif, d), sizes, nres, jac=False)
        rm, _ = factor_eval(kind, kind2, globals_, record, plus(params, sizes, manif, -d), sizes, nres, jac=False)
        Jn[:, k] = (rp - rm) / (2 * h)
    scale = np.abs(Ja).max()
    assert np.abs(Ja - Jn).max() <= rtol * scale, (np.abs(Ja - Jn).max(), scale)


def rand_pose(rng, pscale=1.0):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    if q[3] < 0:
        q = -q
    return np.concatenate([rng.normal(size=3) * pscale, q])


GLOBALS = np.array([-0.005, 0.009, 0.31, 0.3, -5.0, 8.4, 1000 / 1.5, 0, 0, 1000 / 1.5])


def test_projection_factor_jacobians():
    rng = np.random.default_rng(1)
    for _ in range(10):
        pose = rand_pose(rng, 5.0)
        ext = rand_pose(rng, 0.05)
        # a landmark 5-20 m in front of the camera: build it from the camera frame
        w = swgn.SynthWindow(1, int(rng.integers(0, 1000)))
        g = w.graph
        x0 = w.truth()
        offs = w.block_offsets()
        i = int(rng.integers(0, g.n_proj))
        blocks = [g.proj_blocks[3 * i + k] for k in range(3)]
        params = np.concatenate([x0[offs[b]:offs[b] + s] for b, s in zip(blocks, (7, 7, 3))])
        params = params + np.concatenate([np.zeros(7), np.zeros(7), rng.normal(size=3) * 0.05])
        glob = np.array(list(g.Pbg) + list(g.gravity) + list(g.proj_sqrt_info))
        uv = np.array([g.proj_uv[2 * i], g.proj_uv[2 * i + 1]])
        check_numeric(0, 0, glob, uv, params, (7, 7, 3), (1, 1, 0), 2, 1e-6, 2e-6)


def test_imu_factor_jacobians():
    rng = np.random.default_rng(2)
    for wid in range(3):
        w = swgn.SynthWindow(1, wid)
        g = w.graph
        x0 = w.state0()
        offs = w.block_offsets()
        glob = np.array(list(g.Pbg) + list(g.gravity) + list(g.proj_sqrt_info))
        for i in range(g.n_imu):
            rec = np.ctypeslib.as_array(g.imu_data, shape=(g.n_imu * 474,))[474 * i:474 * (i + 1)].copy()
            blocks = [g.imu_blocks[4 * i + k] for k in range(4)]
            params = np.concatenate([x0[offs[b]:offs[b] + s] for b, s in zip(blocks, (7, 9, 7, 9))])
            check_numeric(1, 0, glob, rec, params, (7, 9, 7, 9), (1, 0, 1, 0), 15, 1e-6, 5e-6)


@pytest.mark.parametrize("kind,sizes,manif", [
    (0, (7, 1), (1, 0)), (1, (7, 1, 1), (1, 0, 0)), (2, (7, 1, 1), (1, 0, 0)), (3, (7, 1), (1, 0)),
    (4, (9, 1, 7), (0, 0, 1)), (5, (1, 1), (0, 0))])
def test_gnss_factor_jacobians(kind, sizes, manif):
    rng = np.random.default_rng(10 + kind)
    for _ in range(5):
        rec = np.zeros(16)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        base = np.array([-2323932.39454, 5387298.51324, 2493096.51920])
        rec[0:3] = base + d * 2.2e7
        rec[3:6] = rng.normal(size=3) * 2e3
        rec[6:9] = base
        rec[9] = rng.normal() * 10
        rec[10] = 0.19
        rec[11] = 1.0 / 0.3
        parts = []
        for s, m in zip(sizes, manif):
            if m:
                parts.append(rand_pose(rng, 10.0))
            elif s == 9:
                parts.append(rng.normal(size=9))
            else:
                parts.append(rng.normal(size=s) * 5)
        params = np.concatenate(parts)
        # the range Jacobian ignores the Sagnac term (gnss_factor.cpp:122-127 vs common_function.cpp:133):
        # a relative 1e-5 model error is part of the reference
        check_numeric(2, kind, GLOBALS, rec, params, sizes, manif, 1, 1e-3 if kind != 5 else 1e-6, 2e-4)


def test_varerr2_uses_single_precision_sine():
    el, dt, var = 0.7, 0.1, 9e-6
    v = ob.oracle().oracle_varerr2(el, dt, var)
    s = np.float32(np.sin(np.float32(el)))
    b = 299792458.0 * 5e-12 * dt
    expect = var / float(s) / float(s) + b * b
    assert abs(v - expect) <= 1e-15 * expect
    exact = var / np.sin(el) ** 2 + b * b
    assert v != exact  # the float rounding is visible


def test_cauchy_loss_values_and_derivatives():
    """Includes the points of the reference's loss_function_test.cc:118-126 (CauchyLoss(0.7), CauchyLoss(1.3) at
    s = 0.357, 1.792 and at 0): value against the closed form, derivatives against central differences."""
    rho = np.zeros(3)
    ob.oracle().oracle_cauchy(0.7, 0.0, dp(rho))
    assert rho[0] == 0.0 and rho[1] == 1.0 and abs(rho[2] + 1.0 / 0.49) < 1e-15
    for a in (1.0, 2.5, 0.7, 1.3):
        for s in (0.0, 0.3, 7.0, 1e4, 0.357, 1.792):
            ob.oracle().oracle_cauchy(a, s, dp(rho))
            assert abs(rho[0] - a * a * np.log1p(s / (a * a))) <= 1e-12 * max(1, rho[0])
            h = 1e-6 * max(1.0, s)
            r1, r2 = np.zeros(3), np.zeros(3)
            ob.oracle().oracle_cauchy(a, s + h, dp(r1))
            ob.oracle().oracle_cauchy(a, max(s - h, 0), dp(r2))
            if s > 0:
                assert abs((r1[0] - r2[0]) / (2 * h) - rho[1]) < 1e-6
                assert abs((r1[1] - r2[1]) / (2 * h) - rho[2]) < 1e-6
            assert rho[2] < 0  # corrector always takes the alpha = 0 branch


def test_pose_plus_is_a_retraction():
    rng = np.random.default_rng(4)
    x = rand_pose(rng)
    out = np.zeros(7)
    ob.oracle().oracle_pose_plus(dp(x), dp(np.zeros(6)), dp(out))
    np.testing.assert_allclose(out, x, atol=1e-15)
    d = rng.normal(size=6) * 0.1
    ob.oracle().oracle_pose_plus(dp(x), dp(d), dp(out))
    assert abs(np.linalg.norm(out[3:]) - 1) < 1e-15
    np.testing.assert_allclose(out[:3], x[:3] + d[:3])


def test_corrector_scales_residuals_and_jacobians_like_the_reference_tests():
    """CERES/internal/ceres/corrector_test.cc ScalarCorrection / ScalarCorrectionAlphaClamped (:57-139): with rho'' < 0
    the corrector takes alpha = 0, so residual' = sqrt(rho') residual and J' = sqrt(rho') J, and the cost is rho / 2.
    Checked on every robustified residual block of a cfg1 window: the same graph evaluated without the loss gives
    the raw (r, J); Cauchy(1) has rho'' < 0 everywhere (loss_function.cc:73-80)."""
    import swgn
    w = swgn.SynthWindow(1, 2)
    a = w.graph.proj_cauchy_a
    assert a > 0
    rob = ob.OracleSolver(w.graph_p, w.options())
    cost_rob, r_rob, _, J_rob = rob.evaluate()
    w.graph.proj_cauchy_a = 0.0
    try:
        raw = ob.OracleSolver(w.graph_p, w.options())
        cost_raw, r_raw, _, J_raw = raw.evaluate()
    finally:
        w.graph.proj_cauchy_a = a
    factor, offset = rob.rows()
    n_checked, cost = 0, 0.0
    bounds = list(offset) + [rob.n_res]
    rho = np.zeros(3)
    for k in range(rob.n_row_blocks):
        r0, r1 = bounds[k], bounds[k + 1]
        s = float(r_raw[r0:r1] @ r_raw[r0:r1])
        if r1 - r0 == 2 and not np.array_equal(r_rob[r0:r1], r_raw[r0:r1]):  # a projection factor under the loss
            ob.oracle().oracle_cauchy(a, s, dp(rho))
            assert rho[2] < 0
            np.testing.assert_allclose(r_rob[r0:r1], np.sqrt(rho[1]) * r_raw[r0:r1], rtol=1e-15, atol=0)
            np.testing.assert_allclose(J_rob[r0:r1], np.sqrt(rho[1]) * J_raw[r0:r1], rtol=1e-15, atol=1e-300)
            cost += 0.5 * rho[0]
            n_checked += 1
        else:
            np.testing.assert_array_equal(r_rob[r0:r1], r_raw[r0:r1])
            np.testing.assert_array_equal(J_rob[r0:r1], J_raw[r0:r1])
            cost += 0.5 * s
    assert n_checked == w.graph.n_proj
    assert abs(cost - cost_rob) <= 1e-13 * cost_rob and cost_rob < cost_raw
