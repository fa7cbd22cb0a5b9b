"""Pins the oracle's factor restatements (oracle/oracle_factors.cpp, oracle_preint.cpp) on the reference's OWN factor
classes: RVI/factor/{gnss_factor, projection_factor, imu_factor, integration_base, pose_local_parameterization}.cpp
compiled unmodified from /root/reference into oracle/_ref/libref_gnss.so (oracle/build_ref.sh; Eigen is not installed,
so they are compiled against the minimal eager stand-in in oracle/ref_stubs/ and against this repository's
include/ceres/ headers).  The GNSS factors, which touch Eigen only in the Doppler pose Jacobian, are compared bit for
bit; the Eigen-heavy factors to rounding (the stand-in evaluates products in plain ascending order, a real Eigen build
may associate differently)."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import swgn
from test_oracle_factors import GLOBALS, rand_pose

pytestmark = pytest.mark.skipif(ob.ref() is None, reason="oracle/_ref/libref_gnss.so not built")
P = C.POINTER(C.c_double)


def dp(a):
    return a.ctypes.data_as(P)


def both(kind, kind2, globals_, record, params, sizes, nres):
    out = []
    for lib, fn in ((ob.oracle(), "oracle_factor_eval"), (ob.ref(), "ref_factor_eval")):
        f = getattr(lib, fn)
        f.argtypes = [C.c_int, C.c_int, P, P, P, P, P]
        r = np.zeros(nres)
        J = np.zeros(nres * sum(sizes))
        assert f(kind, kind2, dp(globals_), dp(record), dp(params), dp(r), dp(J)) == 0
        r2 = np.zeros(nres)
        assert f(kind, kind2, dp(globals_), dp(record), dp(params), dp(r2), None) == 0  # residual-only evaluation
        assert np.array_equal(r, r2)
        out.append((r, J))
    return out


def gnss_record(rng, weight_from_var):
    rec = np.zeros(16)
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    base = np.array([-2323932.39454, 5387298.51324, 2493096.51920])
    rec[0:3] = base + d * 2.2e7
    rec[3:6] = rng.normal(size=3) * 2e3
    rec[6:9] = base
    rec[9] = rng.normal() * 10
    rec[10] = 0.19
    rec[12], rec[13], rec[14] = rng.uniform(0.2, 1.5), rng.uniform(0, 1.0), rng.uniform(1e-6, 1e-1)
    if weight_from_var:  # the RTK factors weigh with 1 / sqrt(varerr2(el, dt, var)) themselves (gnss_factor.cpp:98-103)
        rec[11] = 1.0 / np.sqrt(ob.oracle().oracle_varerr2(rec[12], rec[13], rec[14]))
    else:
        rec[11] = 1.0 / rng.uniform(0.05, 3.0)
    return rec


@pytest.mark.parametrize("kind,sizes,manif", [
    (0, (7, 1), (1, 0)), (1, (7, 1, 1), (1, 0, 0)), (2, (7, 1, 1), (1, 0, 0)), (3, (7, 1), (1, 0)),
    (4, (9, 1, 7), (0, 0, 1)), (5, (1, 1), (0, 0))])
def test_gnss_factors_bit_exact_against_the_reference_classes(kind, sizes, manif):
    rng = np.random.default_rng(40 + kind)
    for _ in range(50):
        rec = gnss_record(rng, weight_from_var=kind in (2, 3))
        parts = []
        for s, m in zip(sizes, manif):
            parts.append(rand_pose(rng, 10.0) if m else (rng.normal(size=9) if s == 9 else rng.normal(size=s) * 5))
        params = np.concatenate(parts)
        (ro, Jo), (rr, Jr) = both(2, kind, GLOBALS, rec, params, sizes, 1)
        assert np.array_equal(ro, rr)
        if kind == 4:
            # the Doppler pose Jacobian is the one Eigen expression of the file: (sqrt_info ev') (I - e e') / r
            np.testing.assert_allclose(Jo, Jr, rtol=1e-13, atol=1e-18)
            assert np.array_equal(Jo[:10], Jr[:10])  # velocity and drift blocks: no Eigen involved
        else:
            assert np.array_equal(Jo, Jr)


def test_rtk_carrier_without_istd_has_unit_weight():
    """RTKCarrierPhaseFactor(use_istd = false) weighs with 1 (gnss_factor.cpp:116-117), as the per-epoch phase-bias
    initialisation builds it (swf_gnss.cpp:355,367); the record carries that as WEIGHT = 1, var = 0."""
    rng = np.random.default_rng(7)
    rec = gnss_record(rng, True)
    rec[14] = 0.0
    rec[11] = 1.0
    params = np.concatenate([rand_pose(rng, 10.0), rng.normal(size=1), rng.normal(size=1)])
    (ro, Jo), (rr, Jr) = both(2, 2, GLOBALS, rec, params, (7, 1, 1), 1)
    assert np.array_equal(ro, rr) and np.array_equal(Jo, Jr)
    assert Jr[8] == 1.0  # d r / d clock = sqrt_info = 1


def test_varerr2_bit_exact():
    ob.ref().ref_varerr2.restype = C.c_double
    ob.ref().ref_varerr2.argtypes = [C.c_double] * 3
    rng = np.random.default_rng(0)
    for _ in range(200):
        el, dt, var = rng.uniform(0.05, 1.55), rng.uniform(0, 2), rng.uniform(1e-8, 1.0)
        assert ob.oracle().oracle_varerr2(el, dt, var) == ob.ref().ref_varerr2(el, dt, var)


def test_projection_factor_against_the_reference_class():
    rng = np.random.default_rng(1)
    for trial in range(20):
        w = swgn.SynthWindow(1, int(rng.integers(0, 1000)))
        g = w.graph
        x0 = w.truth()
        offs = w.block_offsets()
        i = int(rng.integers(0, g.n_proj))
        blocks = [g.proj_blocks[3 * i + k] for k in range(3)]
        params = np.concatenate([x0[offs[b]:offs[b] + s] for b, s in zip(blocks, (7, 7, 3))])
        params[14:] += rng.normal(size=3) * 0.05
        params[7:14] = rand_pose(rng, 0.05) if trial % 2 else params[7:14]  # also a rotated, displaced camera
        glob = np.array(list(g.Pbg) + list(g.gravity) + list(g.proj_sqrt_info))
        uv = np.array([g.proj_uv[2 * i], g.proj_uv[2 * i + 1]])
        (ro, Jo), (rr, Jr) = both(0, 0, glob, uv, params, (7, 7, 3), 2)
        np.testing.assert_allclose(ro, rr, rtol=0, atol=1e-12 * max(1.0, np.abs(rr).max()))
        np.testing.assert_allclose(Jo, Jr, rtol=0, atol=1e-13 * np.abs(Jr).max())  # pose, EXTRINSIC and landmark blocks


def test_imu_factor_against_the_reference_class():
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
            (ro, Jo), (rr, Jr) = both(1, 0, glob, rec, params, (7, 9, 7, 9), 15)
            np.testing.assert_allclose(ro, rr, rtol=0, atol=1e-11 * max(1.0, np.abs(rr).max()))
            np.testing.assert_allclose(Jo, Jr, rtol=0, atol=1e-12 * np.abs(Jr).max())


def test_preintegration_against_the_reference_class():
    """IntegrationBase::push_back / propagate / midPointIntegration / get_sqrtinfo run by the reference's own code."""
    rng = np.random.default_rng(11)
    noise = np.array([0.08, 0.004, 0.00004, 2.0e-6])  # yaml acc_n, gyr_n, acc_w, gyr_w
    ref = ob.ref()
    ref.ref_preintegrate.argtypes = [C.c_int, P, P, P, P]
    for n in (11, 101, 201):
        s = np.zeros((n, 7))
        s[:, 0] = 0.005
        s[:, 1:4] = rng.normal(size=(n, 3)) * 0.5 + np.array([0.1, -0.2, 9.8])
        s[:, 4:7] = rng.normal(size=(n, 3)) * 0.05
        bias = rng.normal(size=6) * 0.01
        rec_o, bad = ob.preintegrate_batch(np.array([0, n]), s.ravel(), bias, noise)
        assert bad == 0
        rec_r = np.zeros(474)
        assert ref.ref_preintegrate(n, dp(s), dp(bias), dp(noise), dp(rec_r)) == 0
        ro = rec_o[0]
        np.testing.assert_allclose(ro[:24], rec_r[:24], rtol=0, atol=1e-13 * max(1.0, np.abs(rec_r[:24]).max()))  # deltas, biases, gyr, sum_dt
        np.testing.assert_allclose(ro[24:249], rec_r[24:249], rtol=0, atol=1e-12 * np.abs(rec_r[24:249]).max())  # Jacobian
        # sqrt_info = LLT(covariance^-1): cond(covariance) ~ 1e10 amplifies the rounding of the two inversion algorithms
        Io, Ir = ro[249:474].reshape(15, 15), rec_r[249:474].reshape(15, 15)
        np.testing.assert_allclose(Io.T @ Io, Ir.T @ Ir, rtol=0, atol=1e-5 * np.abs(Ir.T @ Ir).max())


def test_pose_plus_against_the_reference_class():
    rng = np.random.default_rng(4)
    ref = ob.ref()
    ref.ref_pose_plus.argtypes = [P, P, P]
    for _ in range(50):
        x = rand_pose(rng, 10.0)
        d = rng.normal(size=6) * 0.2
        a, b = np.zeros(7), np.zeros(7)
        ob.oracle().oracle_pose_plus(dp(x), dp(d), dp(a))
        ref.ref_pose_plus(dp(x), dp(d), dp(b))
        np.testing.assert_allclose(a, b, rtol=0, atol=4e-16 * max(1.0, np.abs(b).max()))


def test_marginalization_factor_against_the_reference_class():
    """a5: MarginalizationFactor::Evaluate (RVI/factor/marginalization_factor.cpp:410-446) of the reference, compiled into
    oracle/_ref, against the oracle's restatement: residual r0 + J0 (x [-] x0) with the quaternion difference and its sign fix,
    Jacobians = column slices of J0 padded to the global size."""
    R = ob.ref()
    if R is None or not hasattr(R, "ref_prior_eval"):
        pytest.skip("oracle/_ref not built")
    i32, f64, P = C.c_int32, C.c_double, C.POINTER
    sig = [C.c_int, C.c_int, P(i32), P(i32), P(f64), P(f64), P(f64), P(f64), P(f64), P(f64)]
    R.ref_prior_eval.argtypes = sig
    O = ob.oracle()
    O.oracle_prior_eval.argtypes = sig
    rng = np.random.default_rng(5)
    for trial in range(4):
        sizes = np.array([7, 9, 1, 7, 3, 1], np.int32)
        tang = [6 if s == 7 else int(s) for s in sizes]
        idx = np.concatenate([[0], np.cumsum(tang)[:-1]]).astype(np.int32)
        n = int(sum(tang))
        J0, r0 = rng.normal(size=(n, n)), rng.normal(size=n)

        def state(flip):
            x = rng.normal(size=int(sizes.sum()))
            o = 0
            for s in sizes:
                if s == 7:
                    x[o + 3:o + 7] /= np.linalg.norm(x[o + 3:o + 7])
                    if flip:
                        x[o + 3:o + 7] *= -1     # exercises the w < 0 branch of the quaternion difference (:425-428)
                o += s
            return x
        x0, x = state(False), state(trial % 2 == 1)
        nj = n * int(sizes.sum())
        out = []
        for L in (R.ref_prior_eval, O.oracle_prior_eval):
            res, jac = np.zeros(n), np.zeros(nj)
            assert L(len(sizes), n, sizes.ctypes.data_as(P(i32)), idx.ctypes.data_as(P(i32)), ob._dp(x0), ob._dp(np.ascontiguousarray(J0)),
                     ob._dp(r0), ob._dp(x), ob._dp(res), ob._dp(jac)) == 0
            out.append((res, jac))
        (rr, jr), (ro, jo) = out
        assert np.abs(rr - ro).max() < 1e-12 * max(1.0, np.abs(rr).max())
        assert np.array_equal(jr, jo)      # copies of J0's columns: exact
