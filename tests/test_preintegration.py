"""IMU pre-integration (RVI/factor/integration_base.cpp:5-142): the constants of
IMUFactor.  CPU: the oracle restatement against closed forms and against the generator's independent
implementation (the reference holds no test or fixture for IntegrationBase: parity unpinned by the
reference).  GPU: swgn_preintegrate_batch through the C ABI against the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob
import swgn

NOISE = np.array([0.05, 0.005, 0.0005, 0.00005])  # yaml acc_n, gyr_n, acc_w, gyr_w
O_DP, O_DQ, O_DV, O_DT, O_J, O_SQ = 0, 3, 7, 22, 24, 249


def streams(n, seed, lo=2, hi=120):
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, n)
    begin = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    s = np.zeros((begin[-1], 7))
    s[:, 0] = 0.0025
    for f in range(n):
        t = np.arange(lens[f]) * 0.0025
        ph = rng.uniform(0, 6.28, 6)
        s[begin[f]:begin[f + 1], 1:4] = np.array([0.5, -0.3, 9.8]) + 1.5 * np.sin(2.0 * t[:, None] + ph[:3]) + 0.05 * rng.normal(size=(lens[f], 3))
        s[begin[f]:begin[f + 1], 4:7] = 0.4 * np.sin(1.3 * t[:, None] + ph[3:]) + 0.005 * rng.normal(size=(lens[f], 3))
    bias = np.hstack([0.05 * rng.normal(size=(n, 3)), 0.005 * rng.normal(size=(n, 3))])
    return begin, s, bias


def test_oracle_constant_acceleration_closed_form():
    n = 101
    s = np.zeros((n, 7))
    s[:, 0] = 0.0025
    s[:, 1:4] = [0.3, -0.2, 9.8]
    bias = np.array([[0.01, 0.02, -0.03, 0, 0, 0]])
    rec, bad = ob.preintegrate_batch([0, n], s, bias, NOISE)
    assert bad == 0
    T = 0.25
    a = np.array([0.3, -0.2, 9.8]) - bias[0, :3]
    r = rec[0]
    assert abs(r[O_DT] - T) < 1e-12
    assert np.allclose(r[O_DV:O_DV + 3], a * T, rtol=1e-12)
    assert np.allclose(r[O_DP:O_DP + 3], 0.5 * a * T * T, rtol=1e-12)
    assert np.allclose(r[O_DQ:O_DQ + 4], [0, 0, 0, 1], atol=1e-15)
    J = r[O_J:O_J + 225].reshape(15, 15)
    assert np.allclose(J[0:3, 9:12], -0.5 * T * T * np.eye(3), rtol=1e-10, atol=1e-14)   # d delta_p / d ba
    assert np.allclose(J[6:9, 9:12], -T * np.eye(3), rtol=1e-10, atol=1e-14)            # d delta_v / d ba
    sq = r[O_SQ:O_SQ + 225].reshape(15, 15)
    assert np.all(np.tril(sq, -1) == 0) and np.all(np.diag(sq) > 0)                      # matrixL().transpose()


def test_oracle_matches_the_generators_independent_implementation():
    begin, s, bias = streams(12, 7)
    rec, bad = ob.preintegrate_batch(begin, s, bias, NOISE)
    assert bad == 0
    L = swgn.synth_lib()
    L.swgn_synth_preintegrate.restype = C.c_int32
    L.swgn_synth_preintegrate.argtypes = [C.c_int32] + [C.POINTER(C.c_double)] * 4
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for f in range(12):
        out = np.zeros(474)
        seg = np.ascontiguousarray(s[begin[f]:begin[f + 1]])
        assert L.swgn_synth_preintegrate(len(seg), dp(seg), dp(np.ascontiguousarray(bias[f])), dp(NOISE), dp(out)) == 0
        # the generator rotates with the matrix of the un-normalised quaternion where Eigen (and the oracle) use
        # _transformVector: the two differ at O(|omega dt|^2) ~ 1e-7 per step
        for a, b, n in [(O_DP, O_DP + 3, "dp"), (O_DQ, O_DQ + 4, "dq"), (O_DV, O_DV + 3, "dv"), (O_J, O_J + 225, "jac")]:
            assert np.linalg.norm(rec[f, a:b] - out[a:b]) < 1e-5 * max(1e-3, np.linalg.norm(out[a:b])), n
        assert abs(rec[f, O_DT] - out[O_DT]) < 1e-12
        sa, sb = rec[f, O_SQ:O_SQ + 225].reshape(15, 15), out[O_SQ:O_SQ + 225].reshape(15, 15)
        assert np.linalg.norm(sa.T @ sa - sb.T @ sb) < 1e-4 * np.linalg.norm(sb.T @ sb)   # the information matrix


@pytest.mark.gpu
def test_gpu_preintegration_matches_oracle():
    begin, s, bias = streams(300, 11, lo=3)
    ro, bad = ob.preintegrate_batch(begin, s, bias, NOISE)
    rg, info = swgn.preintegrate_batch(begin, s, bias, NOISE)
    assert bad == 0 and not info.any()
    for a, b in [(O_DP, O_DP + 3), (O_DQ, O_DQ + 4), (O_DV, O_DV + 3), (O_DT, O_DT + 1), (O_J, O_J + 225)]:
        err = np.linalg.norm(rg[:, a:b] - ro[:, a:b], axis=1) / np.maximum(1e-300, np.linalg.norm(ro[:, a:b], axis=1))
        assert err.max() < 1e-12, (a, err.max())
    assert np.array_equal(rg[:, 10:22], ro[:, 10:22])  # linearisation biases, gyr_i, gyr_j copied through
    # sqrt_info = LLT(cov^-1).L': cond(cov) ~ 1e10, so two correct evaluations agree to ~1e-6
    err = np.linalg.norm(rg[:, O_SQ:] - ro[:, O_SQ:], axis=1) / np.linalg.norm(ro[:, O_SQ:], axis=1)
    assert err.max() < 1e-5
    sq = rg[:, O_SQ:].reshape(-1, 15, 15)
    assert np.all(np.tril(sq, -1) == 0)


@pytest.mark.gpu
def test_gpu_preintegration_records_drive_the_same_solve():
    """End to end: an IMU factor built from the device's record gives the residual the oracle computes from
    the oracle's record (same states), i.e. the record layout is the one the factor kernels consume."""
    begin, s, bias = streams(4, 3, lo=100, hi=100)
    ro, _ = ob.preintegrate_batch(begin, s, bias, NOISE)
    rg, _ = swgn.preintegrate_batch(begin, s, bias, NOISE)
    gl = np.array([-0.005, 0.009, 0.31, 0.1, -0.2, 9.78, 1, 0, 0, 1.0])
    par = np.array([0, 0, 0, 0, 0, 0, 1, 0.1, 0, 0, 0.01, 0.01, 0.01, 0.001, 0.001, 0.001,
                    0.02, 0.01, 0.3, 0, 0, 0.02, 0.9998, 0.2, 0.1, 2.4, 0.01, 0.01, 0.01, 0.001, 0.001, 0.001])
    for f in range(4):
        r1, r2 = np.zeros(15), np.zeros(15)
        assert ob.oracle().oracle_factor_eval(1, 0, ob._dp(gl), ob._dp(np.ascontiguousarray(ro[f])), ob._dp(par), ob._dp(r1), None) == 0
        assert ob.oracle().oracle_factor_eval(1, 0, ob._dp(gl), ob._dp(np.ascontiguousarray(rg[f])), ob._dp(par), ob._dp(r2), None) == 0
        assert np.linalg.norm(r1 - r2) < 1e-6 * np.linalg.norm(r1)


@pytest.mark.gpu
def test_gpu_preintegration_reports_a_singular_covariance():
    """One midpoint step leaves the 15x15 covariance singular: get_sqrtinfo's LLT cannot succeed (the
    reference would hand NaNs to the solver); both implementations flag the factor instead."""
    begin, s, bias = streams(5, 2, lo=2, hi=2)
    ro, bad = ob.preintegrate_batch(begin, s, bias, NOISE)
    rg, info = swgn.preintegrate_batch(begin, s, bias, NOISE)
    assert bad == 5 and info.tolist() == [1] * 5
    assert np.all(rg[:, O_SQ:] == 0)
    assert np.allclose(rg[:, :O_SQ], ro[:, :O_SQ], rtol=1e-12, atol=1e-300)
