"""Pins the oracle's integer-ambiguity and range-model restatements on the reference's OWN code:
RVI/gnss/src/lambda.cpp and common_function.cpp compiled from /root/reference into
oracle/_ref/libref_gnss.so (oracle/build_ref.sh).  Bit-exact comparisons."""
import ctypes as C

import numpy as np
import pytest

import oracle_binding as ob

pytestmark = pytest.mark.skipif(ob.ref() is None, reason="oracle/_ref/libref_gnss.so not built")


def random_problem(rng, n, scale=1.0):
    """Correlated ambiguity covariance like a short-baseline float solution."""
    B = rng.normal(size=(n, n + 3))
    Q = B @ B.T * scale + 1e-3 * np.eye(n)
    # add strong common-mode correlation (what makes decorrelation do real work)
    u = rng.normal(size=(n, 1))
    Q += 50.0 * scale * (u @ u.T)
    Q = 0.5 * (Q + Q.T)
    a = rng.uniform(-30, 30, size=n)
    return a, Q


@pytest.mark.parametrize("n", [2, 3, 6, 10, 17, 24, 30])
def test_lambda_bit_exact_against_reference(n):
    rng = np.random.default_rng(100 + n)
    for trial in range(20):
        a, Q = random_problem(rng, n, scale=10.0 ** rng.integers(-3, 1))
        io, Fo, so = ob.lambda_search(a, Q, 2, "oracle")
        ir, Fr, sr = ob.lambda_search(a, Q, 2, "ref")
        assert io == ir
        if ir == 0:
            assert np.array_equal(Fo, Fr)
            assert np.array_equal(so, sr)  # bit-exact squared norms
            assert so[0] <= so[1]
            assert np.abs(Fo - np.round(Fo)).max() < 1e-6  # Z^-T E by LU: integers up to rounding


def test_lambda_failure_codes_match_reference():
    a = np.array([0.3, 1.2, -2.2])
    Q = -np.eye(3)  # not positive definite: LD fails
    assert ob.lambda_search(a, Q, 2, "oracle")[0] == ob.lambda_search(a, Q, 2, "ref")[0] == -1


def test_lambda_recovers_known_integers():
    rng = np.random.default_rng(5)
    n = 12
    z = rng.integers(-40, 40, size=n).astype(float)
    a, Q = random_problem(rng, n, scale=1e-4)
    a = z + rng.multivariate_normal(np.zeros(n), Q * 1e-2)
    info, F, s = ob.lambda_search(a, Q)
    assert info == 0
    assert np.array_equal(np.round(F[:, 0]), z)
    assert s[1] / s[0] > 2


@pytest.mark.parametrize("n", [1, 2, 5, 17, 30])
def test_matinv_bit_exact(n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n)) + n * np.eye(n)
    A1 = np.asfortranarray(A.copy())
    A2 = np.asfortranarray(A.copy())
    P = C.POINTER(C.c_double)
    st_o = ob.oracle().oracle_matinv(A1.ctypes.data_as(P), n)
    st_r = ob.ref().ref_matinv(A2.ctypes.data_as(P), n)
    # the reference returns 1 on success (common_function.cpp:348-366), the oracle 0
    assert st_o == 0 and (st_r & 0xff) == 1
    assert np.array_equal(A1, A2)
    np.testing.assert_allclose(A1 @ A, np.eye(n), atol=1e-9)


def test_distance_and_velocity_distance_bit_exact():
    rng = np.random.default_rng(3)
    P = C.POINTER(C.c_double)
    for _ in range(200):
        rr = rng.normal(size=3) * 6.4e6
        rs = rng.normal(size=3) * 2.6e7
        vr = rng.normal(size=3) * 3.0
        vs = rng.normal(size=3) * 3e3
        e1, e2 = np.zeros(3), np.zeros(3)
        d1 = ob.oracle().oracle_distance(rr.ctypes.data_as(P), rs.ctypes.data_as(P), e1.ctypes.data_as(P))
        d2 = ob.ref().ref_distance(rr.ctypes.data_as(P), rs.ctypes.data_as(P), e2.ctypes.data_as(P))
        assert d1 == d2 and np.array_equal(e1, e2)
        v1 = ob.oracle().oracle_velocity_distance(*(x.ctypes.data_as(P) for x in (rr, rs, vr, vs, e1)))
        v2 = ob.ref().ref_velecitydistance(*(x.ctypes.data_as(P) for x in (rr, rs, vr, vs, e2)))
        assert v1 == v2 and np.array_equal(e1, e2)
