"""Pins the traditional dogleg step (a13) on the fixtures of the reference's own
CERES/internal/ceres/dogleg_strategy_test.cc: `DoglegStrategyFixtureEllipse` (:54-91, J'J = Q diag(1..32) Q',
minimum of the quadratic at (1,..,1)) and `DoglegStrategyFixtureValley` (:93-118, J = diag(1,2,4,8,16,32), minimum
at e_3), both with min_lm_diagonal = max_lm_diagonal = 1 and x = 0.

The reference calls DoglegStrategy::ComputeStep directly; here the same 6-D linear problems go through the whole
path (one dense linear factor, DENSE_SCHUR with one eliminated scalar, ONE trust-region iteration from x = 0): the
model is exact for a linear problem, so the step is accepted and the returned state IS the step.  Checked:
  * TrustRegionObeyedTraditional (:124-146): |step| <= radius * (1 + 4 eps) at radius 2;
  * CorrectGaussNewtonStep (:168-191): radius 10 -> the minimum, 1e-5;
  * the Valley expectations (:229-283; the gradient points at the minimum, so the traditional dogleg takes the
    same steps as the subspace variant the reference runs there): 0.25 e_3 at radius 0.25, e_3 at radius 2;
  * beyond the reference: the step equals the textbook dogleg (dogleg_strategy.cc:199-253) computed in numpy.
CPU tests pin the oracle; the gpu tests push the same graphs through the C ABI."""
import numpy as np
import pytest

import oracle_binding as ob
import swgn
from linear_graph import LinearGraph

# dogleg_strategy_test.cc:66-71
BASIS = np.array([
    [-0.1046920933796121, -0.7449367449921986, -0.4190744502875876, -0.4480450716142566, 0.2375351607929440, -0.0363053418882862],
    [0.4064975684355914, 0.2681113508511354, -0.7463625494601520, -0.0803264850508117, -0.4463149623021321, 0.0130224954867195],
    [-0.5514387729089798, 0.1026621026168657, -0.5008316122125011, 0.5738122212666414, 0.2974664724007106, 0.1296020877535158],
    [0.5037835370947156, 0.2668479925183712, -0.1051754618492798, -0.0272739396578799, 0.7947481647088278, -0.1776623363955670],
    [-0.4005458426625444, 0.2939330589634109, -0.0682629380550051, -0.2895448882503687, -0.0457239396341685, -0.8139899477847840],
    [-0.3247764582762654, 0.4528151365941945, -0.0276683863102816, -0.6155994592510784, 0.1489240599972848, 0.5362574892189350]])
DDIAG = np.array([1.0, 2.0, 4.0, 8.0, 16.0, 32.0])
EPS = np.finfo(float).eps


def fixture(name):
    if name == "ellipse":
        J = np.diag(np.sqrt(DDIAG)) @ BASIS
        minimum = np.ones(6)
    else:
        J = np.diag(DDIAG)
        minimum = np.array([0.0, 0.0, 1.0, 0.0, 0.0, 0.0])
    return J, -J @ minimum, minimum


def graph(J, r):
    # six scalar parameter blocks, the first one eliminated; one 6-row block touching all of them
    p = {"col_sizes": [1] * 6, "num_eliminate_blocks": 1, "row_sizes": [6], "row_ptr": [0, 6], "cell_col": list(range(6)),
         "values": [float(J[i, c]) for c in range(6) for i in range(6)], "b": [float(v) for v in r]}
    return LinearGraph(p)


def options(radius):
    opt = swgn.default_options()
    opt.min_lm_diagonal = 1.0
    opt.max_lm_diagonal = 1.0
    opt.initial_trust_region_radius = radius
    opt.max_trust_region_radius = radius
    opt.max_num_iterations = 1
    return opt


def textbook_dogleg(J, r, radius, mu=1e-12):
    g = J.T @ r
    alpha = (g @ g) / ((J @ g) @ (J @ g))
    cauchy = -alpha * g
    gn = -np.linalg.solve(J.T @ J + mu * np.eye(6), g)
    if np.linalg.norm(gn) <= radius:
        return gn
    if np.linalg.norm(cauchy) >= radius:
        return -(radius / np.linalg.norm(g)) * g
    d = gn - cauchy
    a, b, c = d @ d, 2 * cauchy @ d, cauchy @ cauchy - radius * radius
    beta = (-b + np.sqrt(b * b - 4 * a * c)) / (2 * a)
    return cauchy + beta * d


def solve_oracle(lg, opt):
    o = ob.OracleSolver(lg.graph_p, opt)
    _, sm = o.minimize()
    return o.state().copy(), sm


def solve_gpu(lg, opt):
    b = swgn.Batch([lg.graph_p], opt)
    sm = b.solve()[0]
    x = b.get_state(0, 6).copy()
    b.close()
    return x, sm


CASES = [("ellipse", 2.0), ("ellipse", 10.0), ("valley", 0.25), ("valley", 2.0)]


def check(name, radius, solve):
    J, r, minimum = fixture(name)
    lg = graph(J, r)
    # the linear-graph helper reproduces the fixture: residual = r and Jacobian = J at x = 0
    o = ob.OracleSolver(lg.graph_p, options(radius))
    _, res, _, Jd = o.evaluate()
    np.testing.assert_allclose(res, r, rtol=0, atol=0)
    cols, offs, sizes = o.columns()
    assert sorted(cols.tolist()) == list(range(6)) and cols[0] == 0
    np.testing.assert_allclose(Jd[:, np.argsort(cols)], J, rtol=0, atol=0)
    x, sm = solve(lg, options(radius))
    assert sm.termination_type != 2  # not a failure
    assert sm.num_unsuccessful_steps == 0 and sm.num_iterations == 1  # the one step was accepted
    assert np.linalg.norm(x) <= radius * (1.0 + 4.0 * EPS)
    np.testing.assert_allclose(x, textbook_dogleg(J, r, radius), rtol=0, atol=1e-9)
    if (name, radius) == ("ellipse", 10.0):
        np.testing.assert_allclose(x, np.ones(6), atol=1e-5)
    if name == "valley":
        want = np.zeros(6)
        want[2] = min(radius, 1.0)
        np.testing.assert_allclose(x, want, atol=1e-5)
    return x


@pytest.mark.parametrize("name,radius", CASES)
def test_oracle_dogleg_on_the_reference_fixtures(name, radius):
    check(name, radius, solve_oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("name,radius", CASES)
def test_gpu_dogleg_on_the_reference_fixtures(name, radius):
    x = check(name, radius, solve_gpu)
    J, r, _ = fixture(name)
    xo, _ = solve_oracle(graph(J, r), options(radius))
    np.testing.assert_allclose(x, xo, rtol=0, atol=1e-12)


# ---- Levenberg-Marquardt strategy: restated by the oracle only so far (the device answers SWGN_ERR_UNSUPPORTED)
def lm_options(radius, iters):
    opt = options(radius)
    opt.max_num_iterations = iters
    opt.trust_region_strategy = 1  # SWGN_LEVENBERG_MARQUARDT
    return opt


@pytest.mark.parametrize("name", ["ellipse", "valley"])
def test_oracle_levenberg_marquardt_step_and_radius_update(name):
    """levenberg_marquardt_strategy.cc:67-165 on the same fixtures: the first step is -(J'J + D^2)^-1 J'r with
    D^2 = clamp(colnorm^2) / radius (= 1 / radius here, min = max diagonal = 1); the model is exact for a linear
    problem, so rho = 1 and the accepted step multiplies the radius by 3 (radius / max(1/3, 1 - (2 rho - 1)^3)),
    capped by max_radius; a few iterations reach the minimum."""
    J, r, minimum = fixture(name)
    lg = graph(J, r)
    radius = 4.0
    o = ob.OracleSolver(lg.graph_p, lm_options(radius, 1))
    _, sm = o.minimize()
    x = o.state().copy()
    want = -np.linalg.solve(J.T @ J + np.eye(6) / radius, J.T @ r)
    np.testing.assert_allclose(x, want, rtol=0, atol=1e-12)
    assert sm.num_iterations == 1 and sm.num_unsuccessful_steps == 0
    opt = lm_options(radius, 6)
    opt.max_trust_region_radius = 1e16
    o = ob.OracleSolver(lg.graph_p, opt)
    _, sm = o.minimize()
    costs, radii, ok = o.iteration_records()
    assert all(ok) and np.all(np.diff(costs) < 0)
    np.testing.assert_allclose(radii[1:] / radii[:-1], 3.0, rtol=1e-6)  # rho = 1 up to rounding
    np.testing.assert_allclose(o.state(), minimum, atol=1e-4)


def test_oracle_jacobi_scaling_on_the_first_levenberg_marquardt_step():
    """Solver::Options::jacobi_scaling (trust_region_minimizer.cc:261-276,437): columns scaled by s = 1 / (1 + |column|) of
    the initial Jacobian; the strategy works on J diag(s), the step is mapped back with s."""
    J, r, _ = fixture("ellipse")
    J = J * np.array([1.0, 30.0, 0.02, 4.0, 1.0, 0.3])  # unequal column norms so that the scaling matters
    lg = graph(J, r)
    radius = 4.0
    opt = lm_options(radius, 1)
    opt.jacobi_scaling = 1
    # (Marquardt's diagonal diag(J'J) / radius is invariant under column scaling: only the clamp makes the two differ)
    opt.min_lm_diagonal, opt.max_lm_diagonal = 0.5, 1e32
    o = ob.OracleSolver(lg.graph_p, opt)
    _, sm = o.minimize()
    sc = 1.0 / (1.0 + np.linalg.norm(J, axis=0))
    Js = J * sc
    D2 = np.clip((Js * Js).sum(axis=0), 0.5, 1e32) / radius
    want = sc * -np.linalg.solve(Js.T @ Js + np.diag(D2), Js.T @ r)
    np.testing.assert_allclose(o.state(), want, rtol=0, atol=1e-12 * max(1.0, np.abs(want).max()))
    opt.jacobi_scaling = 0
    o2 = ob.OracleSolver(lg.graph_p, opt)
    o2.minimize()
    assert np.abs(o2.state() - want).max() > 1e-6  # the scaling changes the damped step


def test_device_rejects_dogleg_with_jacobi_scaling_loudly():
    J, r, _ = fixture("valley")
    lg = graph(J, r)
    opt = options(2.0)
    opt.jacobi_scaling = 1
    with pytest.raises(RuntimeError, match="jacobi_scaling"):
        swgn.Batch([lg.graph_p], opt)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ellipse", "valley"])
@pytest.mark.parametrize("jacobi", [0, 1])
def test_gpu_levenberg_marquardt_on_the_reference_fixtures(name, jacobi):
    """The device's LEVENBERG_MARQUARDT strategy against the oracle (pinned above on the closed-form damped step and the
    x3 radius update): one iteration and a converged run, with and without jacobi_scaling."""
    J, r, minimum = fixture(name)
    J = J * np.array([1.0, 30.0, 0.02, 4.0, 1.0, 0.3])
    lg = graph(J, r)
    for iters in (1, 2, 8):
        opt = lm_options(4.0, iters)
        opt.jacobi_scaling = jacobi
        opt.min_lm_diagonal, opt.max_lm_diagonal = 0.5, 1e32
        o = ob.OracleSolver(lg.graph_p, opt)
        _, osm = o.minimize()
        b = swgn.Batch([lg.graph_p], opt)
        sm = b.solve()[0]
        x = b.get_state(0, len(lg.state))
        b.close()
        assert (sm.num_iterations, sm.num_successful_steps, sm.num_unsuccessful_steps, sm.termination_type) == \
            (osm.num_iterations, osm.num_successful_steps, osm.num_unsuccessful_steps, osm.termination_type)
        np.testing.assert_allclose(x, o.state(), rtol=0, atol=1e-10 * max(1.0, np.abs(o.state()).max()))
        assert abs(sm.final_cost - osm.final_cost) <= 1e-9 * max(osm.final_cost, 1e-12)
