"""Pins the oracle's Schur eliminate / reduced solve / back-substitute on the reference's own
known-answer fixtures: CERES/internal/ceres/linear_least_squares_problems.cc problems 2, 3, 4
(block-sparse definitions :283-515,517-625; hand-computed A'A, S, r, S\\r, A\\b in the comment at
:135-178) and the dense-reference construction of schur_eliminator_test.cc:83-183 (tolerance
1e-14 relative at :202-222)."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixtures():
    with open(os.path.join(HERE, "golden", "ceres_llsq_problems.json")) as f:
        return json.load(f)


def run_raw(p, use_D):
    col_sizes = np.array(p["col_sizes"], np.int32)
    row_sizes = np.array(p["row_sizes"], np.int32)
    row_ptr = np.array(p["row_ptr"], np.int32)
    cell_col = np.array(p["cell_col"], np.int32)
    values = np.array(p["values"], np.float64)
    b = np.array(p["b"], np.float64)
    D = np.array(p["D"], np.float64)
    ne = p["num_eliminate_blocks"]
    ncols = int(col_sizes.sum())
    nf = int(col_sizes[ne:].sum())
    lhs = np.zeros((nf, nf))
    rhs = np.zeros(nf)
    x = np.zeros(ncols)
    L = ob.oracle()
    st = L.oracle_schur_raw(len(col_sizes), ob._ip(col_sizes), len(row_sizes), ob._ip(row_sizes),
                            ob._ip(row_ptr), ob._ip(cell_col), ob._dp(values), ob._dp(b),
                            ob._dp(D) if use_D else None, ne, ob._dp(lhs), ob._dp(rhs), ob._dp(x))
    return st, lhs, rhs, x


def dense(p):
    col_sizes = p["col_sizes"]
    pos = np.concatenate([[0], np.cumsum(col_sizes)])
    nrows = sum(p["row_sizes"])
    A = np.zeros((nrows, pos[-1]))
    v = 0
    r0 = 0
    for i, rs in enumerate(p["row_sizes"]):
        for k in range(p["row_ptr"][i], p["row_ptr"][i + 1]):
            c = p["cell_col"][k]
            cs = col_sizes[c]
            A[r0:r0 + rs, pos[c]:pos[c] + cs] = np.array(p["values"][v:v + rs * cs]).reshape(rs, cs)
            v += rs * cs
        r0 += rs
    return A


def schur_reference(p, use_D):
    """schur_eliminator_test.cc:83-130: J = [A; diag(D)], H = J'J, g = J'[b;0],
    S = R - Q' P^-1 Q, r = g_f - Q' P^-1 g_e, solution by dense solve."""
    A = dense(p)
    n = A.shape[1]
    ne = sum(p["col_sizes"][:p["num_eliminate_blocks"]])
    J = np.vstack([A, np.diag(p["D"])]) if use_D else A
    f = np.concatenate([p["b"], np.zeros(n)]) if use_D else np.array(p["b"], float)
    H = J.T @ J
    g = J.T @ f
    P, Q, R = H[:ne, :ne], H[:ne, ne:], H[ne:, ne:]
    S = R - Q.T @ np.linalg.solve(P, Q)
    r = g[ne:] - Q.T @ np.linalg.solve(P, g[:ne])
    x = np.linalg.solve(H, g)
    return S, r, x


def test_problem2_hand_computed_comment_values():
    """The numbers printed in the reference's comment (D = 0)."""
    p = load_fixtures()["problem2"]
    st, lhs, rhs, x = run_raw(p, use_D=False)
    assert st == 0
    S = np.triu(lhs) + np.triu(lhs, 1).T
    g = p["golden"]
    np.testing.assert_allclose(S, np.array(g["S"]), atol=6e-5)
    # the comment prints r[2] = 5.0323; its own S\\r = [0.2102 2.1367 0.1388] (and A\\b) only follow
    # from r[2] = 4.0323 = 17 - 30*67/155, i.e. the printed digit is a typo in the reference.
    r_doc = np.array(g["r"])
    assert abs(r_doc[2] - 5.0323) < 1e-12
    r_doc[2] = 4.0323
    np.testing.assert_allclose(rhs, r_doc, atol=6e-5)
    np.testing.assert_allclose(np.linalg.solve(np.array(g["S"]), r_doc), np.array(g["S_solve_r"]), atol=2e-4)
    np.testing.assert_allclose(x[2:], np.array(g["S_solve_r"]), atol=6e-5)
    np.testing.assert_allclose(x, np.array(g["A_solve_b"]), atol=6e-5)
    A = dense(p)
    np.testing.assert_allclose(A.T @ A, np.array(g["AtA"]), atol=0)
    np.testing.assert_allclose(A.T @ np.array(p["b"]), np.array(g["c"]), atol=0)


@pytest.mark.parametrize("name", ["problem2", "problem3", "problem4"])
@pytest.mark.parametrize("use_D", [True, False])
def test_schur_against_dense_reference(name, use_D):
    p = load_fixtures()[name]
    if name == "problem4" and not use_D:
        pytest.skip("rank deficient without the diagonal (reference comment :533-534)")
    st, lhs, rhs, x = run_raw(p, use_D)
    assert st == 0
    S_ref, r_ref, x_ref = schur_reference(p, use_D)
    nf = lhs.shape[0]
    if nf:
        S = np.triu(lhs) + np.triu(lhs, 1).T
        # relative 1e-14 like schur_eliminator_test.cc:202-222
        assert np.linalg.norm(S - S_ref) / np.linalg.norm(S_ref) < 1e-14
        assert np.linalg.norm(rhs - r_ref) / np.linalg.norm(r_ref) < 1e-14
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-13


def test_random_block_sparse_against_dense_reference():
    """Same check on random block-sparse systems with the e-block sizes the window uses
    (1, 3, 9) and mixed row sizes, cf. BlockSparseMatrix::CreateRandomMatrix users."""
    rng = np.random.default_rng(7)
    for trial in range(5):
        e_sizes = list(rng.choice([1, 3, 9], size=6))
        f_sizes = list(rng.choice([1, 6, 9], size=5))
        col_sizes = e_sizes + f_sizes
        ne = len(e_sizes)
        row_sizes, row_ptr, cell_col, values = [], [0], [], []
        for e in range(ne):
            for _ in range(int(rng.integers(2, 5))):
                rs = int(rng.choice([1, 2, 15]))
                fs = sorted(rng.choice(len(f_sizes), size=int(rng.integers(0, 4)), replace=False))
                cells = [e] + [ne + int(f) for f in fs]
                row_sizes.append(rs)
                for c in cells:
                    cell_col.append(c)
                    values += list(rng.normal(size=rs * col_sizes[c]))
                row_ptr.append(len(cell_col))
        for _ in range(3):  # rows without e-block
            rs = int(rng.choice([1, 4]))
            fs = sorted(rng.choice(len(f_sizes), size=2, replace=False))
            row_sizes.append(rs)
            for f in fs:
                cell_col.append(ne + int(f))
                values += list(rng.normal(size=rs * col_sizes[ne + int(f)]))
            row_ptr.append(len(cell_col))
        p = dict(col_sizes=[int(c) for c in col_sizes], row_sizes=row_sizes, row_ptr=row_ptr,
                 cell_col=cell_col, values=values, b=list(rng.normal(size=sum(row_sizes))),
                 D=list(rng.uniform(0.5, 2.0, size=sum(col_sizes))), num_eliminate_blocks=ne)
        st, lhs, rhs, x = run_raw(p, True)
        assert st == 0
        S_ref, r_ref, x_ref = schur_reference(p, True)
        S = np.triu(lhs) + np.triu(lhs, 1).T
        assert np.linalg.norm(S - S_ref) / np.linalg.norm(S_ref) < 1e-13
        assert np.linalg.norm(rhs - r_ref) / np.linalg.norm(r_ref) < 1e-12
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-11


def test_invert_psd_matrix_on_the_reference_test_properties():
    """CERES/internal/ceres/invert_psd_matrix_test.cc: Identity3x3 (:54-60, relative error <= eps) and the full-rank
    5x5 cases, fixed and dynamic (:62-71, :88-98): |m inv(m) - I| / 5 <= 10 eps for m = Q diag(|lambda|) Q'.  The
    eliminator only takes the assume_full_rank branch (schur_eliminator_impl.h:279-280), which is what the oracle
    restates; the eigenvalues are drawn from [0.25, 1] so that the reference's bound holds for every seed."""
    L = ob.oracle()
    eps = np.finfo(float).eps
    eye = np.eye(3)
    inv = np.zeros((3, 3))
    assert L.oracle_invert_psd(ob._dp(eye), 3, ob._dp(inv)) == 1
    assert np.linalg.norm(inv - eye) / np.linalg.norm(eye) <= eps
    rng = np.random.default_rng(5)
    for n in (1, 3, 5, 9, 15):
        for _ in range(20):
            q, _r = np.linalg.qr(rng.standard_normal((n, n)))
            lam = rng.uniform(0.25, 1.0, n)
            m = np.ascontiguousarray((q * lam) @ q.T)
            m = 0.5 * (m + m.T)
            inv = np.zeros((n, n))
            assert L.oracle_invert_psd(ob._dp(m), n, ob._dp(inv)) == 1
            assert np.linalg.norm(m @ inv - np.eye(n)) / n <= 10 * eps
    # not positive definite: the LLT fails and the caller is told
    bad = np.diag([1.0, -1.0, 1.0])
    assert L.oracle_invert_psd(ob._dp(bad), 3, ob._dp(np.zeros((3, 3)))) == 0


def _scalar_graph(rows, groups):
    """x, y, z scalar blocks; rows = parameter lists of the residual blocks in program order."""
    from linear_graph import LinearGraph
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    p = {"col_sizes": [1, 1, 1], "num_eliminate_blocks": 2, "row_sizes": [1] * len(rows), "row_ptr": [int(v) for v in row_ptr],
         "cell_col": [c for r in rows for c in r], "values": [1.0 + 0.1 * k for k in range(int(row_ptr[-1]))],
         "b": [1.0] * len(rows)}
    lg = LinearGraph(p)
    lg.block_group[:] = groups
    return lg


def test_residual_block_order_on_the_reference_fixture():
    """CERES/internal/ceres/reorder_program_test.cc:64-116 (ReorderResidualBlockNormalFunction): x, y in group 0, z in
    group 1, residual blocks (x), (z,x), (z,y), (z), (x,y), (y); expected order 4, 1, 0, 5, 2, 3 -- every e-block's
    bucket filled back to front.  The reference calls LexicographicallyOrderResidualBlocks directly; through the
    whole preprocessor residual block 4 = (x, y) violates the independence of group 0 (program.cc:413-434, rejected
    here as in Ceres), so the pinned vector is the same fixture without it: 1, 0, 5, 2, 3."""
    import swgn
    X, Y, Z = 0, 1, 2
    rows = [[X], [Z, X], [Z, Y], [Z], [Y]]  # program order without (x, y)
    lg = _scalar_graph(rows, [0, 0, 1])
    o = ob.OracleSolver(lg.graph_p, swgn.default_options())
    factor, _ = o.rows()
    assert factor.tolist() == [1, 0, 4, 2, 3]  # = the reference's 1, 0, 5, 2, 3 with block 5 renumbered to 4
    cols, _, _ = o.columns()
    assert cols.tolist() == [X, Y, Z]
    st, pcols, prows = swgn.plan_order(lg.graph_p, 0)  # the host planner of the CUDA path, same vector
    assert st == 0 and pcols.tolist() == [X, Y, Z] and prows.tolist() == [1, 0, 4, 2, 3]
    # with (x, y) present the first elimination group is not an independent set
    lg_bad = _scalar_graph([[X], [Z, X], [Z, Y], [Z], [X, Y], [Y]], [0, 0, 1])
    with pytest.raises(RuntimeError):
        ob.OracleSolver(lg_bad.graph_p, swgn.default_options())
    st, _ = swgn.plan_probe(lg_bad.graph_p, 0)
    assert st != 0 and b"independent" in swgn.lib().swgn_last_error()


def test_apply_ordering_on_the_reference_fixture():
    """reorder_program_test.cc:138-165 (ApplyOrderingNormal): x -> group 0, y -> group 2, z -> group 1 orders the
    parameter blocks x, z, y."""
    import swgn
    X, Y, Z = 0, 1, 2
    lg = _scalar_graph([[X], [Z, X], [Z, Y], [Z], [Y]], [0, 2, 1])
    o = ob.OracleSolver(lg.graph_p, swgn.default_options())
    cols, _, _ = o.columns()
    assert cols.tolist() == [X, Z, Y]
    st, pcols, _ = swgn.plan_order(lg.graph_p, 0)
    assert st == 0 and pcols.tolist() == [X, Z, Y]
