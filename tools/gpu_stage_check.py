"""Stage-by-stage comparison of the CUDA path with the CPU oracle on one synthetic window
(development aid; the assertions live in tests/test_gpu_parity.py).
usage: python tools/gpu_stage_check.py [which=2] [window_id=0]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build_if_needed()
import oracle_binding as ob  # noqa: E402
import swgn  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    which = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    wid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    w = swgn.SynthWindow(which, wid)
    opt = w.options()
    o = ob.OracleSolver(w.graph_p, opt)
    b = swgn.Batch([w.graph_p], opt)
    print("dims: res", o.n_res, "cols", o.n_cols, "n_e", o.n_e, "n_f", o.n_f, "schur bytes", b.schur_bytes(0))
    cb, co, cs = b.columns(0)
    ob_, oo, os_ = o.columns()
    print("columns equal:", np.array_equal(cb, ob_) and np.array_equal(co, oo) and np.array_equal(cs, os_))
    rf, ro = b.rows(0)
    of, oro = o.rows()
    print("rows equal:", np.array_equal(rf, of) and np.array_equal(ro, oro))
    cost, r, g = b.evaluate(0, o.n_res, o.n_cols)
    ocost, orr, og, oJ = o.evaluate()
    J = b.dense_jacobian(0, o.n_res, o.n_cols)
    print("cost", cost, ocost, "rel", abs(cost - ocost) / ocost)
    print("residual rel err", rel(r, orr), "max abs", np.abs(r - orr).max())
    print("jacobian rel err", rel(J, oJ), "max abs", np.abs(J - oJ).max())
    bad = np.argwhere(np.abs(J - oJ) > 1e-6 * (1 + np.abs(oJ)))
    if len(bad):
        print("  first bad J entries:", bad[:10].tolist())
    print("gradient rel err", rel(g, og))
    rng = np.random.default_rng(0)
    D = rng.uniform(0.5, 1.5, o.n_cols) * 1e-3
    x = b.linear_solve(0, D, o.n_cols)
    st, ox, oS, orhs = o.linear_solve(D)
    S, rhs = b.get_reduced(0)
    print("S rel err", rel(np.triu(S), np.triu(oS)), "rhs rel err", rel(rhs, orhs))
    print("linear solve x rel err", rel(x, ox), " (e part", rel(x[:o.n_e], ox[:o.n_e]), " f part", rel(x[o.n_e:], ox[o.n_e:]), ")")
    H = oJ.T @ oJ + np.diag(D * D)
    sv = np.linalg.svd(H, compute_uv=False)
    cond = sv[0] / sv[-1]
    be = lambda v: float(np.linalg.norm(H @ v - og) / (np.linalg.norm(H, 2) * np.linalg.norm(v) + np.linalg.norm(og)))
    print("  cond(J'J + D^2) %.3e -> cond * eps %.3e ; backward error gpu %.3e oracle %.3e" % (cond, cond * 2.2e-16, be(x), be(ox)))
    sm = b.solve()[0]
    ost, osm = o.minimize()
    xs = b.get_state(0, w.n_state)
    xo = o.state()
    print("solve: gpu cost %.12g oracle %.12g | iters %d/%d | successful %d/%d | term %d/%d | linear solves %d/%d" %
          (sm.final_cost, osm.final_cost, sm.num_iterations, osm.num_iterations, sm.num_successful_steps,
           osm.num_successful_steps, sm.termination_type, osm.termination_type, sm.num_linear_solves, osm.num_linear_solves))
    print("state max abs err", np.abs(xs - xo).max(), "max rel", (np.abs(xs - xo) / np.maximum(1, np.abs(xo))).max())
    for name, err in swgn.state_error_by_kind(w, xs, xo).items():
        print("   %-10s max |dx| / max(1, |x|) = %.3e" % (name, err))
    print("timing (total ms, schur ms, schur launches, launches):", b.timing())
    if opt.n_parameter_head > 0:
        Lg = b.get_cholesky(0)
        n, oS2, or2, Lo = o.exports()
        print("cholesky factor rel err", rel(Lg, Lo))
        nt = w.n_amb
        A = b.tail_information(0, nt)
        Ao = ob.tail_information(Lo, nt)
        print("tail information rel err", rel(A, Ao))
        eb, oa, sf = w.ambiguity_epochs()
        offs = w.block_offsets()
        y = np.array([xs[offs[w.first_amb_block + k]] for k in range(nt)])
        yo = np.array([xo[offs[w.first_amb_block + k]] for k in range(nt)])
        pg, Fg, rg = swgn.ambiguity_fix(Ao, yo, eb, oa, sf)
        po, Fo, ro_ = ob.ambiguity_fix(Ao, yo, eb, oa, sf)
        print("fix (same inputs): pairs equal", np.array_equal(pg, po), "F equal", np.array_equal(Fg, Fo), "s", list(rg.s), list(ro_.s),
              "ok", rg.search_ok, ro_.search_ok, "status", rg.status, ro_.status)
        pg2, Fg2, rg2 = swgn.ambiguity_fix(A, y, eb, oa, sf)
        print("fix (gpu inputs): F[:,0] equal oracle", np.array_equal(Fg2[:, 0], Fo[:, 0]) if Fg2.shape == Fo.shape else None, "ok", rg2.search_ok,
              "true N diffs recovered:", np.array_equal(Fg2[:, 0], (w.true_ambiguities()[pg2[:, 0]] - w.true_ambiguities()[pg2[:, 1]])) if rg2.n_dd else None)


if __name__ == "__main__":
    main()
