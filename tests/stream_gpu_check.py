"""Run by tests/test_gpu_stream.py in a subprocess with SWGN_SCHUR_STREAM=1: the streamed Schur kernel (k_schur_stream)
against the CPU oracle through the C ABI -- reduced system of one linear solve, and the full trust-region solve."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_binding as ob  # noqa: E402
import swgn  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    assert os.environ.get("SWGN_SCHUR_STREAM") == "1"
    L = swgn.lib()
    L.swgn_plan_stream_info.argtypes = [C.POINTER(swgn.Graph), C.c_int32, C.POINTER(C.c_int32)]
    for which, wid, kw in [(1, 0, {}), (2, 0, {}), (2, 3, {}), (2, 1, dict(n_keyframes=10, n_landmarks=100, n_gnss_epochs=5))]:
        w = swgn.SynthWindow(which, wid, **kw)
        info = (C.c_int32 * 16)()
        assert L.swgn_plan_stream_info(w.graph_p, 0, info) == 0
        assert info[10] == 1, "window is not routed to the streamed kernel: %r" % (list(info),)
        opt = w.options()
        o = ob.OracleSolver(w.graph_p, opt)
        b = swgn.Batch([w.graph_p], opt)
        rng = np.random.default_rng(wid)
        D = rng.uniform(0.5, 1.5, o.n_cols) * 1e-2
        x = b.linear_solve(0, D, o.n_cols)
        st, ox, oS, orhs = o.linear_solve(D)
        assert st == 0
        S, rhs = b.get_reduced(0)
        assert rel(np.triu(S), np.triu(oS)) < 1e-12, rel(np.triu(S), np.triu(oS))
        assert rel(rhs, orhs) < 1e-9
        assert rel(x, ox) < 1e-6
        sm = b.solve()[0]
        st, osm = o.minimize()
        assert sm.num_iterations == osm.num_iterations and sm.termination_type == osm.termination_type
        assert sm.num_linear_solves == osm.num_linear_solves
        assert abs(sm.final_cost - osm.final_cost) <= 1e-6 * osm.final_cost
        xs, xo = b.get_state(0, w.n_state), o.state()
        assert float(np.max(np.abs(xs - xo) / np.maximum(1.0, np.abs(xo)))) < 1e-6
        b.close()
        print("stream ok", which, wid, kw)
    # a batch mixing window shapes, and a window beyond the on-chip budget next to streamed ones (both kernels launch)
    ws = [swgn.SynthWindow(2, 5), swgn.SynthWindow(2, 0, n_keyframes=40, n_landmarks=300, n_gnss_epochs=20), swgn.SynthWindow(1, 1)]
    opt = ws[0].options()
    opt.n_parameter_head = 0  # (the small VI-only window has fewer retained blocks than the cfg2 head)
    b = swgn.Batch([w.graph_p for w in ws], opt)
    sms = b.solve()
    for i, w in enumerate(ws):
        o = ob.OracleSolver(w.graph_p, opt)
        st, osm = o.minimize()
        xs, xo = b.get_state(i, w.n_state), o.state()
        err = float(np.max(np.abs(xs - xo) / np.maximum(1.0, np.abs(xo))))
        print("mixed batch window", i, "iterations", sms[i].num_iterations, osm.num_iterations, "cost", sms[i].final_cost, osm.final_cost, "err", err)
        assert sms[i].num_iterations == osm.num_iterations
        assert err < 1e-6
    b.close()
    print("stream mixed batch ok")


if __name__ == "__main__":
    main()
