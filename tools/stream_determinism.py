"""Development aid: run-to-run determinism of the streamed Schur kernel.  Solves the same batch repeatedly and compares
S | rhs, the E-buffer-dependent solution x and the final states bit for bit with the first run.
usage: SWGN_SCHUR_STREAM=1 python tools/stream_determinism.py [n_windows] [repeats]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import swgn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ws = bench.make_windows(n, 0, os.cpu_count() or 1)
opt = ws[0].options()
b = swgn.Batch([w.graph_p for w in ws], opt)
x0 = np.concatenate([w.state0() for w in ws])
D = np.random.default_rng(0).uniform(0.5, 1.5, 1412) * 1e-2
ref = None
bad = {"S": 0, "rhs": 0, "x": 0, "state": 0}
for r in range(reps):
    b.set_states(x0)
    x = b.linear_solve(0, D, 1412)
    S, rhs = b.get_reduced(0)
    b.set_states(x0)
    b.solve()
    st = b.get_states().copy()
    cur = (np.triu(S).copy(), rhs.copy(), x.copy(), st)
    if ref is None:
        ref = cur
        continue
    for k, name in enumerate(bad):
        if not np.array_equal(cur[k], ref[k]):
            bad[name] += 1
            if bad[name] <= 2:
                d = np.abs(cur[k] - ref[k])
                print("run", r, name, "differs: max abs", d.max(), "at", np.unravel_index(np.argmax(d), d.shape), "count", int((d > 0).sum()))
print("differences over", reps - 1, "repeats:", bad)
