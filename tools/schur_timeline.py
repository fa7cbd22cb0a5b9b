"""Development aid: per-phase clock64 timeline of k_schur for every window of a batch.
usage: SWGN_DEBUG_TIMELINE=1 python tools/schur_timeline.py [n_windows]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)
os.environ["SWGN_DEBUG_TIMELINE"] = "1"
import bench  # noqa: E402
import swgn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ws = bench.make_windows(n, 0, os.cpu_count() or 1)
b = swgn.Batch([w.graph_p for w in ws], ws[0].options())
for _ in range(2):
    b.set_states(np.concatenate([w.state0() for w in ws]))
    b.solve()
L = swgn.lib()
L.swgn_batch_debug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
out = np.zeros(24 * n, np.int64)
assert L.swgn_batch_debug_timeline(b.h, out.ctypes.data_as(C.POINTER(C.c_int64))) == 0
t = out[:8 * n].reshape(n, 8)
tc = out[8 * n:16 * n].reshape(n, 8)
names = ["P0 zero", "P1a chunks", "wait barrier", "P1b rows", "P2 gather", "tail barrier"]
d = np.diff(t[:, :7], axis=1).astype(float)
print("per-CTA phase durations [cycles]: median / p90 / max")
for i, nm in enumerate(names):
    print("  %-14s %9.0f %9.0f %9.0f" % (nm, np.median(d[:, i]), np.percentile(d[:, i], 90), d[:, i].max()))
if len(sys.argv) > 2 and sys.argv[2] == "chol":
    names = ["stage panel", "in-panel", "write-back", "trailing", "back-solve", "total", " potrf+sync", " trsm+sync"]
    print("k_chol per-CTA accumulated phase cycles: median / p90 / max")
    for i, nm in enumerate(names):
        print("  %-14s %9.0f %9.0f %9.0f" % (nm, np.median(tc[:, i]), np.percentile(tc[:, i], 90), tc[:, i].max()))
    sys.exit(0)
tot = (t[:, 6] - t[:, 0]).astype(float)
print("  %-14s %9.0f %9.0f %9.0f" % ("total", np.median(tot), np.percentile(tot, 90), tot.max()))
span = t[:, 6].max() - t[:, 0].min()
print("kernel span (cycles, all SMs share one clock domain approx):", span, " CTAs:", n, " SMs used:", len(set(t[:, 7])))
print("timing:", b.timing())
