"""Development aid: per-phase clock64 accounting of k_schur_stream for every window of a batch (thread 0 of each CTA):
barrier-to-barrier time of the TMA wait, phase A (chunk products), B (factors, W), C (S terms) summed over the batches,
and the write-out of S.  usage: python tools/stream_timeline.py [n_windows]"""
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
t = out[:8 * n].reshape(n, 8).astype(float)
rows = [("batches total", t[:, 1] - t[:, 0]), ("  TMA wait", t[:, 2]), ("  A products", t[:, 3]), ("  B factors/W", t[:, 4]),
        ("  C S terms", t[:, 5]), ("write-out", t[:, 6] - t[:, 1]), ("total", t[:, 6] - t[:, 0])]
print("k_schur_stream per-CTA cycles: median / p90 / max")
for nm, v in rows:
    print("  %-16s %9.0f %9.0f %9.0f" % (nm, np.median(v), np.percentile(v, 90), v.max()))
pr = out[16 * n:].reshape(n, 8).astype(float)
print("  sample batches (cycles, median): C of batch 2 / 10 / n-3 / n-2: %.0f %.0f %.0f %.0f; A of 2 / 10: %.0f %.0f; B of 2 / 10: %.0f %.0f"
      % tuple(np.median(pr[:, k]) for k in range(8)))
print("SMs used:", len(set(t[:, 7])), " timing:", b.timing())
