"""Latency of the literal drop-in case: ONE cfg2 window per ceres::Solve, i.e. swgn_batch_create (preprocess +
upload) + solve + read-back + destroy for a batch of one, next to the CPU restatement on one host thread.
Run on the GPU box: python tools/single_window_latency.py"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
import swgn  # noqa: E402

ws = bench.make_windows(8, 0, 8, 2)
opt = ws[0].options()
rows = []
for k, w in enumerate(ws):
    t0 = time.perf_counter()
    b = swgn.Batch([w.graph_p], opt)
    t1 = time.perf_counter()
    sm = b.solve()
    t2 = time.perf_counter()
    b.get_state(0, w.n_state)
    t3 = time.perf_counter()
    b.close()
    t4 = time.perf_counter()
    rows.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, sm[0].num_iterations, b.timing()[0] if False else 0.0))
    print("window %d: create %.2f solve %.2f get %.2f destroy %.2f ms (%d iterations)" % (
        k, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), sm[0].num_iterations), flush=True)
tot = np.array([sum(r[:4]) for r in rows[2:]])
print("GPU, one window per call (median of %d after 2 warm-up calls): %.2f ms" % (len(tot), 1e3 * np.median(tot)))
cpu = []
for w in ws[:4]:
    it, t = bench.cpu_leg([w], opt, 1)
    cpu.append(t)
print("CPU restatement, one window on one thread, minimiser only: %.2f ms" % (1e3 * np.median(cpu)))
