"""Cold path of one batch (512 cfg2 windows): swgn_batch_create (with SWGN_DEBUG_TIMING=1 the library prints its
plan / alloc / pack + upload phases), solve, read-back, destroy -- three times, the first one pays the CUDA
context and the pinned staging pool.  Run on the GPU box: SWGN_DEBUG_TIMING=1 python tools/cold_parts.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import swgn  # noqa: E402

ws = bench.make_windows(512, 0, os.cpu_count() or 1, 2)
opt = ws[0].options()
for k in range(3):
    t0 = time.perf_counter()
    b = swgn.Batch([w.graph_p for w in ws], opt)
    t1 = time.perf_counter()
    sm = b.solve()
    t2 = time.perf_counter()
    b.get_states()
    t3 = time.perf_counter()
    b.close()
    t4 = time.perf_counter()
    print("create %.1f solve %.1f get_states %.1f destroy %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)), flush=True)
