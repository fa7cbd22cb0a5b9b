"""Development aid: time swgn_preintegrate_batch (device kernel + H2D/D2H through the C ABI) on the IMU load of
the BASELINE batch: 4096 windows x 29 factors x 101 samples (0.25 s at 400 Hz)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import swgn  # noqa: E402
import test_preintegration as tp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 29
begin, s, bias = tp.streams(n, 1, lo=101, hi=101)
swgn.preintegrate_batch(begin[:65], s[:begin[64]], bias[:64], tp.NOISE)
t0 = time.perf_counter()
rec, info = swgn.preintegrate_batch(begin, s, bias, tp.NOISE)
dt = time.perf_counter() - t0
print("factors %d samples %d failures %d: %.1f ms through the C ABI (%.2f M samples/s)" % (n, begin[-1], int(info.sum()), 1e3 * dt, begin[-1] / dt / 1e6))
if len(sys.argv) > 2:
    import oracle_binding as ob
    m = 2000
    t0 = time.perf_counter()
    ob.preintegrate_batch(begin[:m + 1], s[:begin[m]], bias[:m], tp.NOISE)
    dc = time.perf_counter() - t0
    print("CPU restatement, 1 thread: %.2f M samples/s" % (begin[m] / dc / 1e6))
