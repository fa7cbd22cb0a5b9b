import sys, time, os
sys.path.insert(0, 'rtk-visual-inertial-navigation_b200'); sys.path.insert(0, '.')
import numpy as np, swgn, bench
ws = bench.make_windows(512, 0, 16, 2)
opt = ws[0].options()
for k in range(3):
    t0 = time.perf_counter(); b = swgn.Batch([w.graph_p for w in ws], opt); t1 = time.perf_counter(); sm = b.solve(); t2 = time.perf_counter(); b.get_states(); t3 = time.perf_counter(); b.close(); t4 = time.perf_counter()
    print("create %.1f solve %.1f get %.1f close %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3)), flush=True)
