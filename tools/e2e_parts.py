"""Where the end-to-end time of one bench step goes (update_inputs / solve / get_states), and the same work
double-buffered over two half batches driven by two host threads, the second started half a period late so
that one half's copies fall into the other half's solve.  Run on the GPU box: python tools/e2e_parts.py"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
import swgn  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = 6
ws = bench.make_windows(W, 0, os.cpu_count() or 1, 2)
opt = ws[0].options()
b = swgn.Batch([w.graph_p for w in ws], opt)
sms = (swgn.Summary * W)()
out = np.zeros(b.states_size())
for k in range(4):
    t0 = time.perf_counter()
    b.update_inputs()
    t1 = time.perf_counter()
    b.solve(sms)
    t2 = time.perf_counter()
    b.get_states(out)
    t3 = time.perf_counter()
    print("serial: update %.1f ms  solve %.1f ms (device %.1f)  get_states %.1f ms  -> %.0f it/s" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), b.timing()[0], 1e3 * (t3 - t2),
        sum(sms[i].num_iterations for i in range(W)) / (t3 - t0)), flush=True)
b.close()

half = W // 2
parts = [swgn.Batch([w.graph_p for w in ws[:half]], opt), swgn.Batch([w.graph_p for w in ws[half:]], opt)]
psm = [(swgn.Summary * p.n)() for p in parts]
pout = [np.zeros(p.states_size()) for p in parts]
for stagger in (False, True):
    go = threading.Barrier(3)
    first_update = threading.Event()
    done = [0, 0]

    def drive(i):
        p = parts[i]
        p.update_inputs()
        p.solve(psm[i])
        go.wait()
        if stagger and i == 1:
            first_update.wait()
        for k in range(K):
            p.update_inputs()
            if i == 0 and k == 0:
                first_update.set()
            p.solve(psm[i])
            p.get_states(pout[i])
            done[i] += sum(psm[i][j].num_iterations for j in range(p.n))

    th = [threading.Thread(target=drive, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    go.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    print("two half batches, stagger=%s: %.1f ms per step of %d windows -> %.0f it/s" % (stagger, 1e3 * dt / K, W, sum(done) / dt), flush=True)
print("states equal to the serial leg:", np.allclose(np.concatenate(pout), out, rtol=1e-9, atol=1e-12))
