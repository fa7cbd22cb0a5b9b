"""Window-size sweep (BASELINE.json configs[4]): keyframes x landmarks, GNSS epochs = KF / 2.
For every shape a batch of windows is solved on the device and the Schur kernel's achieved algorithmic GB/s
(algorithmic bytes of the eliminations that ran / time of the k_schur launches, CUDA events) is reported next to
the whole-solve throughput.  usage: python tools/schur_sweep.py [out.md]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)
import swgn  # noqa: E402
from concurrent.futures import ThreadPoolExecutor  # noqa: E402

peak = 6553.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for kf in (10, 20, 40):
    for lm in (100, 300, 1000):
        # enough windows that the batch's working set is far beyond the 126 MB L2
        n = 888  # two full waves of 3 CTAs x 148 SMs
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            ws = list(ex.map(lambda i: swgn.SynthWindow(2, i, n_keyframes=kf, n_landmarks=lm, n_gnss_epochs=kf // 2), range(n)))
        opt = ws[0].options()
        b = swgn.Batch([w.graph_p for w in ws], opt)
        x0 = np.concatenate([w.state0() for w in ws])
        sm = None
        tot = sch = 0.0
        nb = 0.0
        its = 0
        for rep in range(3):
            b.set_states(x0)
            sm = b.solve()
            if rep == 0:
                continue
            t, s, nl, nk = b.timing()
            tot += t
            sch += s
            nls = np.array([sm[i].num_linear_solves for i in range(n)], np.float64)
            nb += float((nls * np.array([b.schur_bytes(i) for i in range(n)], np.float64)).sum())
            its += sum(sm[i].num_iterations for i in range(n))
        gbs = nb / (sch * 1e-3) / 1e9
        rows.append((kf, lm, n, int(sm[0].n_e), int(sm[0].n_f), int(sm[0].n_residuals), b.schur_bytes(0) / 1e6, gbs, gbs / peak,
                     its / (tot * 1e-3), sch / tot, sum(1 for i in range(n) if sm[i].termination_type == 2)))
        b.close()
        del ws
out = ["| KF | landmarks | windows | n_e | n_f | residuals | Schur MB/window-it | k_schur GB/s | frac of %.0f GB/s | window-it/s | k_schur share | failed |" % peak,
       "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    out.append("| %d | %d | %d | %d | %d | %d | %.2f | %.0f | %.3f | %.0f | %.2f | %d |" % r)
txt = "\n".join(out) + "\n\n(one B200, fp64, DOGLEG <= 8 iterations, device-resident inputs, CUDA-event timing; first column block of window 0 shown)\n"
print(txt)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt)
