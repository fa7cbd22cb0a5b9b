"""Throughput of swgn_gnss_preprocess (per-epoch GNSS linearisation): R receivers, one epoch each per
call, 20 satellites (18 usable), the default RTK configuration; host buffers in, host buffers out (the C ABI as a host
calls it).  The CPU figure beside it is the oracle restatement of GnssPreprocess on one host thread (test infrastructure,
here as the timed baseline only).  Usage: python tools/gnss_epoch_bench.py [receivers] [epochs]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gnss_scenario as S  # noqa: E402
import oracle_binding as ob  # noqa: E402
import swgn_gnss as G  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n_epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    cfg = G.default_config()
    n_sc = 16  # distinct scenarios, reused round-robin over the receivers
    scs = [S.Scenario(100 + s, cfg=cfg) for s in range(n_sc)]
    trackers = [G.Tracker(cfg) for _ in range(R)]
    outputs = [G.OutputBuffers(cap_keep=64, cap_n=80) for _ in range(R)]
    times = []
    dt = [np.zeros(G.NCLK) for _ in range(n_sc)]
    for k in range(n_epochs):
        base = [sc.epoch(k) for sc in scs]
        epochs, frames, keep = [], [], []
        for r in range(R):
            e, obs, f = base[r % n_sc]
            e2, o2 = S.copy_epoch(e, obs)
            f2 = S.copy_frame(f)
            for c in range(G.NCLK):
                f2.gnss_dt[c] = dt[r % n_sc][c]
            epochs.append(e2), frames.append(f2), keep.append(o2)
        t0 = time.perf_counter()
        G.preprocess(trackers, epochs, frames, outputs)
        times.append(time.perf_counter() - t0)
        for s in range(n_sc):
            dt[s] = np.array(frames[s].gnss_dt[:])
    # CPU restatement, one thread, a bounded sample
    n_cpu = min(R, 64)
    To = [ob.OracleGnssTracker(cfg) for _ in range(n_cpu)]
    scs = [S.Scenario(100 + s, cfg=cfg) for s in range(n_sc)]
    cpu = []
    for k in range(n_epochs):
        base = [sc.epoch(k) for sc in scs]
        t = 0.0
        for r in range(n_cpu):
            e, obs, f = base[r % n_sc]
            e2, o2 = S.copy_epoch(e, obs)
            f2 = S.copy_frame(f)
            t0 = time.perf_counter()
            To[r].preprocess(e2, f2, outputs[r])
            t += time.perf_counter() - t0
        cpu.append(t / n_cpu)
    best = min(times[1:]) if len(times) > 1 else times[0]
    print(json.dumps({"metric": "gnss epochs preprocessed per second", "receivers": R, "epochs_timed": n_epochs,
                      "ms_per_call": [round(1e3 * t, 2) for t in times], "epochs_per_s": round(R / best, 1),
                      "cpu_oracle_ms_per_epoch": [round(1e3 * t, 3) for t in cpu], "cpu_epochs_per_s_one_thread": round(1.0 / min(cpu), 1),
                      "note": "first call includes module load / context creation; the python list marshalling of the ctypes arrays is inside the timed call"}))


if __name__ == "__main__":
    main()
