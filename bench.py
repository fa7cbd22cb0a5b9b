#!/usr/bin/env python
"""bench.py -- Gauss-Newton window-iterations/s on batched synthetic 20-KF x 300-landmark x
10-GNSS-epoch sliding windows (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--windows WPG] [--impl swgn|reference]

A "step" is one pass of the hot path over one batch: the full trust-region solve
(<= 8 DOGLEG iterations, DENSE_SCHUR) of every window of the batch through the C ABI
(swgn_batch_solve).  One process per GPU (torchrun for N > 1, weak scaling: every rank solves its
own `--windows` windows, no collective on the solve path); value = iterations executed by all
ranks / max-over-ranks device time (CUDA events on the library's stream).

--impl reference times the CPU path on the host cores: the reference's own modified-Ceres cannot
be built in this image (needs Eigen3/ROS/OpenCV), so the timed code is the oracle
restatement (oracle/, kind "port"), one window per OpenMP thread on all host cores.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200"))
sys.path.insert(0, ROOT)

METRIC = "gauss_newton_window_iterations_per_sec"
UNIT = "iterations/s"
WORKLOAD = "cfg3: batch of independent cfg2 windows (20 KF + 10 GNSS frames, 300 landmarks, 20 sats, fp64, DOGLEG<=8 it)"


WORKLOAD_A = ("cfg3-A: batch of independent composition-A windows (20 KF, 300 landmarks, 9 IMUGNSSFactor chains hiding 18 GNSS "
              "frames, 20 sats, fp64, DOGLEG<=8 it) -- secondary workload, not the BASELINE metric's configuration")


def make_windows(n, first_id, threads, which=2):
    import swgn
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        return list(ex.map(lambda i: swgn.SynthWindow(which, first_id + i), range(n)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_leg(windows, opt, threads):
    """Oracle (CPU restatement of the modified-Ceres path) on the host cores: one window per OpenMP
    thread, minimiser time only (the reference's own `minimizer_time_in_seconds` convention,
    RVI/swf/swf_core.cpp:409).  TEST INFRASTRUCTURE used as the measured baseline only."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    import swgn
    n = len(windows)
    arr = (C.POINTER(swgn.Graph) * n)(*[w.graph_p for w in windows])
    it = C.c_int64()
    t = ob.oracle().oracle_solve_batch_timed(n, arr, C.byref(opt), threads, C.byref(it), None, 0)
    return it.value, t


def gnss_epoch_leg(receivers=2048, n_epochs=4):
    """SURVEY 8f rank 4 next to the headline: swgn_gnss_preprocess for `receivers` receivers per call (20 satellites each,
    host buffers in and out), and the CPU restatement of GnssPreprocess on one thread on a bounded sample."""
    import ctypes as C
    import gnss_scenario as S
    import swgn_gnss as G
    cfg = G.default_config()
    n_sc = 16
    scs = [S.Scenario(100 + s, cfg=cfg) for s in range(n_sc)]
    trackers = [G.Tracker(cfg) for _ in range(receivers)]
    outputs = [G.OutputBuffers(cap_keep=64, cap_n=80) for _ in range(receivers)]
    dt = [np.zeros(G.NCLK) for _ in range(n_sc)]
    times = []
    for k in range(n_epochs):
        base = [sc.epoch(k) for sc in scs]
        epochs, frames, keep = [], [], []
        for r in range(receivers):
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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob  # TEST INFRASTRUCTURE, here as the timed CPU baseline only
    n_cpu = 32
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
    best = min(times[1:])
    return {"receivers_per_call": receivers, "satellites": 20, "ms_per_call": 1e3 * best, "epochs_per_s": receivers / best,
            "cpu_port_epochs_per_s_one_thread": 1.0 / min(cpu),
            "note": "GnssPreprocess (swf_gnss.cpp:265-587) for one epoch of every receiver per call through the C ABI, host buffers in "
                    "and out, ctypes marshalling included; the device solves are ~3 ms of a call, the rest is host planning / packing"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge.build_oracle()
    ge.build_synth()
    cores = os.cpu_count() or 1
    sample = max(cores, args.ref_windows)
    ws = make_windows(sample, 0, cores, 3 if args.composition == "A" else 2)
    opt = ws[0].options()
    for _ in range(args.warmup):
        cpu_leg(ws[:cores], opt, cores)
    its, tt = 0, 0.0
    for _ in range(args.steps):
        i, t = cpu_leg(ws, opt, cores)
        its += i
        tt += t
    v = its / tt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_A if args.composition == "A" else WORKLOAD, "windows_per_step": sample, "threads": cores},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d cfg2 windows per step, one window per OpenMP thread, minimiser time only "
                                   "(CPU restatement of the modified-Ceres path; the reference itself needs Eigen3/ROS/OpenCV)" % sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_swgn(args, rank, local_rank, world):
    import __graft_entry__ as ge
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ge.build_if_needed()  # a no-op when the in-tree libraries are fresh
    if dist is not None:
        dist.barrier()
    import swgn
    L = swgn.lib()
    if L.swgn_device_count() <= local_rank:
        raise RuntimeError("no CUDA device for rank %d: the solver has no CPU fallback" % rank)
    cores = os.cpu_count() or 1
    threads = max(1, cores // max(1, min(world, 8)))
    W = args.windows
    t0 = time.time()
    ws = make_windows(W, rank * W, threads, 3 if args.composition == "A" else 2)
    t_gen = time.time() - t0
    opt = ws[0].options()
    opt.device = local_rank
    t0 = time.perf_counter()
    b = swgn.Batch([w.graph_p for w in ws], opt)
    t_create = time.perf_counter() - t0
    x0 = np.concatenate([w.state0() for w in ws])
    sms = (swgn.Summary * W)()
    schur_bytes_w = np.array([b.schur_bytes(i) for i in range(W)], np.float64)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident leg: inputs already in HBM when the timed region starts
    # composition A: the chains' hidden states and history are part of the step's input, so the
    # untimed reset re-uploads the inputs instead of only the window states
    reset = (lambda: b.update_inputs()) if args.composition == "A" else (lambda: b.set_states(x0))
    for _ in range(args.warmup):
        reset()
        b.solve(sms)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms, schur_ms, schur_bytes, iters, launches, n_schur = 0.0, 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        reset()  # untimed: restores the initial point in HBM
        b.solve(sms)
        tot, sch, nsch, nk = b.timing()
        dev_ms += tot
        schur_ms += sch
        launches += nk
        n_schur += nsch
        nls = np.array([sms[i].num_linear_solves for i in range(W)], np.float64)
        schur_bytes += float((nls * schur_bytes_w).sum())
        iters += sum(sms[i].num_iterations for i in range(W))
    barrier()
    clocks = sampler.stop() if sampler else None
    final_costs = np.array([sms[i].final_cost for i in range(W)])
    init_costs = np.array([sms[i].initial_cost for i in range(W)])
    n_fail = sum(1 for i in range(W) if sms[i].termination_type == 2)

    # ---- end-to-end leg through the C ABI with host buffers: H2D of the step's inputs (factor constants + initial
    # states, pinned staging), solve, D2H of states and summaries.  (a) serial: update_inputs -> solve -> get_states;
    # (b) double-buffered, the headline: the NEXT step's inputs are packed and uploaded into a shadow block on a copy stream
    # by a second host thread (swgn_batch_prefetch_inputs) while the current step solves on the one compute stream, and
    # become live between two solves (swgn_batch_commit_inputs).  Every step's H2D and D2H are inside the timed region.
    import threading
    e2e_serial_s, e2e_s, e2e_iters, h2d = 0.0, 0.0, 0, 0
    out = np.zeros(b.states_size())
    for k in range(1 + max(1, args.steps // 2)):
        barrier()
        t0 = time.perf_counter()
        h2d = b.update_inputs()
        b.solve(sms)
        b.get_states(out)
        dt = time.perf_counter() - t0
        if k == 0:
            continue  # first call allocates the pinned staging buffer
        e2e_serial_s += dt
    e2e_serial_s /= max(1, args.steps // 2)
    b.prefetch_inputs()  # (allocates the shadow blocks; untimed)
    b.commit_inputs()
    b.prefetch_inputs()
    b.commit_inputs()
    b.prefetch_inputs()  # inputs of the first timed step, staged like every later step's: during the step before it
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        b.commit_inputs()
        t = threading.Thread(target=b.prefetch_inputs)  # the next step's inputs (ctypes releases the GIL)
        t.start()
        b.solve(sms)
        b.get_states(out)
        t.join()
        e2e_iters += sum(sms[i].num_iterations for i in range(W))
    e2e_s = time.perf_counter() - t0
    d2h = out.nbytes + W * 160  # states + TRState records
    # ---- cfg4 leg (BASELINE configs[3]): covariance recovery + LAMBDA fix decision of every window of the solved batch in
    # two launches (swgn_batch_ambiguity_fix); the bit-exact check against the oracle lives in tests/test_gpu_parity.py
    cfg4 = None
    if args.composition == "B" and ws[0].n_amb > 0:
        epochs = swgn.Batch.pack_epochs([w.ambiguity_epochs() for w in ws])
        b.ambiguity_fix_all(ws[0].n_amb, epochs)
        t_calls = []
        for _ in range(3):
            t0 = time.perf_counter()
            res, _, _ = b.ambiguity_fix_all(ws[0].n_amb, epochs)
            t_calls.append(time.perf_counter() - t0)
        t_fix = min(t_calls)
        cfg4 = {"windows": W, "n_ambiguities": int(ws[0].n_amb), "ms": 1e3 * t_fix, "ms_calls": [round(1e3 * t, 2) for t in t_calls], "windows_per_s": W / t_fix,
                "searched": sum(1 for i in range(W) if res[i].status == 0), "ratio_test_passed": sum(1 for i in range(W) if res[i].search_ok),
                "note": "wall time of the C-ABI call incl. H2D of the epoch lists and D2H of the results; best of three calls"}
    # ---- strong-scaling leg (BASELINE configs[2] as written: `--windows` windows in total, sharded over the ranks)
    strong = None
    if world > 1:
        Ws = max(1, W // world)
        bs = swgn.Batch([w.graph_p for w in ws[:Ws]], opt)
        xs0 = np.concatenate([w.state0() for w in ws[:Ws]])
        sms_s = (swgn.Summary * Ws)()
        s_ms, s_it = 0.0, 0
        for k in range(1 + args.steps):
            bs.set_states(xs0)
            bs.solve(sms_s)
            if k == 0:
                continue
            s_ms += bs.timing()[0]
            s_it += sum(sms_s[i].num_iterations for i in range(Ws))
        bs.close()
        strong = [s_ms, float(s_it), Ws]
    # cold path: planning + allocation + upload + solve + read-back of a fresh batch
    # (twice, the second one reported: the first grows the process-wide pinned staging pool to this batch size)
    for _ in range(2):
        t0 = time.perf_counter()
        b2 = swgn.Batch([w.graph_p for w in ws[:min(W, 512)]], opt)
        sm2 = b2.solve()
        b2.get_states()
        t_cold = time.perf_counter() - t0
        cold_iters = sum(sm2[i].num_iterations for i in range(b2.n))
        b2.close()

    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, e2e_s, e2e_serial_s, strong[0]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([float(iters), float(e2e_iters), float(launches), float(n_fail), strong[1]], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, e2e_serial_s, strong[0] = t.tolist()
        iters, e2e_iters, launches, n_fail, strong[1] = s.tolist()
    if rank != 0:
        b.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    value = iters / (dev_ms * 1e-3)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = schur_bytes / (schur_ms * 1e-3) / 1e9
    # DRAM traffic of the same kernel from the committed ncu capture (dram__bytes_read.sum + dram__bytes_write.sum of one
    # launch), scaled per window to this launch
    traffic, traffic_src = None, None
    try:
        cap = "r02_k_schur_4096win.md" if os.path.exists(os.path.join(ROOT, "profiles", "r02_k_schur_4096win.md")) else "r01b_k_schur_592win.md"
        cap_windows = 4096.0 if cap.startswith("r02") else 592.0
        txt = open(os.path.join(ROOT, "profiles", cap)).read()

        def grab(name):
            import re
            m = re.search(r"\| %s \| ([0-9.]+) \| (\w+) \|" % re.escape(name), txt)
            return float(m.group(1)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m.group(2)]
        per_window = (grab("dram__bytes_read.sum") + grab("dram__bytes_write.sum")) / cap_windows
        traffic = per_window * (schur_bytes / max(1, n_schur)) / float(np.mean(schur_bytes_w))
        traffic_src = "profiles/%s (ncu, %d-window launch), per window x windows of this launch" % (cap, int(cap_windows))
    except Exception:
        pass
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ncpu = min(W, max(cores, args.ref_windows))
        ci, ct = cpu_leg(ws[:ncpu], ws[0].options(), cores)
        # BASELINE.md 3 variants.  The port solves one window on ONE thread (the reference parallelises a single solve over 4
        # threads, RVI/swf/swf.cpp:29; the port does not): (i) one window at a time on one thread = the latency of one solve;
        # (i') four windows side by side on four threads = what a perfectly scaling 4-thread solve would reach; (iii) above
        c1i, c1t = cpu_leg(ws[:min(W, 4)], ws[0].options(), 1)
        c4i, c4t = cpu_leg(ws[:min(W, 16)], ws[0].options(), min(4, cores))
        cpu = {"value": ci / ct, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of the same cfg2 windows, one window per OpenMP thread on %d threads, minimiser time only "
                         "(CPU restatement of the modified-Ceres path)" % (ncpu, cores),
               "variants": {"one_window_at_a_time_1_thread": c1i / c1t, "four_windows_on_4_threads": c4i / c4t,
                            "one_window_per_core_%d_threads" % cores: ci / ct}}
    gnss = None
    if world == 1 and not args.no_cpu_baseline:
        gnss = gnss_epoch_leg()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_A if args.composition == "A" else WORKLOAD, "windows_per_gpu": W, "iterations_per_window": iters / (args.steps * W * world),
                   "n_f": int(sms[0].n_f), "n_e": int(sms[0].n_e), "n_residuals": int(sms[0].n_residuals),
                   "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed" % (2.9e-3 * W),
                   "failed_windows": int(n_fail), "median_cost_reduction": float(np.median(final_costs / init_costs)),
                   "host_generate_s": round(t_gen, 2), "host_plan_upload_s": round(t_create, 2)},
        "roofline": {"kernel": "k_schur", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                     "bytes_per_launch": schur_bytes / max(1, n_schur), "launches": n_schur, "avg_launch_ms": schur_ms / max(1, n_schur), "share_of_step": schur_ms / dev_ms},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_iters / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "mode": "double-buffered inputs: step k+1 is packed and uploaded (copy stream, second host thread) while step k solves",
                "serial_value": (e2e_iters / args.steps) / e2e_serial_s, "serial_ms_per_step": 1e3 * e2e_serial_s,
                "cold_value": cold_iters / t_cold, "cold_note": "fresh batch of %d windows in a warm process: host planning, device allocation, upload, solve, read-back" % min(W, 512)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "cfg4_ambiguity_fix": cfg4,
        "gnss_epoch_preprocess": gnss,
        "strong_scaling": None if strong is None else {
            "value": strong[1] / (strong[0] * 1e-3), "unit": UNIT, "windows_total": int(strong[2]) * world, "windows_per_gpu": int(strong[2]),
            "ms_per_step": strong[0] / args.steps, "note": "the same `--windows` windows in total, sharded over the ranks (device-resident, max over ranks)"},
    }
    print(json.dumps(line), flush=True)
    b.close()
    if dist is not None:
        dist.destroy_process_group()


def run_sweep(args):
    """BASELINE configs[4]: keyframes x landmarks sweep, GNSS epochs = KF / 2; per shape the Schur kernel's achieved
    algorithmic GB/s (CUDA events around its launches) next to the whole-solve throughput.  One JSON line."""
    import __graft_entry__ as ge
    ge.build_if_needed()
    import swgn
    peak = 6553.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []
    n = args.sweep_windows
    for kf in (10, 20, 40):
        for lm in (100, 300, 1000):
            with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
                ws = list(ex.map(lambda i: swgn.SynthWindow(2, i, n_keyframes=kf, n_landmarks=lm, n_gnss_epochs=kf // 2), range(n)))
            b = swgn.Batch([w.graph_p for w in ws], ws[0].options())
            x0 = np.concatenate([w.state0() for w in ws])
            tot = sch = nb = 0.0
            its = 0
            sm = None
            for rep in range(1 + args.steps):
                b.set_states(x0)
                sm = b.solve()
                if rep == 0:
                    continue
                t, s_, nl, nk = b.timing()
                tot += t
                sch += s_
                nls = np.array([sm[i].num_linear_solves for i in range(n)], np.float64)
                nb += float((nls * np.array([b.schur_bytes(i) for i in range(n)], np.float64)).sum())
                its += sum(sm[i].num_iterations for i in range(n))
            gbs = nb / (sch * 1e-3) / 1e9
            rows.append({"keyframes": kf, "landmarks": lm, "windows": n, "n_e": int(sm[0].n_e), "n_f": int(sm[0].n_f),
                         "schur_mb_per_window_iteration": b.schur_bytes(0) / 1e6, "k_schur_gbs": gbs, "frac": gbs / peak,
                         "iterations_per_s": its / (tot * 1e-3), "k_schur_share": sch / tot,
                         "failed": sum(1 for i in range(n) if sm[i].termination_type == 2)})
            b.close()
            del ws
    print(json.dumps({"metric": "k_schur algorithmic GB/s over the window-size sweep", "unit": "GB/s", "peak": peak, "sweep": rows}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--windows", type=int, default=4096, help="windows per GPU")
    ap.add_argument("--ref-windows", type=int, default=128, help="windows per CPU step / cpu_baseline sample")
    ap.add_argument("--impl", default="swgn", choices=["swgn", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--composition", default="B", choices=["A", "B"],
                    help="B (default): the BASELINE cfg2 window with explicit GNSS frames; A: GNSS frames hidden in IMUGNSSFactor chains")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4]: window-size sweep of the Schur kernel (one GPU), one JSON line")
    ap.add_argument("--sweep-windows", type=int, default=888)
    args = ap.parse_args()
    if args.sweep:
        return run_sweep(args)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_swgn(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
