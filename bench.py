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
        os.environ["NCCL_DEBUG"] = "WARN"  # stdout carries exactly one JSON line (no NCCL version banner)
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

    # ---- end-to-end leg through the C ABI with host buffers: H2D of the step's inputs (factor
    # constants + initial states, pinned staging), solve, D2H of states and summaries
    e2e_s, e2e_iters, h2d = 0.0, 0, 0
    out = np.zeros(b.states_size())
    for k in range(1 + args.steps):
        barrier()
        t0 = time.perf_counter()
        h2d = b.update_inputs()
        b.solve(sms)
        b.get_states(out)
        dt = time.perf_counter() - t0
        if k == 0:
            continue  # first call allocates the pinned staging buffer
        e2e_s += dt
        e2e_iters += sum(sms[i].num_iterations for i in range(W))
    d2h = out.nbytes + W * 160  # states + TRState records
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
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([float(iters), float(e2e_iters), float(launches), float(n_fail)], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s = t.tolist()
        iters, e2e_iters, launches, n_fail = s.tolist()
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
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (592-window launch),
    # scaled per window to this launch: dram__bytes_read.sum + dram__bytes_write.sum
    traffic, traffic_src = None, None
    try:
        txt = open(os.path.join(ROOT, "profiles", "r01b_k_schur_592win.md")).read()

        def grab(name):
            import re
            m = re.search(r"\| %s \| ([0-9.]+) \| (\w+) \|" % re.escape(name), txt)
            return float(m.group(1)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m.group(2)]
        per_window = (grab("dram__bytes_read.sum") + grab("dram__bytes_write.sum")) / 592.0
        traffic = per_window * (schur_bytes / max(1, n_schur)) / float(np.mean(schur_bytes_w))
        traffic_src = "profiles/r01b_k_schur_592win.md (ncu --set full, 592-window launch), scaled per window"
    except Exception:
        pass
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ncpu = min(W, max(cores, args.ref_windows))
        ci, ct = cpu_leg(ws[:ncpu], ws[0].options(), cores)
        cpu = {"value": ci / ct, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of the same cfg2 windows, one window per OpenMP thread on %d threads, minimiser time only "
                         "(CPU restatement of the modified-Ceres path)" % (ncpu, cores)}
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
                "cold_value": cold_iters / t_cold, "cold_note": "fresh batch of %d windows in a warm process: host planning, device allocation, upload, solve, read-back" % min(W, 512)},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    b.close()
    if dist is not None:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_swgn(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
