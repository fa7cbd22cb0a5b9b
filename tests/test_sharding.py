"""Multi-GPU path: windows are sharded over ranks with no data-path collective; the
only communication is the barrier and the max-over-ranks / sum reductions of the benchmark.  That
host-side logic is exercised here with world_size 2 on the gloo backend (CPU), with the CPU oracle
standing in for the per-rank solve (the CUDA library refuses to run without a device)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = textwrap.dedent('''
    import os, sys, ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "rtk-visual-inertial-navigation_b200")); sys.path.insert(0, ROOT)
    import numpy as np, torch, torch.distributed as dist
    import bench, swgn, oracle_binding as ob
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    W = 3
    # window ids [rank*W, (rank+1)*W): disjoint shards, same rule as bench.run_swgn
    ws = bench.make_windows(W, rank * W, 2)
    ids = torch.tensor([rank * W + i for i in range(W)])
    gathered = [torch.zeros_like(ids) for _ in range(world)]
    dist.all_gather(gathered, ids)
    all_ids = torch.cat(gathered).tolist()
    assert sorted(all_ids) == list(range(world * W)), all_ids          # every window exactly once
    its, t = bench.cpu_leg(ws, ws[0].options(), 2)                      # per-rank work, no collective inside
    tmax = torch.tensor([t], dtype=torch.float64); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tot = torch.tensor([float(its)], dtype=torch.float64); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    assert tot.item() == 8.0 * W * world and tmax.item() >= t - 1e-12
    # different ranks really got different windows
    chk = torch.tensor([float(np.sum(ws[0].state0()))], dtype=torch.float64)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    assert both[0].item() != both[1].item()
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


def test_world_size_2_gloo_sharding():
    port = free_port()
    code = "ROOT = %r\n" % ROOT + WORKER
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=280)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\\n%s" % (r, o)
        assert "rank %d ok" % r in o
