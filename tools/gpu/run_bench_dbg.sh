#!/bin/bash
mkdir -p gpurun_out
SWGN_GNSS_DEBUG=1 SWGN_DEBUG_TIMING=1 python bench.py --steps 1 --warmup 3 2>&1 >/dev/null | grep "gnss\]\|swgn_batch_create" | tail -36 | tee gpurun_out/bench_dbg.log
