#!/bin/bash
mkdir -p gpurun_out
SWGN_DEBUG_TIMING=1 SWGN_GNSS_DEBUG=1 python tools/gnss_epoch_bench.py 4096 3 2>&1 | tail -42 | tee gpurun_out/gnss_epoch_bench_dbg2.log
