#!/bin/bash
mkdir -p gpurun_out
SWGN_GNSS_DEBUG=1 python tools/gnss_epoch_bench.py 4096 4 2>&1 | tail -50 | tee gpurun_out/gnss_epoch_bench_dbg.log
