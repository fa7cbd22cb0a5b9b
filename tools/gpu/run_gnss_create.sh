#!/bin/bash
# phases of swgn_batch_create inside swgn_gnss_preprocess (SWGN_DEBUG_TIMING) -- 3 epochs of 4096 receivers
mkdir -p gpurun_out
SWGN_DEBUG_TIMING=1 SWGN_GNSS_DEBUG=1 python tools/gnss_epoch_bench.py 4096 3 > /dev/null 2> gpurun_out/gnss_create.err
grep -v "^+" gpurun_out/gnss_create.err | tail -60
