#!/bin/bash
# per-epoch GNSS path: tests of every caller of the batched prior read-back, then the throughput record
mkdir -p gpurun_out
python -m pytest tests/test_gnss_epoch.py tests/test_gpu_parity.py tests/test_ceres_shim.py -q -m gpu -x -k "gnss or marginal or epoch or oldest_frame or fixed_integer" 2>&1 | tail -3
SWGN_DEBUG_TIMING=1 SWGN_GNSS_DEBUG=1 python tools/gnss_epoch_bench.py 4096 3 2>&1 >/dev/null | grep -v "^+" | tail -24 | grep "priors\|create  \|read-back\|scatter"
python tools/gnss_epoch_bench.py 4096 8 2>/dev/null | tail -1 > gpurun_out/r02_gnss_epoch_bench.json
cat gpurun_out/r02_gnss_epoch_bench.json
