#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gnss_epoch.py tests/test_gpu_parity.py -x -q -m gpu -k "gnss or marginal_prior" 2>&1 | tail -5 | tee gpurun_out/run_gnss.log
python tools/gnss_epoch_bench.py 4096 4 2>&1 | tail -3 | tee gpurun_out/gnss_epoch_bench.json
