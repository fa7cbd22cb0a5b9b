#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gnss_epoch.py tests/test_gpu_parity.py -q -m gpu -x -k "gnss or marginal or epoch" 2>&1 | tail -3
python tools/gnss_epoch_bench.py 4096 8 2>/dev/null | tail -1 > gpurun_out/r02_gnss_epoch_bench.json
cat gpurun_out/r02_gnss_epoch_bench.json
