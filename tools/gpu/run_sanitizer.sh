#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gnss_epoch.py tests/test_gpu_parity.py -q -m gpu -x -k "gnss or marginal or marginalize or fixed_integer" 2>&1 | tail -12 | tee gpurun_out/sanitizer_r02_epoch.log
