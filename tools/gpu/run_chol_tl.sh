#!/bin/bash
mkdir -p gpurun_out
python tools/schur_timeline.py 1024 chol > gpurun_out/chol_tl.log 2>&1
tail -25 gpurun_out/chol_tl.log
