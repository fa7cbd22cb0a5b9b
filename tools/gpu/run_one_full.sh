#!/bin/bash
mkdir -p gpurun_out
python -m pytest "$@" > gpurun_out/run_one_full.log 2>&1
grep -n "Fatal\|Segmentation\|Aborted\|libref_estimator\|Current thread\|File \"/root/repo" gpurun_out/run_one_full.log | head -30
tail -5 gpurun_out/run_one_full.log
