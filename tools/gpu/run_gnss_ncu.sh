#!/bin/bash
# launch list of swgn_gnss_preprocess (3 epochs of 4096 receivers)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/gnss_launches.csv python tools/gnss_epoch_bench.py 4096 3 > gpurun_out/gnss_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/gnss_launches.csv gpurun_out/gnss_launches.md; cat gpurun_out/gnss_launches.md
