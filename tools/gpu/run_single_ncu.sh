#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/r02_launches_single_window.csv python tools/single_window_latency.py > gpurun_out/single.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r02_launches_single_window.csv | tee gpurun_out/r02_launches_single_window.md
tail -4 gpurun_out/single.log
