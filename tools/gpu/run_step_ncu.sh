#!/bin/bash
# ncu --set full of one k_step launch and one k_backsub launch at 1024 windows + per-source-line stall samples
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 3 -c 1 -f -o gpurun_out/k_step_1024 python bench.py --windows 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/step_ncu.log 2>&1
python tools/ncu_summary.py kernel gpurun_out/k_step_1024.ncu-rep gpurun_out/k_step_1024.md
cat gpurun_out/k_step_1024.md
ncu -i gpurun_out/k_step_1024.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/k_step_1024_source.csv 2>/dev/null
python tools/ncu_hot_lines.py gpurun_out/k_step_1024_source.csv 14 | tee gpurun_out/k_step_hot_lines.txt
