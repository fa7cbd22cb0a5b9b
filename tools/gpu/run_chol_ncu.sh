#!/bin/bash
# ncu --set full of one k_chol launch (592 windows = one full wave at 2 CTAs/SM) + per-source-line stall samples
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chol -s 6 -c 1 -f -o gpurun_out/k_chol_592 python bench.py --windows 592 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/chol_ncu.log 2>&1
ncu -i gpurun_out/k_chol_592.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/k_chol_592_source.csv 2>/dev/null
python tools/ncu_hot_lines.py gpurun_out/k_chol_592_source.csv 45 | tee gpurun_out/k_chol_hot_lines.txt
