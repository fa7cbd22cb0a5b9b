#!/bin/bash
# ncu --set full of one k_eval (full mode) and one k_schur launch at 592 windows on the round's final kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eval -s 5 -c 1 -f -o gpurun_out/r02_k_eval_592 python bench.py --windows 592 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_schur -s 3 -c 1 -f -o gpurun_out/r02_k_schur_592 python bench.py --windows 592 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py kernel gpurun_out/r02_k_eval_592.ncu-rep gpurun_out/r02_k_eval_592win.md
python tools/ncu_summary.py kernel gpurun_out/r02_k_schur_592.ncu-rep gpurun_out/r02_k_schur_592win.md
head -12 gpurun_out/r02_k_eval_592win.md; head -12 gpurun_out/r02_k_schur_592win.md
