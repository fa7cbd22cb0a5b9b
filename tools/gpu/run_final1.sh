#!/bin/bash
mkdir -p gpurun_out
python bench.py 2> gpurun_out/bench_final.err | tail -1 > gpurun_out/r02_bench_4096win.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_4096win.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['serial_value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'], d['cfg4_ambiguity_fix'], d['clocks'])"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 400 --csv --log-file gpurun_out/r02_launches_1024win.csv python bench.py --windows 1024 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu1.err
python tools/ncu_summary.py launches gpurun_out/r02_launches_1024win.csv gpurun_out/r02_launches_1024win.md; cat gpurun_out/r02_launches_1024win.md
