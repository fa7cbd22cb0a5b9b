#!/bin/bash
mkdir -p gpurun_out
python bench.py 2> gpurun_out/bench_final.err | tail -1 > gpurun_out/r02_bench_4096win.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_4096win.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['gnss_epoch_preprocess'], d['cpu_baseline']['variants'])"
tail -3 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_arm.json; head -c 600 gpurun_out/r02_bench_reference_arm.json
