#!/bin/bash
mkdir -p gpurun_out
python bench.py --windows 4096 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'],'e2e',d['e2e']['value'],'schur ms',d['roofline']['avg_launch_ms'])" | tee gpurun_out/bench_quick.log
