#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/tests_bench.log
python bench.py --windows 4096 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'],'e2e',d['e2e']['value'])" | tee -a gpurun_out/tests_bench.log
