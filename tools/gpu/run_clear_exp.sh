#!/bin/bash
for v in 0 1; do
if [ $v = 1 ]; then export SWGN_SCHUR_SKIP_CLEAR=1; fi
python bench.py --windows 4096 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('skip=$v value',d['value'],'ms_per_step',d['ms_per_step'],'schur ms',d['roofline']['avg_launch_ms'], 'failed', d['config']['failed_windows'])"
done
