#!/bin/bash
mkdir -p gpurun_out
for nb in 16 32; do
echo "== NB $nb"
SWGN_CHOL_NB=$nb python bench.py --windows 4096 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'])"
done | tee gpurun_out/chol_exp.log
python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -2
