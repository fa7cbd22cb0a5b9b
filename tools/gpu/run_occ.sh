#!/bin/bash
mkdir -p gpurun_out
for f in 0 77824 116736; do
  echo "== floor $f" 
  SWGN_SCHUR_SMEM_FLOOR=$f python bench.py --windows 4096 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'],'roofline',d['roofline'])"
done 2>&1 | tee gpurun_out/occ.log
for f in 77824 116736; do
SWGN_SCHUR_SMEM_FLOOR=$f timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_schur -s 9 -c 2 --csv --log-file gpurun_out/occ_$f.csv python bench.py --windows 4096 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
tail -7 gpurun_out/occ_$f.csv
done
