#!/bin/bash
# k_step occupancy experiment: rebuild k_tr.cu with SWGN_STEP_CTAS = 4 (64 regs), 5, 6, 8 and bench each
mkdir -p gpurun_out
P=rtk-visual-inertial-navigation_b200
python -c "import __graft_entry__ as g; g.build_if_needed()" 2>/dev/null
for n in 4 5 6 8; do
  nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I include -gencode arch=compute_100a,code=sm_100a -DSWGN_STEP_CTAS=$n -c $P/csrc/k_tr.cu -o $P/build/k_tr.cu.o
  nvcc -shared -o $P/libswgn.so $P/build/*.o -gencode arch=compute_100a,code=sm_100a
  echo "== SWGN_STEP_CTAS $n"
  python bench.py --windows 4096 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value'],'ms_per_step',d['ms_per_step'])"
done 2>&1 | tee gpurun_out/step_occ.log
