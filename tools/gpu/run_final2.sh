#!/bin/bash
# end-of-round check: whole GPU suite, smoke, default bench line, launch list
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/final_tests.log
python bench.py 2> gpurun_out/bench_final.err | tail -1 > gpurun_out/r02_bench_4096win.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_4096win.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['serial_value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['share_of_step'], d['cfg4_ambiguity_fix'], d['clocks'])" | tee -a gpurun_out/final_tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 400 --csv --log-file gpurun_out/r02_launches_1024win.csv python bench.py --windows 1024 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu1.err
python tools/ncu_summary.py launches gpurun_out/r02_launches_1024win.csv gpurun_out/r02_launches_1024win.md; cat gpurun_out/r02_launches_1024win.md
