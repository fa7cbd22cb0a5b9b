#!/bin/bash
# final records: default bench line and the reference arm
mkdir -p gpurun_out
python bench.py 2> gpurun_out/bench_final.err | tail -1 > gpurun_out/r02_bench_4096win.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference_arm.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_4096win.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cfg4', d['cfg4_ambiguity_fix']['ms'], 'gnss', d.get('gnss_epoch_preprocess'), d['cpu_baseline'], d['clocks'])
r=json.load(open('gpurun_out/r02_bench_reference_arm.json'))
print('reference arm', r['value'], r['unit'], r['cpu_baseline'])"
