#!/bin/bash
# usage: tools/gpu/run_bench_n.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/r02_bench_${N}gpu.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'serial', d['e2e']['serial_value'], 'strong', d['strong_scaling'])"
tail -2 gpurun_out/bench_${N}gpu.err
