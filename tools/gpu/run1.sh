set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tools/gpu_stage_check.py 1 0 > gpurun_out/stage1.log 2>&1; tail -20 gpurun_out/stage1.log
timeout 600 python tools/gpu_stage_check.py 2 0 > gpurun_out/stage2.log 2>&1; tail -20 gpurun_out/stage2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest1.log 2>&1; tail -15 gpurun_out/pytest1.log
timeout 600 python bench.py --windows 1024 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err; cat gpurun_out/bench_stream.json
SWGN_SCHUR_STREAM=0 timeout 600 python bench.py --windows 1024 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_gather.json 2> gpurun_out/bench_gather.err; cat gpurun_out/bench_gather.json
