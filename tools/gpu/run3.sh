cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02_bench_4096win.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/r02_bench_4096win.json
