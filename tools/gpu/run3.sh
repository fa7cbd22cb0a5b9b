cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/single_window_latency.py 2>&1 | tail -3
