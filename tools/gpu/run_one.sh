#!/bin/bash
# usage: tools/gpu/run_one.sh <pytest args...>   -- runs on the GPU box, log into gpurun_out/
mkdir -p gpurun_out
python -m pytest "$@" 2>&1 | tail -60 | tee gpurun_out/run_one.log
