#!/bin/bash
mkdir -p gpurun_out
SWGN_GNSS_DEBUG=1 python -m pytest tests/test_gnss_epoch.py -x -q -m gpu  2>&1 | grep -v "^E  \|^    " | tail -30 | tee gpurun_out/dbg_gnss.log
