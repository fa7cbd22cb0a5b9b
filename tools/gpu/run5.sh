cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# (1) launch list of a short bench (1024 windows)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 400 --csv --log-file gpurun_out/r02_launches_1024win.csv python bench.py --windows 1024 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu1.err
# (2) DRAM traffic of k_schur at the benched batch size
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_schur -s 9 -c 2 --csv --log-file gpurun_out/r02_k_schur_4096win.csv python bench.py --windows 4096 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu2.err
# (3) full capture of the streamed kernel, 592 windows
SWGN_SCHUR_STREAM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_schur_stream -s 9 -c 1 -o gpurun_out/r02_k_schur_stream_592win python bench.py --windows 592 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu3.err
# (4) sweep through bench.py
timeout 1200 python bench.py --sweep --steps 2 > gpurun_out/r02_sweep.json 2> gpurun_out/sweep.err
ls -la gpurun_out | tail -12; tail -2 gpurun_out/ncu1.err gpurun_out/ncu2.err gpurun_out/ncu3.err gpurun_out/sweep.err
