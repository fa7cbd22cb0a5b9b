#!/bin/bash
# memcheck over the whole solver parity file and the reference drop-in tests; racecheck + synccheck over one full solve of a
# small window (shared-memory hazards and barrier divergence of k_eval / k_schur / k_chol / k_backsub / k_step / k_end)
mkdir -p gpurun_out
{
echo "== memcheck: tests/test_gpu_parity.py tests/test_ceres_shim.py"
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_ceres_shim.py -q -m gpu -x -k "not full_size and not sweep" 2>&1 | tail -6
echo "== racecheck: full solve of a cfg1 window and a small cfg2-shaped window"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_full_solve_matches_oracle and 1-0" 2>&1 | tail -6
echo "== synccheck: same"
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_full_solve_matches_oracle and 1-0" 2>&1 | tail -6
} 2>&1 | tee gpurun_out/r02_sanitizer_solver.txt
