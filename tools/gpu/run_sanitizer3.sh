#!/bin/bash
# racecheck over a cfg2 window (GNSS factors, epoch-clock chunks), a composition-A window (k_chain) and the LM / jacobi-scaling path
mkdir -p gpurun_out
{
echo "== racecheck: cfg2 full solve"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_full_solve_matches_oracle and 2-0" 2>&1 | tail -4
echo "== racecheck: composition A through the ceres shim (k_chain)"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ceres_shim.py -q -m gpu -x -k "test_imu_gnss_factor_through_ceres_api and 4-0" 2>&1 | tail -4
echo "== racecheck: LEVENBERG_MARQUARDT + jacobi_scaling"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_levenberg_marquardt_solve_matches_oracle" 2>&1 | tail -4
echo "== racecheck: export pass + marginal priors + LAMBDA batch"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_marginal_priors_of_a_whole_batch or test_lambda_batch_bit_exact or test_batched_ambiguity_fix" 2>&1 | tail -4
} 2>&1 | tee gpurun_out/r02_sanitizer_racecheck.txt
