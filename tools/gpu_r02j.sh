#!/bin/bash
# r02j: left-looking Cholesky after the tensor-core Gram (cholesky_ll.cu), now the default route of
# 256-column factors: parity, memcheck + racecheck on the K = 256 half-steps, configs[2] at 5 % with a
# launch list, at full size, and one ncu --set full capture of the new kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_wgram.py tests/test_oracle_vs_reference_trainer.py -m gpu -q \
  -k "cholesky or CHOLESKY or gram_of_256 or reference or failures" > gpurun_out/t_chol_ll.log 2>&1
echo "== chol tests rc=$?"; tail -n 15 gpurun_out/t_chol_ll.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 0 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_half_steps and CHOLESKY and 256" > gpurun_out/sanitize_ll_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Race|Invalid|hazard" gpurun_out/sanitize_ll_$tool.log | head -n 12
done
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_ll.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_ll.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c3_ll_launches.csv \
  python tools/time_config.py --config c3 --scale 0.05 --epochs 1 > gpurun_out/c3_ll_launches.log 2>&1
echo "== launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholesky_ll -s 2 -c 1 -o gpurun_out/prof_chol_ll -f \
  python tools/time_config.py --config c3 --scale 0.05 --epochs 1 > gpurun_out/ncu_chol_ll.log 2>&1
echo "== ncu rc=$?"
timeout 900 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3_full_ll.log 2>&1
echo "== c3 full rc=$?"; tail -n 1 gpurun_out/c3_full_ll.log | cut -c1-700
