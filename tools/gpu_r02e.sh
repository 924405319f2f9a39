#!/bin/bash
# r02e: whole GPU suite (blocked Cholesky factorisation); the tensor-core Cholesky route on the
# many-rows case (plain, then under memcheck if it fails); configs[2] at 5 %; ncu of the fused
# scoring kernel.
mkdir -p gpurun_out
TAIL=8 tools/gpu_check.sh tests
IALS_CHOL=tc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cholesky_k256_many_rows or (half_steps and 256)" > gpurun_out/t_chol_tc.log 2>&1
rc=$?; echo "== chol tc tests rc=$rc"; tail -n 5 gpurun_out/t_chol_tc.log
if [ $rc -ne 0 ]; then
  IALS_CHOL=tc timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cholesky_k256_many_rows" > gpurun_out/sanitize_chol_tc.log 2>&1; echo "== memcheck chol tc rc=$?"
  grep -E "Invalid|ERROR SUMMARY|at 0x|at .*\+0x|by thread|Address|kernel" gpurun_out/sanitize_chol_tc.log | head -n 30
fi
for m in "" tc; do
  IALS_CHOL=$m timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_${m:-tile}.log 2>&1
  echo "== c3 x 0.05 [IALS_CHOL=$m] rc=$?"; tail -n 1 gpurun_out/c3_scaled_${m:-tile}.log | cut -c1-400
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -c 1 \
  -f -o gpurun_out/prof_score python tools/profile_epoch.py --epochs 1 --recommend 16384 > gpurun_out/ncu_score.log 2>&1; echo "== ncu score rc=$?"; tail -n 2 gpurun_out/ncu_score.log
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend.log | cut -c1-300
