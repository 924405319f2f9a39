#!/bin/bash
# r02f: whole GPU suite (FFMA2 in the Cholesky tile kernel); the tensor-core Cholesky route on the
# many-rows and several-jobs-per-row cases (under memcheck where it fails); configs[2] at 5 %.
mkdir -p gpurun_out
TAIL=8 tools/gpu_check.sh tests
IALS_CHOL=tc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cholesky_k256 or (half_steps and 256)" > gpurun_out/t_chol_tc.log 2>&1
rc=$?; echo "== chol tc tests rc=$rc"; tail -n 6 gpurun_out/t_chol_tc.log
if [ $rc -ne 0 ]; then
  IALS_CHOL=tc timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cholesky_k256_rows_cut" > gpurun_out/sanitize_chol_tc.log 2>&1; echo "== memcheck chol tc rc=$?"
  grep -E "Invalid|ERROR SUMMARY|at 0x|at .*\+0x|by thread|Address|kernel|Saved host" gpurun_out/sanitize_chol_tc.log | head -n 40
fi
for m in "" tc; do
  IALS_CHOL=$m timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_${m:-tile}.log 2>&1
  echo "== c3 x 0.05 [IALS_CHOL=$m] rc=$?"; tail -n 1 gpurun_out/c3_scaled_${m:-tile}.log | cut -c1-400
done
