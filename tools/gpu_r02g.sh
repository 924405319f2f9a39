#!/bin/bash
# r02g: find the smallest Netflix-shaped scale at which IALS_CHOL=tc dies and run it under memcheck;
# scoring kernel after the epilogue fast path.
mkdir -p gpurun_out
fail=""
for sc in 0.005 0.01 0.02 0.03; do
  IALS_CHOL=tc timeout 300 python tools/time_config.py --config c3 --scale $sc --epochs 1 > gpurun_out/c3_tc_$sc.log 2>&1
  rc=$?; echo "== c3 x $sc tc rc=$rc"; tail -n 1 gpurun_out/c3_tc_$sc.log | cut -c1-300
  if [ $rc -ne 0 ] && [ -z "$fail" ]; then fail=$sc; fi
done
if [ -n "$fail" ]; then
  IALS_CHOL=tc timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
    python tools/time_config.py --config c3 --scale $fail --epochs 1 > gpurun_out/sanitize_c3_tc.log 2>&1; echo "== memcheck c3 x $fail tc rc=$?"
  grep -E "Invalid|ERROR SUMMARY|at .*\+0x|by thread|Address|kernel|Saved host|Error" gpurun_out/sanitize_c3_tc.log | head -n 40
fi
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend.log | cut -c1-300
timeout 300 python -m pytest tests/test_score_tc.py tests/test_gpu_parity.py -m gpu -x -q -k "topk or recommend or evaluator or score" > gpurun_out/t_score.log 2>&1; echo "== score tests rc=$?"; tail -n 3 gpurun_out/t_score.log
