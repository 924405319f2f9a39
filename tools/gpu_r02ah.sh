#!/bin/bash
# r02ah: allow-lists fused into score_tc (ials_trainer_recommend_allowed): parity + timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_score_tc.py tests/test_evaluator_wide.py -m gpu -q -x > gpurun_out/t_ah.log 2>&1
echo "== score/evaluator tests rc=$?"; tail -n 15 gpurun_out/t_ah.log
timeout 300 python tools/time_recommend.py > gpurun_out/recommend_ah.log 2>&1
echo "== time_recommend rc=$?"; cat gpurun_out/recommend_ah.log | cut -c1-300
