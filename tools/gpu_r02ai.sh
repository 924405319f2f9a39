#!/bin/bash
# r02ai: users picked by index (ials_trainer_recommend_users) + IDMapper fused serving: parity + timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_score_tc.py tests/test_id_mapping.py tests/test_evaluator_wide.py -m gpu -q -x > gpurun_out/t_ai.log 2>&1
echo "== score/id_mapping/evaluator tests rc=$?"; tail -n 15 gpurun_out/t_ai.log
timeout 300 python tools/time_recommend.py > gpurun_out/recommend_ai.log 2>&1
echo "== time_recommend rc=$?"; cat gpurun_out/recommend_ai.log | cut -c1-300
