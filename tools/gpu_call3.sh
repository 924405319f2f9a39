#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_score_tc.py -x -q > gpurun_out/t_score.log 2>&1; echo "rc=$?" >> gpurun_out/t_score.log
tail -n 30 gpurun_out/t_score.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_score_tc.py > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -n 15 gpurun_out/t_all.log
timeout 300 python tools/time_recommend.py --k 10 > gpurun_out/recommend.log 2>&1; echo "rc=$?" >> gpurun_out/recommend.log
tail -n 6 gpurun_out/recommend.log
tools/gpu_ab.sh "A=0" "IALS_HEAVY_THRESHOLD=3072" "IALS_HEAVY_THRESHOLD=4096" "IALS_HEAVY_JOB_LEN=4096" "IALS_HEAVY_THRESHOLD=3072 IALS_HEAVY_JOB_LEN=4096" "IALS_HEAVY_THRESHOLD=1536"
