#!/bin/bash
# One GPU call for the hot-column cache of cg_rows.cu and the serving top-k (f3):
# full GPU suite, then A/B bench runs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "rc=$?" >> gpurun_out/t_all.log; tail -n 25 gpurun_out/t_all.log
tools/gpu_ab.sh "A=0" "IALS_HOT_SLOTS=0" "IALS_HOT_MIN_COVERAGE=0" "IALS_ROWS_PER_WARP=1" \
  "IALS_HEAVY_THRESHOLD=4096" "IALS_HOT_SLOTS=128"
