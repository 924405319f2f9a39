#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/diag_tiny.py > gpurun_out/diag_tiny.log 2>&1; echo "rc=$?" >> gpurun_out/diag_tiny.log; cat gpurun_out/diag_tiny.log | tail -60
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log; tail -n 15 gpurun_out/t_all.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cholesky_row_kernel -c 2 -f -o gpurun_out/prof_chol \
  python tools/profile_epoch.py --shape netflix --scale 0.02 --solver CHOLESKY --epochs 1 > gpurun_out/ncu_chol.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_chol.log; tail -n 3 gpurun_out/ncu_chol.log
