#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cholesky_tile_kernel -c 2 -f -o gpurun_out/prof_chol_tile \
  python tools/profile_epoch.py --shape netflix --scale 0.02 --solver CHOLESKY --epochs 1 > gpurun_out/ncu_chol_tile.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_chol_tile.log; tail -n 3 gpurun_out/ncu_chol_tile.log
