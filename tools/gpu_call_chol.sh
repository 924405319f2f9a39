#!/bin/bash
# Register-tiled Cholesky (cholesky_tile.cu): GPU suite, A/B against the v0 kernel on a scaled
# Netflix shape, then the full C3 epoch.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log; tail -n 12 gpurun_out/t_all.log
for mode in tile row; do
  IALS_CHOL=$mode timeout 200 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_$mode.log 2>&1
  echo "[$mode] rc=$?"; tail -n 1 gpurun_out/c3_scaled_$mode.log
done
IALS_CHOL=tile timeout 200 python tools/time_config.py --config c1 --cpu-epochs 0 > gpurun_out/c1_final.log 2>&1; tail -n 1 gpurun_out/c1_final.log
timeout 420 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3.log 2>&1; echo "rc=$?" >> gpurun_out/c3.log; tail -n 2 gpurun_out/c3.log
