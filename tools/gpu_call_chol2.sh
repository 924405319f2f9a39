#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log; tail -n 6 gpurun_out/t_all.log
timeout 200 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_tile2.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/c3_scaled_tile2.log
