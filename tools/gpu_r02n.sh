#!/bin/bash
# r02n: whole GPU suite on the one-pass 256-column Gram + ll kernel; configs[2] timings; scoring with
# 128-key candidate buffers.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 8 gpurun_out/t_all.log
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_n.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_n.log | cut -c1-700
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend4.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend4.log | cut -c1-300
timeout 900 python tools/time_c3_sharded.py --cpu-sample 0.02 > gpurun_out/c3_sharded_1n.log 2>&1
echo "== c3 sharded driver, 1 GPU rc=$?"; tail -n 1 gpurun_out/c3_sharded_1n.log | cut -c1-400
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholesky_ll -s 2 -c 1 -o gpurun_out/prof_chol_ll4 -f \
  python tools/time_config.py --config c3 --scale 0.05 --epochs 1 > gpurun_out/ncu_chol_ll4.log 2>&1
echo "== ncu rc=$?"
