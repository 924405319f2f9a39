#!/bin/bash
# r02l: ll kernel with rotated warp roles; scoring epilogue on 16-column chunks with the select
# tree; configs[2] through the sharded driver on one GPU (device-built matrix) with the CPU port beside it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cholesky or CHOLESKY or topk or recommend or evaluator or score or reference" > gpurun_out/t_l.log 2>&1
echo "== gpu tests rc=$?"; tail -n 4 gpurun_out/t_l.log
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_ll3.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_ll3.log | cut -c1-700
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend3.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend3.log | cut -c1-300
timeout 900 python tools/time_c3_sharded.py --cpu-sample 0.02 > gpurun_out/c3_sharded_1.log 2>&1
echo "== c3 sharded driver, 1 GPU rc=$?"; tail -n 1 gpurun_out/c3_sharded_1.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c3_ll3_launches.csv \
  python tools/time_config.py --config c3 --scale 0.05 --epochs 1 > gpurun_out/c3_ll3_launches.log 2>&1
echo "== launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_tc -s 1 -c 1 -o gpurun_out/prof_score3 -f \
  python tools/time_recommend.py > gpurun_out/ncu_score3.log 2>&1
echo "== ncu score rc=$?"
