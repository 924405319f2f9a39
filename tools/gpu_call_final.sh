#!/bin/bash
# Round-end evidence on one B200: GPU suite, bench line (with the CPU baseline), ncu launch list,
# then the other BASELINE configurations that fit one GPU (C1, C3).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
TAIL=8 tools/gpu_check.sh tests bench launches
timeout 200 python tools/time_config.py --config c1 > gpurun_out/c1.log 2>&1; echo "rc=$?" >> gpurun_out/c1.log; tail -n 3 gpurun_out/c1.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "rc=$?" >> gpurun_out/bench_reference.log; tail -n 2 gpurun_out/bench_reference.log
timeout 480 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3.log 2>&1; echo "rc=$?" >> gpurun_out/c3.log; tail -n 3 gpurun_out/c3.log
