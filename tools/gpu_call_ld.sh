#!/bin/bash
# K <= 128 padded to the 128-float stride: GPU suite, C1 timing (default and tight stride), bench sanity.
mkdir -p gpurun_out
TAIL=8 tools/gpu_check.sh tests
timeout 200 python tools/time_config.py --config c1 > gpurun_out/c1_ld128.log 2>&1; echo "rc=$?" >> gpurun_out/c1_ld128.log; tail -n 2 gpurun_out/c1_ld128.log
IALS_LD_MIN=32 timeout 200 python tools/time_config.py --config c1 --cpu-epochs 0 > gpurun_out/c1_ld64.log 2>&1; echo "rc=$?" >> gpurun_out/c1_ld64.log; tail -n 2 gpurun_out/c1_ld64.log
tools/gpu_ab.sh "A=0"
