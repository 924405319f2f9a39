#!/bin/bash
# One short pytest selection under a hard timeout:  tools/gpu_quick.sh <timeout_s> <pytest args...>
mkdir -p gpurun_out
t=$1; shift
timeout "$t" python -m pytest "$@" > gpurun_out/quick.log 2>&1
rc=$?
echo "rc=$rc" >> gpurun_out/quick.log
tail -n 40 gpurun_out/quick.log
exit 0
