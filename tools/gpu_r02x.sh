#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/time_half_step.py > gpurun_out/half_full.log 2>&1; echo "== product rc=$?"; tail -n 3 gpurun_out/half_full.log
cp irspack_b200/lib/libials_b200.so /tmp/lib_keep.so
cp gpurun_exp_half.so irspack_b200/lib/libials_b200.so
timeout 300 python tools/time_half_step.py > gpurun_out/half_half.log 2>&1; echo "== half sweep rc=$?"; tail -n 3 gpurun_out/half_half.log
cp /tmp/lib_keep.so irspack_b200/lib/libials_b200.so
