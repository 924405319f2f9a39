#!/bin/bash
# r02w: (1) feature-aware tests incl. the comparison with the reference's own trainer; (2) timing
# experiment: how much of cg_rows_kernel is the sweep over P?  gpurun_exp_half.so = the library with
# half of the sweep skipped (wrong numbers, failure test off) -- only its phase times are read.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_feature_aware.py -m gpu -q > gpurun_out/t_w.log 2>&1
echo "== feature tests rc=$?"; tail -n 5 gpurun_out/t_w.log
timeout 300 python tools/time_config.py --config c2 --epochs 5 > gpurun_out/c2_full_sweep.log 2>&1
echo "== c2 (product library) rc=$?"; tail -n 1 gpurun_out/c2_full_sweep.log | cut -c1-700
cp irspack_b200/lib/libials_b200.so /tmp/lib_keep.so
cp gpurun_exp_half.so irspack_b200/lib/libials_b200.so
timeout 300 python tools/time_config.py --config c2 --epochs 5 > gpurun_out/c2_half_sweep.log 2>&1
echo "== c2 (half sweep experiment) rc=$?"; tail -n 1 gpurun_out/c2_half_sweep.log | cut -c1-700
cp /tmp/lib_keep.so irspack_b200/lib/libials_b200.so
