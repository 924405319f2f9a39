#!/bin/bash
# r02v: feature-aware iALS on the device (feature.cu + prior in cg.cu / cholesky_tile.cu).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_feature_aware.py -m gpu -q > gpurun_out/t_v.log 2>&1
echo "== feature tests rc=$?"; tail -n 40 gpurun_out/t_v.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_feature_aware.py -m gpu -x -q -k "warmup_errors or (epochs_match and 32 and True)" > gpurun_out/sanitize_v_memcheck.log 2>&1
echo "== memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_v_memcheck.log | head -n 8
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_feature_aware.py > gpurun_out/t_all.log 2>&1
echo "== all other gpu tests rc=$?"; tail -n 6 gpurun_out/t_all.log
