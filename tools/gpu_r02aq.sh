#!/bin/bash
# r02aq: ialspp_dense factorisation on packed FP32 (FFMA2 + 128-bit pivot-row loads): parity, timing, sharded test.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_dist.py -m gpu -q -k "ialspp or IALSPP or golden or two_ranks" > gpurun_out/t_aq.log 2>&1
echo "== ialspp tests rc=$?"; tail -n 3 gpurun_out/t_aq.log
timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 3 > gpurun_out/ialspp_c2_aq.log 2>&1
echo "== c2 IALSPP rc=$?"; tail -n 1 gpurun_out/ialspp_c2_aq.log | cut -c1-600
timeout 300 python tools/time_config.py --config c2 --solver IALSPP --subspace 32 --epochs 3 > gpurun_out/ialspp_c2_aq_s32.log 2>&1
echo "== c2 IALSPP S=32 rc=$?"; tail -n 1 gpurun_out/ialspp_c2_aq_s32.log | cut -c1-250
