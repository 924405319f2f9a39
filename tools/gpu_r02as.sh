#!/bin/bash
# r02as: ialspp_dense with 8-unknown substitution steps (inverted 8x8 diagonal sub-blocks): parity, phases, timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "ialspp or IALSPP or golden" > gpurun_out/t_as.log 2>&1
echo "== ialspp tests rc=$?"; tail -n 6 gpurun_out/t_as.log
timeout 300 python tools/_dbg_phase.py 2>&1 | tail -2
timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 3 > gpurun_out/ialspp_c2_as.log 2>&1
echo "== c2 IALSPP rc=$?"; tail -n 1 gpurun_out/ialspp_c2_as.log | cut -c1-600
