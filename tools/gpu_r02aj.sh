#!/bin/bash
# r02aj: iALS++ as block Gauss-Seidel on the tensor-core Gram (ialspp_dense.cu): parity + timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -k "ialspp or IALSPP or golden" > gpurun_out/t_aj.log 2>&1
echo "== ialspp tests rc=$?"; tail -n 12 gpurun_out/t_aj.log
for chunk in 1024 4096; do
  IALS_GS_CHUNK=$chunk timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 3 > gpurun_out/ialspp_c2_$chunk.log 2>&1
  echo "== c2 IALSPP chunk $chunk rc=$?"; tail -n 1 gpurun_out/ialspp_c2_$chunk.log | cut -c1-600
done
IALS_IALSPP=simt timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 2 --scale 0.25 > gpurun_out/ialspp_c2_simt_quarter.log 2>&1
echo "== c2 x 0.25 IALSPP simt rc=$?"; tail -n 1 gpurun_out/ialspp_c2_simt_quarter.log | cut -c1-400
timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 2 --scale 0.25 > gpurun_out/ialspp_c2_quarter.log 2>&1
echo "== c2 x 0.25 IALSPP tensor rc=$?"; tail -n 1 gpurun_out/ialspp_c2_quarter.log | cut -c1-400
