#!/bin/bash
# r02h: tensor-core Cholesky route after the cross-kernel fix (groups = stages): parity cases,
# configs[2] at 5 % and at full size next to the SIMT route; scoring after the split change.
mkdir -p gpurun_out
IALS_CHOL=tc timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_wgram.py -m gpu -q -k "cholesky_k256 or (half_steps and 256) or gram_of_256" > gpurun_out/t_chol_tc.log 2>&1
echo "== chol tc tests rc=$?"; tail -n 4 gpurun_out/t_chol_tc.log
for m in tc ""; do
  IALS_CHOL=$m timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_${m:-tile}.log 2>&1
  echo "== c3 x 0.05 [IALS_CHOL=$m] rc=$?"; tail -n 1 gpurun_out/c3_scaled_${m:-tile}.log | cut -c1-400
done
if grep -q '^{' gpurun_out/c3_scaled_tc.log; then
  IALS_CHOL=tc timeout 900 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3_full_tc.log 2>&1; echo "== c3 full tc rc=$?"; tail -n 1 gpurun_out/c3_full_tc.log | cut -c1-400
fi
timeout 900 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3_full_tile.log 2>&1; echo "== c3 full tile rc=$?"; tail -n 1 gpurun_out/c3_full_tile.log | cut -c1-400
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend.log | cut -c1-300
