#!/bin/bash
# r02ab: the ll kernel with only the shortened pivot chain (committed state): timing.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cholesky or CHOLESKY" > gpurun_out/t_ab.log 2>&1
echo "== chol tests rc=$?"; tail -n 3 gpurun_out/t_ab.log
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_ab.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_ab.log | cut -c1-400
timeout 900 python tools/time_c3_sharded.py > gpurun_out/c3_sharded_1ab.log 2>&1
echo "== c3 sharded driver, 1 GPU rc=$?"; tail -n 1 gpurun_out/c3_sharded_1ab.log | cut -c1-400
