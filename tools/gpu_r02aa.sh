#!/bin/bash
# r02aa: ll kernel with per-warp staging pipelines (no block barrier in the accumulation loop) and the
# shortened pivot chain of the diagonal block; launch list of bench.py --c4 off (configs[1] only).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_wgram.py tests/test_oracle_vs_reference_trainer.py -m gpu -q -k "cholesky or CHOLESKY or gram_of_256 or reference or failures" > gpurun_out/t_aa.log 2>&1
echo "== chol tests rc=$?"; tail -n 5 gpurun_out/t_aa.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 0 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_half_steps and CHOLESKY and 256" > gpurun_out/sanitize_aa_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" gpurun_out/sanitize_aa_$tool.log | head -n 6
done
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_aa.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_aa.log | cut -c1-500
timeout 900 python tools/time_c3_sharded.py > gpurun_out/c3_sharded_1aa.log 2>&1
echo "== c3 sharded driver, 1 GPU rc=$?"; tail -n 1 gpurun_out/c3_sharded_1aa.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2only.csv \
  python bench.py --steps 2 --warmup 1 --c4 off --cpu-epochs 1 > gpurun_out/launches_c2only.log 2>&1
echo "== launch list (configs[1] only) rc=$?"
