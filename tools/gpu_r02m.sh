#!/bin/bash
# r02m: one-pass 256 x 256 Gram kernel (wgram256_kernel) in the Cholesky route; operator test,
# parity, memcheck, timings, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_wgram.py tests/test_gpu_parity.py tests/test_oracle_vs_reference_trainer.py -m gpu -x -q -k "gram or cholesky or CHOLESKY or reference or failures" > gpurun_out/t_m.log 2>&1
echo "== gpu tests rc=$?"; tail -n 12 gpurun_out/t_m.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py tests/test_wgram.py -m gpu -x -q -k "(test_half_steps and CHOLESKY and 256) or gram_of_256" > gpurun_out/sanitize_m_memcheck.log 2>&1
echo "== memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_m_memcheck.log | head -n 8
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_m.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_m.log | cut -c1-700
timeout 900 python tools/time_c3_sharded.py > gpurun_out/c3_sharded_1m.log 2>&1
echo "== c3 sharded driver, 1 GPU rc=$?"; tail -n 1 gpurun_out/c3_sharded_1m.log | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c3_m_launches.csv \
  python tools/time_config.py --config c3 --scale 0.05 --epochs 1 > gpurun_out/c3_m_launches.log 2>&1
echo "== launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:wgram256 -s 2 -c 1 -o gpurun_out/prof_wgram256 -f \
  python tools/time_config.py --config c3 --scale 0.05 --epochs 1 > gpurun_out/ncu_wgram256.log 2>&1
echo "== ncu rc=$?"
