#!/bin/bash
# r02k: cholesky_ll v2 (column-parallel panel, warp-0 diagonal block, 128 threads, 4 CTAs per SM) and
# the scoring epilogue without per-column votes: parity, sanitizer, timings, ncu of both kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 6 gpurun_out/t_all.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 0 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(test_half_steps and CHOLESKY and 256) or test_topk_canonical_ties_and_minus_inf" > gpurun_out/sanitize_k_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Race|Invalid|hazard" gpurun_out/sanitize_k_$tool.log | head -n 12
done
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_ll2.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_ll2.log | cut -c1-700
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend2.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend2.log | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cholesky_ll -s 2 -c 1 -o gpurun_out/prof_chol_ll2 -f \
  python tools/time_config.py --config c3 --scale 0.05 --epochs 1 > gpurun_out/ncu_chol_ll2.log 2>&1
echo "== ncu ll rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_tc -s 1 -c 1 -o gpurun_out/prof_score2 -f \
  python tools/time_recommend.py > gpurun_out/ncu_score2.log 2>&1
echo "== ncu score rc=$?"
timeout 900 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3_full_ll2.log 2>&1
echo "== c3 full rc=$?"; tail -n 1 gpurun_out/c3_full_ll2.log | cut -c1-700
