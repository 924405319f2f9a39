#!/bin/bash
# r02o: scoring kernel with 8 epilogue warps / 4 producer warps; ll kernel with fixed staging roles.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "cholesky or CHOLESKY or topk or recommend or evaluator or score or reference or id_mapping" > gpurun_out/t_o.log 2>&1
echo "== gpu tests rc=$?"; tail -n 6 gpurun_out/t_o.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(test_half_steps and CHOLESKY and 256) or test_topk_canonical_ties_and_minus_inf or test_recommend_matches_reference_ordering" > gpurun_out/sanitize_o_memcheck.log 2>&1
echo "== memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitize_o_memcheck.log | head -n 8
for i in 1 2; do
timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_o$i.log 2>&1
echo "== c3 x 0.05 rc=$?"; tail -n 1 gpurun_out/c3_scaled_o$i.log | cut -c1-500
timeout 300 python tools/time_recommend.py > gpurun_out/time_recommend5_$i.log 2>&1; echo "== time_recommend rc=$?"; tail -n 4 gpurun_out/time_recommend5_$i.log | cut -c1-300
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_tc -s 1 -c 1 -o gpurun_out/prof_score5 -f \
  python tools/time_recommend.py > gpurun_out/ncu_score5.log 2>&1
echo "== ncu score rc=$?"
