#!/bin/bash
# r02ap: ialspp_dense with the L2 prefetch of the next row: parity, timing, memcheck + racecheck of the new kernels
# (ialspp_dense, the allow-list / row-map forms of score_tc, gather_rows).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -k "ialspp or IALSPP or golden" > gpurun_out/t_ap.log 2>&1
echo "== ialspp tests rc=$?"; tail -n 3 gpurun_out/t_ap.log
timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 3 > gpurun_out/ialspp_c2_ap.log 2>&1
echo "== c2 IALSPP rc=$?"; tail -n 1 gpurun_out/ialspp_c2_ap.log | cut -c1-600
IALS_GS_CHUNK=1024 timeout 300 python tools/time_config.py --config c2 --solver IALSPP --epochs 3 > gpurun_out/ialspp_c2_ap1024.log 2>&1
echo "== c2 IALSPP chunk 1024 rc=$?"; tail -n 1 gpurun_out/ialspp_c2_ap1024.log | cut -c1-200
for TOOL in memcheck racecheck; do
  timeout 500 compute-sanitizer --tool $TOOL --error-exitcode 86 --launch-timeout 0 \
    python -m pytest tests/test_gpu_parity.py tests/test_score_tc.py -m gpu -x -q \
    -k "test_ialspp_half_steps or test_ialspp_long_rows or test_recommend_users_picked_by_index or test_fused_topk_allow_lists_exact" > gpurun_out/sanitize_ap_$TOOL.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_ap_$TOOL.log
  echo "== $TOOL"; grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/sanitize_ap_$TOOL.log | tail -n 4
done
