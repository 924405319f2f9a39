#!/bin/bash
# r02b: the one-touch light-row kernel (cg_tile.cu) -- parity, A/B against cg_rows, ncu; and a
# compute-sanitizer run of the K = 256 tensor-core Cholesky route that crashed in r02a.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q \
  -k "not evaluator_ndcg" > gpurun_out/t_tile.log 2>&1; echo "== parity (tile) rc=$?"; tail -n 6 gpurun_out/t_tile.log
IALS_TILE_WARPS=16 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "half_steps or team or staged or empty_rows or overfit_cg or c1_config" > gpurun_out/t_tile16.log 2>&1; echo "== parity (tile16) rc=$?"; tail -n 4 gpurun_out/t_tile16.log
tools/gpu_ab.sh "A=0" "IALS_TILE_WARPS=16" "IALS_LIGHT=rows" "IALS_HEAVY_THRESHOLD=4096" "IALS_HEAVY_THRESHOLD=1024" \
  "IALS_TILE_WARPS=16 IALS_HEAVY_THRESHOLD=4096"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cg_tile_kernel -s 2 -c 2 \
  -f -o gpurun_out/prof_tile python tools/profile_epoch.py --epochs 2 > gpurun_out/ncu_tile.log 2>&1; echo "== ncu tile rc=$?"; tail -n 2 gpurun_out/ncu_tile.log
IALS_CHOL=tc timeout 600 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python tools/parity_chol256.py > gpurun_out/sanitize_chol_tc.log 2>&1; echo "== memcheck chol tc rc=$?"
grep -E "Invalid|ERROR SUMMARY|at 0x|by thread|in .*kernel|Address" gpurun_out/sanitize_chol_tc.log | head -n 24
