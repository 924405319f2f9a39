#!/bin/bash
# compute-sanitizer over a small slice of the parity suite (the reference has no
# sanitizer configuration; the ownership rule "a half-epoch reads the other factor matrix and
# writes disjoint rows of its own" is what memcheck / racecheck verify here).
#   gpurun --timeout 900 -- tools/gpu_sanitize.sh [memcheck|racecheck|synccheck|initcheck]
# tcgen05 / TMEM kernels are covered by memcheck and synccheck; racecheck only sees shared memory
# accessed through the generic proxy (the SIMT kernels: cg_rows, dense_cg, cholesky_tile, top-k).
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL='test_half_steps or test_gram or test_empty_rows_and_columns or test_user_scores_and_errors or test_topk_canonical_ties_and_minus_inf'
timeout 800 compute-sanitizer --tool "$TOOL" --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$TOOL.log 2>&1
echo "rc=$?" >> gpurun_out/sanitize_$TOOL.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/sanitize_$TOOL.log | tail -n 8
