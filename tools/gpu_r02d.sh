#!/bin/bash
# r02d: whole GPU suite after the clean-up; configs[2] with the tensor-core rank updates at 5 %
# and at full size; configs[3] at FULL size on one GPU (heavy-row threshold 2048 vs 512); ncu of the
# fused scoring kernel; compute-sanitizer over the tensor-core paths.
mkdir -p gpurun_out
TAIL=8 tools/gpu_check.sh tests
for m in "" tc; do
  IALS_CHOL=$m timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_${m:-tile}.log 2>&1
  echo "== c3 x 0.05 [IALS_CHOL=$m] rc=$?"; tail -n 1 gpurun_out/c3_scaled_${m:-tile}.log | cut -c1-400
done
if grep -q '^{' gpurun_out/c3_scaled_tc.log; then
  IALS_CHOL=tc timeout 600 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3_tc.log 2>&1; echo "== c3 full tc rc=$?"; tail -n 1 gpurun_out/c3_tc.log | cut -c1-400
fi
for thr in 2048 512; do
  IALS_HEAVY_THRESHOLD=$thr timeout 900 python tools/time_c4.py --scale 1.0 --steps 2 --e2e-steps 1 --score-users 32768 > gpurun_out/c4_full_1gpu_thr$thr.log 2>&1
  echo "== c4 full, 1 GPU, thr=$thr rc=$?"
  python - gpurun_out/c4_full_1gpu_thr$thr.log <<'PY'
import json, sys
l=[x for x in open(sys.argv[1]) if x.startswith("{")]
if not l: print(open(sys.argv[1]).read()[-1200:])
else:
    d=json.loads(l[-1]); print("ms/epoch %.2f  G int/s %.3f  GB/s %.0f build %.1fs plan %.1fs"%(d["ms_per_epoch"], d["interactions_per_s"]/1e9, d["achieved_gbs"], d["build_s"], d["plan_s"]), {k: round(v,2) for k,v in d["phases_ms_max_over_ranks"].items()}, "e2e ms", d["e2e"]["ms_per_step"] if d["e2e"] else None, "topk users/s", round(d["score_topk"]["users_per_s"]) if d["score_topk"] else None, d["schedule"])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 \
  -f -o gpurun_out/prof_score python tools/profile_epoch.py --epochs 1 --recommend 16384 > gpurun_out/ncu_score.log 2>&1; echo "== ncu score rc=$?"; tail -n 2 gpurun_out/ncu_score.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py tests/test_wgram.py tests/test_score_tc.py -m gpu -x -q \
  -k "low_thresholds or gram_of_256 or test_gram or gathered_weighted or (half_steps and CHOLESKY-256) or (fused_topk_matches and 64-100)" \
  > gpurun_out/sanitize_memcheck.log 2>&1; echo "== memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize_memcheck.log | tail -n 8
