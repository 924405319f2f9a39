#!/bin/bash
# r02a: first call of round 2 on one B200 -- suite, bench, the gated variant checks, and the
# A/B lines that decide which kernel variants stay (tools/gpu_round2_first.sh, trimmed).
mkdir -p gpurun_out
TAIL=8 tools/gpu_check.sh tests bench
IALS_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_zz_experimental.py -m gpu -q -k "not ialspp" > gpurun_out/t_experimental.log 2>&1; tail -n 15 gpurun_out/t_experimental.log
tools/gpu_ab.sh "A=0" "IALS_WGRAM=kmajor" "IALS_WGRAM=kmajor IALS_HEAVY_THRESHOLD=1024" \
  "IALS_WGRAM=fused" "IALS_WGRAM=fused IALS_HEAVY_THRESHOLD=1024" "IALS_WGRAM=fused IALS_HEAVY_THRESHOLD=512" \
  "IALS_WGRAM=fused IALS_HEAVY_THRESHOLD=256" "IALS_ROWS_LDG=na" "IALS_ROWS_PER_WARP=4" "IALS_HEAVY_THRESHOLD=1024"
for v in "" kmajor; do
  IALS_WGRAM=$v timeout 300 python tools/time_wgram.py > gpurun_out/time_wgram_${v:-default}.log 2>&1
  echo "== time_wgram [IALS_WGRAM=$v] rc=$?"; tail -n 6 gpurun_out/time_wgram_${v:-default}.log
done
for m in "" tc; do
  IALS_CHOL=$m timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_${m:-tile}.log 2>&1
  echo "== c3 x 0.05 [IALS_CHOL=$m] rc=$?"; tail -n 1 gpurun_out/c3_scaled_${m:-tile}.log
done
IALS_CHOL=tc timeout 480 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3_tc.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/c3_tc.log
