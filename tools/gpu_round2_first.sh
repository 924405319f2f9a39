#!/bin/bash
# First GPU call of a round on one B200: the parity suite, the bench line, the ncu launch list,
# then the A/B candidates that were never measured: cg_rows with four rows per
# sweep over P, the heavy-row threshold around the current one, and the shared-memory counters
# of the tensor-core Gram (the bandwidth model of DESIGN.md 8.2 predicts the LSU/shared pipe,
# not the tensor pipe, as its limit) next to its K-major variant (wgram_k.cu).
mkdir -p gpurun_out
TAIL=8 tools/gpu_check.sh tests bench launches
IALS_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_experimental.py -m gpu -q > gpurun_out/t_experimental.log 2>&1; tail -n 5 gpurun_out/t_experimental.log
tools/gpu_ab.sh "A=0" "IALS_WGRAM=kmajor" "IALS_WGRAM=kmajor IALS_HEAVY_THRESHOLD=1024" "IALS_WGRAM=kmajor IALS_HEAVY_THRESHOLD=512" \
  "IALS_WGRAM=fused" "IALS_WGRAM=fused IALS_HEAVY_THRESHOLD=1024" "IALS_WGRAM=fused IALS_HEAVY_THRESHOLD=512" \
  "IALS_WGRAM=fused IALS_HEAVY_THRESHOLD=256" "IALS_ROWS_LDG=na" "IALS_ROWS_PER_WARP=4" "IALS_ROWS_PER_WARP=4 IALS_HEAVY_THRESHOLD=1024" \
  "IALS_HEAVY_THRESHOLD=1024" "IALS_HEAVY_THRESHOLD=3072"
M=l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum
M=$M,smsp__inst_executed_pipe_lsu.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
M=$M,sm__inst_executed_pipe_uniform.sum,sm__cycles_elapsed.max,gpu__time_duration.sum
M=$M,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:wgram_kernel -s 4 -c 4 --csv \
  --log-file gpurun_out/wgram_counters.csv python tools/profile_epoch.py --epochs 2 > gpurun_out/wgram_counters.log 2>&1
echo "rc=$?" >> gpurun_out/wgram_counters.log; tail -n 3 gpurun_out/wgram_counters.log
for v in "" kmajor; do
  IALS_WGRAM=$v timeout 300 python tools/time_wgram.py > gpurun_out/time_wgram_${v:-default}.log 2>&1
  echo "== time_wgram [IALS_WGRAM=$v] rc=$?"; tail -n 6 gpurun_out/time_wgram_${v:-default}.log
done
# iALS++: the v0 block solver is paced by its heaviest rows (64-thread CTAs stage
# their neighbours 16 at a time per warp); wider CTAs are one environment switch away
for th in "" 128 256 512; do
  IALS_IALSPP_THREADS=$th timeout 300 python tools/time_config.py --config c2 --solver IALSPP --scale 0.25 --epochs 2 \
    > gpurun_out/ialspp_threads_${th:-default}.log 2>&1
  echo "== iALS++ [IALS_IALSPP_THREADS=$th] rc=$?"; tail -n 1 gpurun_out/ialspp_threads_${th:-default}.log
done
# configs[2] (Netflix shape, K = 256, Cholesky): the register-tiled kernel against the
# tensor-core rank updates (IALS_CHOL=tc), 5 % sample first, then the full shape
for m in "" tc; do
  IALS_CHOL=$m timeout 300 python tools/time_config.py --config c3 --scale 0.05 --epochs 2 > gpurun_out/c3_scaled_${m:-tile}.log 2>&1
  echo "== c3 x 0.05 [IALS_CHOL=$m] rc=$?"; tail -n 1 gpurun_out/c3_scaled_${m:-tile}.log
done
IALS_CHOL=tc timeout 480 python tools/time_config.py --config c3 --epochs 2 > gpurun_out/c3_tc.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/c3_tc.log
