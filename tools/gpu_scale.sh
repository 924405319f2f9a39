#!/bin/bash
# Multi-GPU evidence on one box (gpurun --gpus N -- tools/gpu_scale.sh N):
#   1. the one-rank-per-device NCCL parity test (tests/test_dist.py: sharded == single process,
#      replicas identical, own-row upload pushed to the peers),
#   2. bench.py --gpus N: BASELINE configs[3] (1 B interactions, 10 M x 2 M) row-sharded across
#      the N GPUs, strong scaling, per-phase times, own-row e2e, configs[4] top-100 sample
#      (IALS_BENCH_C4_SCALE shrinks the shape for dry runs).
# Every log lands in gpurun_out/ even when a step times out.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/scale${N}_topo.log 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 420 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/scale${N}_dist_tests.log 2>&1
  echo "rc=$?" >> gpurun_out/scale${N}_dist_tests.log; tail -n 6 gpurun_out/scale${N}_dist_tests.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29517 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 \
  > gpurun_out/scale${N}_bench.json 2> gpurun_out/scale${N}_bench.err
echo "rc=$?" >> gpurun_out/scale${N}_bench.err; tail -n 3 gpurun_out/scale${N}_bench.err; cat gpurun_out/scale${N}_bench.json
