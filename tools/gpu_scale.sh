#!/bin/bash
# Multi-GPU evidence on one box (gpurun --gpus N -- tools/gpu_scale.sh N):
#   1. the one-rank-per-device NCCL parity test (tests/test_dist.py),
#   2. bench.py --gpus N (weak scaling: N stacked ML-20M-shaped blocks),
#   3. tools/time_c4.py (BASELINE configs[3]: the 1 B-interaction matrix row-sharded across the
#      N GPUs, strong scaling; configs[4]: top-100 of its users) at SCALE (default 1.0).
# Every log lands in gpurun_out/ even when a step times out.
N=${1:-2}
SCALE=${SCALE:-1.0}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/scale${N}_topo.log 2>&1
timeout 420 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/scale${N}_dist_tests.log 2>&1
echo "rc=$?" >> gpurun_out/scale${N}_dist_tests.log; tail -n 6 gpurun_out/scale${N}_dist_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/scale${N}_bench.json 2> gpurun_out/scale${N}_bench.err
echo "rc=$?" >> gpurun_out/scale${N}_bench.err; tail -n 3 gpurun_out/scale${N}_bench.err; cat gpurun_out/scale${N}_bench.json
timeout 900 $TR --master-port 29519 tools/time_c4.py --scale $SCALE --epochs 3 \
  > gpurun_out/scale${N}_c4.json 2> gpurun_out/scale${N}_c4.err
echo "rc=$?" >> gpurun_out/scale${N}_c4.err; tail -n 3 gpurun_out/scale${N}_c4.err; cat gpurun_out/scale${N}_c4.json
