#!/bin/bash
# configs[2] (Netflix shape, K = 256, Cholesky) row-sharded over N GPUs of one box, and optionally
# (BENCH=1) bench.py --gpus N = configs[3] / [4]:   gpurun --gpus N -- tools/gpu_scale_c3.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29519 tools/time_c3_sharded.py --steps 3 --warmup 1 \
  > gpurun_out/c3_sharded_$N.log 2> gpurun_out/c3_sharded_$N.err
echo "rc=$?" >> gpurun_out/c3_sharded_$N.err; tail -n 2 gpurun_out/c3_sharded_$N.err; grep '^{' gpurun_out/c3_sharded_$N.log | cut -c1-900
if [ "${BENCH:-0}" = "1" ]; then
  timeout 900 $TR --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/scale${N}_bench.json 2> gpurun_out/scale${N}_bench.err
  echo "rc=$?" >> gpurun_out/scale${N}_bench.err; tail -n 2 gpurun_out/scale${N}_bench.err; grep '^{' gpurun_out/scale${N}_bench.json | cut -c1-600
fi
