#!/bin/bash
# r02au (2 GPUs): bench.py --gpus 2 (configs[3] strong scaling + configs[4] sample) and the NCCL test, after the
# recommend_impl / run_solver changes of the end of the round.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/scale2_au_bench.json 2> gpurun_out/scale2_au_bench.err
echo "== bench N=2 rc=$?"; tail -n 2 gpurun_out/scale2_au_bench.err | cut -c1-300; grep '^{' gpurun_out/scale2_au_bench.json | cut -c1-700
timeout 300 python -m pytest tests/test_dist.py -m gpu -q > gpurun_out/t_au.log 2>&1
echo "== dist tests rc=$?"; tail -n 3 gpurun_out/t_au.log
