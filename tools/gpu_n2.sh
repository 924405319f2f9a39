#!/bin/bash
# Two-GPU check (gpurun --gpus 2 -- tools/gpu_n2.sh): real one-rank-per-device NCCL parity
# test, then the weak-scaling bench at N=2.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n2_topo.log 2>&1
timeout 420 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/n2_dist_tests.log 2>&1; echo "rc=$?" >> gpurun_out/n2_dist_tests.log
tail -n 15 gpurun_out/n2_dist_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "rc=$?" >> gpurun_out/n2_bench.err
tail -n 5 gpurun_out/n2_bench.err
cat gpurun_out/n2_bench.json
