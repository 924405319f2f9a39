#!/usr/bin/env python
"""BASELINE configs[3] / [4] outside bench.py: irspack_b200.dist.run_c4 at a chosen scale, on one
GPU or (under torch.distributed.run) row-sharded over N.  Prints one JSON line on rank 0.

    python tools/time_c4.py --scale 0.1 --steps 2
    IALS_HEAVY_THRESHOLD=512 python tools/time_c4.py --scale 0.1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29511 tools/time_c4.py --scale 1.0
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the 10M x 2M x 1B shape")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--score-users", type=int, default=100_000, help="users scored per rank (0: skip)")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    from bench import HYPER, oracle_row_check
    from irspack_b200.dist import run_c4

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    res = run_c4(HYPER, steps=a.steps, warmup=a.warmup, scale=a.scale, e2e_steps=a.e2e_steps,
                 score_users_per_rank=a.score_users, parity_check=oracle_row_check)
    if res is not None:
        res["env"] = {k: v for k, v in os.environ.items() if k.startswith("IALS_")}
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
