#!/usr/bin/env python
"""BASELINE configs[2] (Netflix shape 480 189 x 17 770, 100.5 M interactions, K = 256, Cholesky
solver) on one GPU or, under torch.distributed.run, row-sharded over N: irspack_b200.dist.run_c4
with that shape and solver (device-built power-law matrix of the shape; the host generator of
tools/time_config.py needs two minutes for it).  Prints one JSON line on rank 0.

    python tools/time_c3_sharded.py [--scale 0.05] [--cpu-sample 0.02]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29513 tools/time_c3_sharded.py

--cpu-sample f: rank 0 also times the CPU port (oracle, all host threads) on a matrix of the same
generator with f of the users and interactions (the items, hence the item-side row lengths, shrink
with it: stated in the output) and reports the per-interaction extrapolation.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-sample", type=float, default=0.0)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    from bench import HYPER, oracle_row_check
    from irspack_b200.dist import run_c4
    from irspack_b200.synth import SHAPES

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    U, I, nnz, K = SHAPES["netflix"]
    shape = (max(int(U * a.scale), world), I, int(nnz * a.scale), K)
    res = run_c4(HYPER, steps=a.steps, warmup=a.warmup, e2e_steps=0, score_users_per_rank=0, shape=shape,
                 solver="CHOLESKY", parity_check=oracle_row_check)
    if res is not None:
        U, I, nnz = res["n_users"], res["n_items"], res["nnz"]
        flops = 2.0 * nnz * K * (K + 1) + (U + I) * (K ** 3 / 3.0 + 2.0 * K * K) + 4.0 * nnz * K  # SURVEY 8 d
        res["algorithmic_tflop_per_epoch"] = flops / 1e12
        res["tflops"] = flops / (res["ms_per_epoch"] / 1e3) / 1e12
        res.pop("achieved_gbs", None)
        res.pop("algorithmic_bytes_per_epoch", None)
        if a.cpu_sample > 0:
            import numpy as np
            import scipy.sparse as sps

            import oracle
            from irspack_b200.dist import synth_user_block_device
            from irspack_b200.synth import init_factors

            f = a.cpu_sample
            Us, nnzs = max(int(U * f), 1), int(nnz * f)
            ip, ix, dt = synth_user_block_device(Us, I, nnzs, seed=1003, device=torch.device(f"cuda:{local_rank}"),
                                                 item_seed=1003)
            X = sps.csr_matrix((dt.cpu().numpy(), ix.cpu().numpy(), ip.cpu().numpy()), shape=(Us, I))
            nt = oracle.hardware_threads()
            o = oracle.OracleTrainer(X, K, HYPER["alpha0"], HYPER["reg"], HYPER["nu"], oracle.LOSS_IALSPP)
            o.user, o.item = init_factors(Us, K, 1), init_factors(I, K, 2)
            t0 = time.perf_counter()
            o.step(oracle.SOLVER_CHOLESKY, 3, nt)
            cdt = time.perf_counter() - t0
            fs = 2.0 * nnzs * K * (K + 1) + (Us + I) * (K ** 3 / 3.0 + 2.0 * K * K) + 4.0 * nnzs * K
            res["cpu_port"] = {"kind": "port", "cores": nt, "sample": f"{f:g} of the users and interactions "
                               f"({Us} x {I}, {nnzs} nnz), one epoch", "ms_per_epoch_sample": 1e3 * cdt,
                               "tflops": fs / cdt / 1e12,
                               "full_epoch_ms_extrapolated_by_flops": 1e3 * cdt * flops / fs}
        res["env"] = {k: v for k, v in os.environ.items() if k.startswith("IALS_")}
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
