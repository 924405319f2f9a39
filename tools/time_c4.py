#!/usr/bin/env python
"""BASELINE.json configs[3] and [4]: iALS K=128 (CG) on a synthetic power-law matrix of
10 M users x 2 M items with 1 B interactions, row-sharded across the GPUs of one box, and the
batch scoring + seen-mask + top-100 of the same model.  Prints one JSON line on rank 0.

    python tools/time_c4.py --scale 0.01                                  # one GPU, 1 % of the shape
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29511 tools/time_c4.py [--scale 1.0] [--epochs 3] [--score-users 200000]

STRONG scaling: the matrix is fixed (``--scale`` shrinks users, items and nnz together for dry
runs), every rank generates its own user block ON THE DEVICE (irspack_b200.dist
synth_user_block_device), the rows of X^T are exchanged device to device, and the trainer takes
the device CSR as is: no interaction ever touches the host.  Timing as in bench.py: CUDA events
on the launching stream around exactly ``--epochs`` epochs after ``--warmup`` ones, barrier +
synchronize on both sides, MAX over ranks.  Not a bench line: bench.py owns that contract; this
records the two 8-GPU configurations (profiles/).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the 10M x 2M x 1B shape")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--score-users", type=int, default=200_000,
                    help="users scored per rank for configs[4] (0: skip, -1: the whole shard)")
    ap.add_argument("--score-block", type=int, default=16384)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    import irspack_b200
    from bench import measured_peaks, solve_bytes
    from irspack_b200 import _ials_core as core
    from irspack_b200.dist import (ShardedIALSTrainer, exchange_transposed_shards_device,
                                   global_item_bounds_device, synth_user_block_device)
    from irspack_b200.synth import SHAPES

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    U0, I0, nnz0, K = SHAPES["powerlaw1b"]
    U, I, nnz = max(int(U0 * a.scale), world), max(int(I0 * a.scale), 128), int(nnz0 * a.scale)
    # user blocks of equal row count and equal nnz (the generator draws exactly nnz/world pairs per
    # block): nnz-balanced by construction; items are cut by their global degrees
    ub = np.linspace(0, U, world + 1).astype(np.int64)
    nb = np.linspace(0, nnz, world + 1).astype(np.int64)
    n_rows, n_nnz = int(ub[rank + 1] - ub[rank]), int(nb[rank + 1] - nb[rank])

    t0 = time.perf_counter()
    ip, ix, dt = synth_user_block_device(n_rows, I, n_nnz, seed=1004 + rank, device=dev, item_seed=1004)
    item_bounds = global_item_bounds_device(ix, I)
    t_ip, t_ix, t_dt = exchange_transposed_shards_device(ip, ix, dt, int(ub[rank]), U, item_bounds)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0

    hyper = dict(alpha0=0.1, reg=1e-3, nu=1.0, max_cg_steps=3)
    cfg = (core.IALSModelConfigBuilder().set_K(K).set_alpha0(hyper["alpha0"]).set_reg(hyper["reg"])
           .set_nu(hyper["nu"]).build())
    sc = core.IALSSolverConfigBuilder().set_max_cg_steps(hyper["max_cg_steps"]).build()
    t0 = time.perf_counter()
    tr = ShardedIALSTrainer(cfg, (ip, ix, dt), int(ub[rank]), U, (t_ip, t_ix, t_dt),
                            int(item_bounds[rank]), I, init_on_device=True)
    del ip, ix, dt, t_ip, t_ix, t_dt
    torch.cuda.empty_cache()
    t_plan = time.perf_counter() - t0

    def barrier() -> None:
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(a.warmup):
        tr.step_async(sc)
    tr.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(a.epochs):
        tr.step_async(sc)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    tr.sync()  # raises if a solver flagged a failure
    ms = max_over_ranks(ev0.elapsed_time(ev1)) / a.epochs

    # configs[4]: score + seen-mask + top-k of this rank's own users (item factors are replicated:
    # no collective, SURVEY.md 8 e), fused tcgen05 kernel, only k (index, score) pairs per user
    # leave the device
    score = None
    n_score = n_rows if a.score_users < 0 else min(a.score_users, n_rows)
    if n_score > 0:
        lib = irspack_b200._lib.lib
        k = min(a.topk, I)
        b0 = int(ub[rank])
        idx = np.empty((a.score_block, k), dtype=np.int32)
        cnt = np.empty((a.score_block,), dtype=np.int32)
        p = core._ptr

        def recommend(b: int, e: int) -> None:
            irspack_b200._lib.check(lib.ials_trainer_recommend(
                tr._handle, b, e, k, 0, p(None), p(None), p(idx), p(None), p(cnt)))

        recommend(b0, b0 + min(a.score_block, n_score))  # warm-up (scratch allocation)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for b in range(b0, b0 + n_score, a.score_block):
            recommend(b, min(b + a.score_block, b0 + n_score))
        torch.cuda.synchronize()
        dt_score = max_over_ranks(time.perf_counter() - t0)
        scored = n_score
        if world > 1:
            tot = torch.tensor([n_score], device=dev, dtype=torch.int64)
            dist.all_reduce(tot)
            scored = int(tot.item())
        score = {"users_scored": scored, "k": k, "seconds": dt_score,
                 "users_per_s": scored / dt_score,
                 "algorithmic_tflops": 2.0 * scored * I * K / dt_score / 1e12,
                 "what": "ials_trainer_recommend(mask='train') over each rank's own users in blocks of "
                         f"{a.score_block}; host wall clock incl. the D2H of k indices per user, max over ranks"}

    if rank == 0:
        peak, peak_kind = measured_peaks()
        w = dict(n_users=U, n_items=I, nnz=nnz, K=K)
        algo = solve_bytes(w) + (U + I) * 4 * K
        line = {
            "config": "c4" if a.scale == 1.0 else f"c4 x {a.scale}", "n_gpus": world, "scaling": "strong",
            "n_users": U, "n_items": I, "nnz": nnz, "K": K, "solver": "CG", **hyper,
            "epochs_timed": a.epochs, "warmup": a.warmup, "ms_per_epoch": ms,
            "epochs_per_s": 1e3 / ms, "interactions_per_s": nnz / (ms / 1e3),
            "algorithmic_gb_per_epoch": algo / 1e9, "achieved_gbs": algo / (ms / 1e3) / 1e9,
            "hbm_peak_gbs_per_gpu": peak, "peak_kind": peak_kind,
            "frac_of_hbm_peak": algo / (ms / 1e3) / 1e9 / (peak * world),
            "build_s": round(t_build, 2), "plan_s": round(t_plan, 2),
            "data": "synthetic power-law block per rank, generated and transposed on the device",
            "score_topk": score,
        }
        print(json.dumps(line), flush=True)
    del tr
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
