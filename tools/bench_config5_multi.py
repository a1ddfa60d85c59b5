#!/usr/bin/env python3
"""BASELINE config 5 (10^6 candidate segment-time vectors of one problem) on several GPUs of one box, both ways the path offers:

  torchrun --nproc-per-node N tools/bench_config5_multi.py     one rank per GPU, per-rank tg_sweep_costs on its shard, then ONE
                                                               NCCL all_gather of (cost, global index, S times) = 8 (2 + S) bytes
                                                               per rank (sharding.sweep_best_distributed)
  python tools/bench_config5_multi.py --contexts N             one process, one context + host thread per GPU behind the C ABI
                                                               (tg_sweep_best)

Rank 0 prints one JSON line; the winner is checked against the oracle's scan of the candidates around it."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mrs_uav_trajectory_generation_b200 as tg  # noqa: E402
from mrs_uav_trajectory_generation_b200 import sharding, workloads as W  # noqa: E402


def problem(K):
    import oracle_lib as O

    O.build_oracle(ref=False)
    O.set_math_mode(O.MATH_DET)
    path = W.random_flier_path(0xB200, 11)
    V = len(path)
    mask = np.ones(V, np.uint8)
    mask[0] = mask[-1] = 0b111
    vals = np.zeros((V, 5, 4))
    vals[:, 0] = path
    base = O.estimate_times(path)[0]
    rng = np.random.Generator(np.random.Philox(key=0xB200))
    cand = np.maximum(base * np.exp(rng.uniform(-0.5, 0.5, (K, V - 1))), 0.01)
    return O, mask, vals, cand


def check(O, mask, vals, cand, bi, bc):
    K = len(cand)
    n = min(20000, K)
    lo = max(0, min(bi - n // 2, K - n))
    cc = O.sweep_costs(mask, vals, 2, cand[lo: lo + n])
    return bool(float(cc[bi - lo]) == bc and int(np.argmin(cc)) + lo == bi)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=1000000)
    ap.add_argument("--contexts", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    O, mask, vals, cand = problem(a.candidates)
    if a.contexts > 0:
        import torch

        n_dev = max(1, torch.cuda.device_count())
        lib = tg.Library(os.environ.get("TG_LIB") or None)
        ctxs = [tg.Context(lib, g % n_dev) for g in range(a.contexts)]
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            bc, bi, bt = tg.Context.sweep_best(ctxs, mask, vals, cand)
            ts.append(time.perf_counter() - t0)
        print(json.dumps({"config": 5, "mode": "tg_sweep_best, one process", "contexts": a.contexts, "devices": min(a.contexts, n_dev), "candidates": a.candidates,
                          "candidates_per_s_host_buffers": a.candidates / min(ts), "best_index": int(bi), "best_cost": bc,
                          "oracle_agrees_on_best": check(O, mask, vals, cand, bi, bc)}))
        return
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    ctx = tg.Context(tg.Library(os.environ.get("TG_LIB") or None), local)
    ts = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bc, bi, bt = sharding.sweep_best_distributed(ctx, mask, vals, cand, r=2, rank=rank, world=world)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ts.append(float(dt[0]))
    if rank == 0:
        print(json.dumps({"config": 5, "mode": "one rank per GPU, NCCL all_gather of (cost, index, times)", "n_gpus": world, "candidates": a.candidates,
                          "candidates_per_s_host_buffers": a.candidates / min(ts), "exchange_bytes_per_rank": 8 * (2 + cand.shape[1]), "best_index": int(bi), "best_cost": bc,
                          "best_times_match_candidate": bool(np.array_equal(bt, cand[bi])), "oracle_agrees_on_best": check(O, mask, vals, cand, bi, bc)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
