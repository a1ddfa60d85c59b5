"""Worker of tests/test_sharding.py: one process per rank, gloo backend on 127.0.0.1, host-emulation library.
Checks (a) that the shards of optimize() reassemble to the single-process result bit for bit and (b) that the
distributed sweep (the path's single all_gather) finds the same first minimum as a one-rank scan."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist

    from mrs_uav_trajectory_generation_b200 import Context, Library, sharding, workloads as W

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = Context(Library(os.path.join(ROOT, "tests", "host_emu", "libtg_emu.so")), 0)
    out_dir = sys.argv[1]
    # (a) sharded optimize: every rank writes its block; the test compares with the unsharded run
    B = 11
    wp_off, wp = W.random_flier_paths(B, first_index=500)
    res, out, p0 = sharding.optimize_sharded(ctx, wp_off, wp, ctx.L.default_params(), rank, world)
    np.savez(os.path.join(out_dir, f"shard_{rank}.npz"), p0=p0, n_samples=res["n_samples"], rounds=res["rounds"], coef=out["coef"], times=out["times"],
             samples=out["samples"], seg_off=out["seg_off"], smp_off=out["smp_off"])
    # (b) distributed sweep with a tie on purpose: candidate 7 is repeated at index 301 (rank 1's slice) -> index 7 must win
    path = W.random_flier_path(0xB200 & 0xFFF, 11)
    V = len(path)
    mask = np.ones(V, np.uint8)
    mask[0] = mask[-1] = 0b111
    vals = np.zeros((V, 5, 4))
    vals[:, 0] = path
    rng = np.random.default_rng(9)
    base = np.full(V - 1, 1.3)
    cand = np.maximum(base * np.exp(rng.uniform(-0.5, 0.5, (400, V - 1))), 0.01)
    cand[301] = cand[7]
    bc, bi, bt = sharding.sweep_best_distributed(ctx, mask, vals, cand, r=2, rank=rank, world=world)
    np.savez(os.path.join(out_dir, f"sweep_{rank}.npz"), cost=bc, index=bi, times=bt, cand=cand)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
