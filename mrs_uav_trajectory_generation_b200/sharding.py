"""Multi-GPU use of the path: one process per GPU (torch.distributed for the plumbing), problems sharded by index.

Every path is an independent problem (the reference is stateless per request, one path at a time on one thread --
src/mrs_trajectory_generation.cpp:1513), so the solve / feasibility / sampling path needs NO collective: rank g owns the
contiguous block [g*B/G, (g+1)*B/G) and copies its own inputs and outputs.  The one exchange step is the best-candidate
reduction of a segment-time sweep (BASELINE config 5): each rank evaluates its slice of the candidates, then a single
all_gather of (cost, global index, S times) -- 8*(2+S) bytes per rank -- lets every rank take the minimum.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous block of problem (or candidate) indices owned by `rank`: [rank*n/world, (rank+1)*n/world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_paths(wp_off, wp, rank, world, stop_at=None, init14=None):
    """Slices a ragged batch to the block owned by `rank`.  Returns (wp_off_local, wp_local, stop_local, init_local, p0)."""
    wp_off = np.asarray(wp_off, dtype=np.int32)
    B = len(wp_off) - 1
    p0, p1 = shard_bounds(B, rank, world)
    v0, v1 = int(wp_off[p0]), int(wp_off[p1])
    loc_off = (wp_off[p0:p1 + 1] - v0).astype(np.int32)
    return (loc_off, np.ascontiguousarray(wp[v0:v1]), None if stop_at is None else np.ascontiguousarray(stop_at[v0:v1]),
            None if init14 is None else np.ascontiguousarray(init14[p0:p1]), p0)


def optimize_sharded(ctx, wp_off, wp, params=None, rank=0, world=1, stop_at=None, init14=None):
    """Runs optimize() on this rank's block only.  Returns (results, outputs, first_problem_index).  No collective."""
    loc_off, loc_wp, loc_stop, loc_init, p0 = shard_paths(wp_off, wp, rank, world, stop_at, init14)
    if len(loc_off) < 2:
        return None, None, p0
    res, _ = ctx.optimize_batch(loc_off, loc_wp, loc_stop, loc_init, params)
    return res, ctx.fetch_outputs(), p0


def sweep_best_distributed(ctx, vmask, vval, cand, r=2, rank=0, world=1, group=None):
    """Best of K candidate segment-time vectors for ONE problem, candidates sharded over the ranks.

    cand: [K, S] (every rank passes the same array, or at least its own slice at the right rows).
    Returns (best_cost, best_global_index, best_times[S]); ties go to the lowest global index (the first minimum, as a
    serial scan would find it).  With world == 1 no process group is needed."""
    cand = np.ascontiguousarray(cand, dtype=np.float64)
    K, S = cand.shape
    k0, k1 = shard_bounds(K, rank, world)
    mine = np.empty(2 + S)
    if k1 > k0:
        _, bi, bc = ctx.sweep_costs(vmask, vval, cand[k0:k1], r=r, want_costs=False)
        mine[0], mine[1] = bc, float(k0 + bi)
        mine[2:] = cand[k0 + bi]
    else:
        mine[0], mine[1] = np.inf, float(K)
        mine[2:] = 0.0
    if world == 1:
        allr = mine[None]
    else:
        import torch
        import torch.distributed as dist

        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.from_numpy(mine.copy())
        if on_gpu:
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t, group=group)  # the single collective of the path
        allr = np.stack([o.cpu().numpy() for o in out])
    order = np.lexsort((allr[:, 1], allr[:, 0]))  # by cost, then by global index
    best = allr[order[0]]
    return float(best[0]), int(best[1]), best[2:].copy()
