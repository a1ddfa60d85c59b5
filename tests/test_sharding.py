"""Host-side multi-GPU logic on CPU: world_size 2, gloo backend (the GPU runs use nccl through the same code)."""
import os
import subprocess
import sys

import numpy as np

import oracle_lib as O
from mrs_uav_trajectory_generation_b200 import sharding, workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_gloo(emu_ctx, oracle, tmp_path):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE="2")  # concurrent runs must not collide
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)], env=dict(env, RANK=str(r)))
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    # (a) shards == unsharded run == oracle
    B = 11
    wp_off, wp = W.random_flier_paths(B, first_index=500)
    res, _ = emu_ctx.optimize_batch(wp_off, wp, None, None, emu_ctx.L.default_params())
    out = emu_ctx.fetch_outputs()
    sh = [np.load(tmp_path / f"shard_{r}.npz") for r in range(2)]
    assert int(sh[0]["p0"]) == 0 and int(sh[1]["p0"]) == 5
    assert np.array_equal(np.concatenate([s["n_samples"] for s in sh]), res["n_samples"])
    assert np.array_equal(np.concatenate([s["rounds"] for s in sh]), res["rounds"])
    for k in ("coef", "times", "samples"):
        assert np.array_equal(np.concatenate([s[k] for s in sh]), out[k]), k
    # (b) the sweep: both ranks agree, the tie goes to the lower global index, and the cost is the oracle's
    sw = [np.load(tmp_path / f"sweep_{r}.npz") for r in range(2)]
    assert int(sw[0]["index"]) == int(sw[1]["index"]) and float(sw[0]["cost"]) == float(sw[1]["cost"])
    cand = sw[0]["cand"]
    path = W.random_flier_path(0xB200 & 0xFFF, 11)
    V = len(path)
    mask = np.ones(V, np.uint8)
    mask[0] = mask[-1] = 0b111
    vals = np.zeros((V, 5, 4))
    vals[:, 0] = path
    costs = O.sweep_costs(mask, vals, 2, cand)
    assert int(sw[0]["index"]) == int(np.argmin(costs)) and float(sw[0]["cost"]) == float(costs.min())
    assert int(sw[0]["index"]) != 301 or int(np.argmin(costs)) == 301
    assert np.array_equal(sw[0]["times"], cand[int(sw[0]["index"])])
