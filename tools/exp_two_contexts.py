#!/usr/bin/env python3
"""Experiment: C contexts on ONE GPU, each driven by its own host thread over its own 65 536-path batches (staggered start), against one
context.  The tail rounds of a batch (a few hundred paths in rounds 4-6, each round a chain of latency-bound launches) leave the GPU
idle; a second context's dense phases can run there.  Prints total trajectories/s per C."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mrs_uav_trajectory_generation_b200 as tg  # noqa: E402
from mrs_uav_trajectory_generation_b200 import workloads as W  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = 6
lib = tg.Library()
for C in (1, 2, 3):
    ctxs = [tg.Context(lib, 0) for _ in range(C)]
    batches = [W.random_flier_paths_fast(B, first_index=c) for c in range(C)]
    dev = [torch.from_numpy(b[1]).cuda() for b in batches]
    P = lib.default_params()
    for c in range(C):  # warm-up
        ctxs[c].optimize_batch(batches[c][0], dev[c].data_ptr(), None, None, P, inputs_on_device=True)
    torch.cuda.synchronize()

    def run(c):
        if c:
            time.sleep(0.06 * c)  # stagger the phases
        for _ in range(steps):
            ctxs[c].optimize_batch(batches[c][0], dev[c].data_ptr(), None, None, P, inputs_on_device=True)

    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(c,)) for c in range(C)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"contexts {C}: {C * steps * B / dt:,.0f} trajectories/s ({1e3 * dt / steps:.1f} ms per round of {C} batches)")
    del ctxs
