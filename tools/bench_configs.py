#!/usr/bin/env python3
"""Measures BASELINE.json configs 1, 2, 4 and 5 on one GPU beside the CPU oracle (config 3 is bench.py's headline).
Prints one JSON object per config.  Usage (on the GPU box): python tools/bench_configs.py [--quick] > profiles/rNN_configs.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mrs_uav_trajectory_generation_b200 as tg  # noqa: E402
import oracle_lib as O  # noqa: E402
from mrs_uav_trajectory_generation_b200 import workloads as W  # noqa: E402


def p50(xs):
    return float(np.median(np.asarray(xs)))


def main():
    quick = "--quick" in sys.argv
    O.build_oracle(ref=False)
    O.set_math_mode(O.MATH_DET)
    ctx = tg.Context(tg.Library(os.environ.get("TG_LIB") or None), 0)
    P = ctx.L.default_params()
    cores = os.cpu_count()

    # ---- config 1: single paths, full pipeline, latency (host call -> samples on the host)
    for name, wp, init in (("F1a 4-waypoint test path + prepended start", W.F1A_WAYPOINTS, W.init14(W.F1A_INIT_HEADING)),
                           ("F1b 10-waypoint zig-zag + prepended hover", W.F1B_WAYPOINTS, W.init14(W.F1B_INIT_HEADING))):
        wp_off = np.array([0, len(wp)], np.int32)
        gpu, cpu = [], []
        for rep in range(5 if quick else 21):
            t0 = time.perf_counter()
            res, _ = ctx.optimize_batch(wp_off, wp, None, init[None], P)
            out = ctx.fetch_outputs()
            gpu.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            r = O.optimize_path(wp, init=init)
            cpu.append(time.perf_counter() - t0)
        same = np.array_equal(out["coef"], r["coeffs"]) and np.array_equal(out["samples"], r["samples"])
        print(json.dumps({"config": 1, "case": name, "metric": "p50 latency", "unit": "ms", "gpu_ms": 1e3 * p50(gpu[1:]), "cpu_oracle_ms": 1e3 * p50(cpu[1:]),
                          "rounds": int(res["rounds"][0]), "final_segments": int(res["n_waypoints"][0] - 1), "samples": int(res["n_samples"][0]),
                          "n_evals": int(res["n_evals"][0]), "scale_passes": int(res["n_scale_passes"][0]), "bit_exact_vs_oracle": bool(same)}))

    # ---- config 2: 4096 random paths, linear solve + sampling
    B = 4096
    wp_off, wp = W.random_flier_paths_fast(B, first_index=2)
    for r in (2, 4):
        P2 = ctx.L.default_params(run_time_alloc=0, check_deviation=0, derivative_to_optimize=r)
        dev, wall = [], []
        for rep in range(4 if quick else 12):
            t0 = time.perf_counter()
            res, tot = ctx.optimize_batch(wp_off, wp, None, None, P2)
            out = ctx.fetch_outputs(want=("smp_off", "samples"))
            wall.append(time.perf_counter() - t0)
            dev.append(ctx.last_device_ms() * 1e-3)
        n_cpu = 512
        t0 = time.perf_counter()
        O.optimize_batch(wp_off[: n_cpu + 1], wp[: wp_off[n_cpu]], params=O.default_params(run_time_alloc=0, check_deviation=0, derivative_to_optimize=r),
                         cap_wp=16, cap_samples=400)
        cpu_s = time.perf_counter() - t0
        print(json.dumps({"config": 2, "r": r, "paths": B, "metric": "trajectories/s (linear solve + dt sampling)", "gpu_device": B / p50(dev[2:]),
                          "gpu_e2e_host_buffers": B / p50(wall[2:]), "cpu_oracle": n_cpu / cpu_s, "cpu_threads": cores, "cpu_sample_paths": n_cpu,
                          "mean_samples": float(res["n_samples"].mean())}))

    # ---- config 4: 200-waypoint path, p50 single-problem latency
    gpu, cpu, segs = [], [], []
    for seed in range(3 if quick else 11):
        path = W.random_flier_path(7000 + seed, 200)
        wp_off4 = np.array([0, 200], np.int32)
        t0 = time.perf_counter()
        res, _ = ctx.optimize_batch(wp_off4, path, None, None, P)
        out = ctx.fetch_outputs(want=("smp_off", "samples"))
        gpu.append(time.perf_counter() - t0)
        segs.append(int(res["n_waypoints"][0] - 1))
        if seed < (1 if quick else 3):
            t0 = time.perf_counter()
            O.optimize_path(path, cap_wp=13000, cap_samples=60000)
            cpu.append(time.perf_counter() - t0)
    print(json.dumps({"config": 4, "metric": "p50 latency, 200-waypoint path, full pipeline", "unit": "ms", "gpu_ms": 1e3 * p50(gpu[1:] or gpu),
                      "cpu_oracle_ms": 1e3 * p50(cpu), "final_segments_median": int(np.median(segs)), "note": "one problem at a time: the GPU has no batch to fill its SMs"}))

    # ---- config 5: candidate segment-time sweep for one problem
    K = 100000 if quick else 1000000
    path = W.random_flier_path(0xB200, 11)
    V = len(path)
    mask = np.ones(V, np.uint8)
    mask[0] = mask[-1] = 0b111
    vals = np.zeros((V, 5, 4))
    vals[:, 0] = path
    base = O.estimate_times(path)[0]
    rng = np.random.Generator(np.random.Philox(key=0xB200))
    cand = np.maximum(base * np.exp(rng.uniform(-0.5, 0.5, (K, V - 1))), 0.01)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        costs, bi, bc = ctx.sweep_costs(mask, vals, cand, r=2, want_costs=False)
        ts.append(time.perf_counter() - t0)
    dev_s = ctx.last_device_ms() * 1e-3
    n_cpu = 20000
    lo = max(0, min(bi - n_cpu // 2, K - n_cpu))
    t0 = time.perf_counter()
    cc = O.sweep_costs(mask, vals, 2, cand[lo: lo + n_cpu])
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"config": 5, "candidates": K, "metric": "candidates/s (update times + solve + cost, argmin)", "gpu_device": K / dev_s,
                      "gpu_e2e_host_buffers": K / min(ts), "cpu_oracle": n_cpu / cpu_s, "cpu_threads": cores, "best_index": int(bi), "best_cost": bc,
                      "oracle_agrees_on_best": bool(float(cc[bi - lo]) == bc and int(np.argmin(cc)) + lo == bi)}))


if __name__ == "__main__":
    main()
