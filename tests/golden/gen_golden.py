#!/usr/bin/env python3
"""Generates the committed golden fixtures of tests/golden/ (run in the build container, where /root/reference exists).

  rpoly_reference.npz   zeros returned by the REFERENCE's own Jenkins-Traub file
                        (/root/reference/src/eth_trajectory_generation/rpoly/rpoly_ak1.cpp, compiled unmodified by
                        oracle/Makefile into oracle/_ref/libref_rpoly.so) for 600 polynomials of the kinds the path meets:
                        extremum polynomials of real trajectory segments (degrees 15/13/11/7/6/5), random polynomials
                        with exact zeros at the origin, vanishing leading coefficients, clustered zeros.
  pipeline_oracle.npz   outputs of the oracle's full pipeline (findTrajectory + validation + subdivision) for the two
                        config-1 fixtures and eight random-flier paths: final waypoints, segment
                        times, coefficients, samples, verdicts and counts.  A regression pin of the restatement itself
                        (the reference cannot be built here: Eigen / NLopt / ROS absent).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
from mrs_uav_trajectory_generation_b200 import workloads as W  # noqa: E402


def extremum_polys(coef):
    """Increasing-power coefficient vectors handed to findRootsJenkinsTraub for one segment (eth/segment.cpp:122-145)."""
    out = []
    fact = lambda k, j: float(np.prod(np.arange(j - k + 1, j + 1))) if j >= k else 0.0
    for k in (1, 2, 3):
        nd, ndd = 10 - k, 9 - k
        acc = np.zeros(nd + ndd - 1)
        for dim in (0, 1):
            c = coef[dim]
            d = np.array([c[j + k] * fact(k, j + k) for j in range(nd)])
            dd = np.array([c[j + k + 1] * fact(k + 1, j + k + 1) for j in range(ndd)])
            acc += np.convolve(d, dd)
        out.append(acc)
        for dim in (2, 3):
            c = coef[dim]
            v = np.zeros(10)
            for j in range(10 - k - 1):
                v[j] = c[j + k + 1] * fact(k + 1, j + k + 1)
            out.append(v)
    return out


def main():
    O.build_oracle(ref=True)
    assert O.ref_lib() is not None, "needs /root/reference (build container)"
    O.set_math_mode(O.MATH_DET)
    rng = np.random.default_rng(20261017)
    polys = []
    # extremum polynomials of real segments
    for p in range(12):
        r = O.optimize_path(W.random_flier_path(p, 11), params=O.default_params(check_deviation=0))
        for s in range(0, len(r["times"]), 3):
            polys += extremum_polys(r["coeffs"][s])
    # synthetic stress cases
    for i in range(240):
        n = int(rng.integers(3, 17))
        c = rng.standard_normal(n) * np.exp(rng.uniform(-4, 4, n))
        if i % 5 == 0:
            c[: int(rng.integers(1, 3))] = 0.0          # zeros at the origin
        if i % 7 == 0:
            c[-int(rng.integers(1, 3)):] = 0.0          # vanishing leading coefficients (trimmed by the wrapper)
        if i % 11 == 0:
            roots = np.concatenate([np.full(3, rng.uniform(-2, 2)), rng.uniform(-3, 3, max(n - 4, 1))])
            c = np.poly(roots)[::-1]                    # a triple zero
        polys.append(c)
    polys = polys[:600]
    L = max(len(c) for c in polys)
    P = np.zeros((len(polys), L))
    n_c = np.zeros(len(polys), np.int32)
    RE = np.zeros((len(polys), L))
    IM = np.zeros((len(polys), L))
    n_r = np.zeros(len(polys), np.int32)
    ok = np.zeros(len(polys), np.uint8)
    for i, c in enumerate(polys):
        re, im, success = O.find_roots(c, use_ref=True)
        P[i, : len(c)] = c
        n_c[i] = len(c)
        RE[i, : len(re)] = re
        IM[i, : len(im)] = im
        n_r[i] = len(re)
        ok[i] = success
    np.savez_compressed(os.path.join(HERE, "rpoly_reference.npz"), coeffs=P, n_coeffs=n_c, re=RE, im=IM, n_roots=n_r, ok=ok)
    print("rpoly_reference.npz:", len(polys), "polynomials,", int(n_r.sum()), "zeros")

    paths = [W.F1A_WAYPOINTS, W.F1B_WAYPOINTS] + [W.random_flier_path(100 + p, 11) for p in range(8)]
    inits = [W.init14(W.F1A_INIT_HEADING), W.init14(W.F1B_INIT_HEADING)] + [None] * 8
    store = {}
    for p, (wp, init) in enumerate(zip(paths, inits)):
        r = O.optimize_path(wp, init=init)
        res = r["res"]
        store[f"wp_in_{p}"] = np.asarray(wp, dtype=np.float64)
        store[f"init_{p}"] = np.zeros(0) if init is None else init
        store[f"wp_{p}"] = r["wp"]
        store[f"times_{p}"] = r["times"]
        store[f"coef_{p}"] = r["coeffs"]
        store[f"samples_{p}"] = r["samples"]
        store[f"meta_{p}"] = np.array([res.success, res.rounds, res.safe, res.n_waypoints, res.n_samples, res.nlopt_code, res.n_evals, res.n_scale_passes],
                                      dtype=np.int64)
    store["n"] = np.array(len(paths))
    np.savez_compressed(os.path.join(HERE, "pipeline_oracle.npz"), **store)
    print("pipeline_oracle.npz:", len(paths), "paths")


if __name__ == "__main__":
    main()
