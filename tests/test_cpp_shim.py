"""The C++ host side (include/eth_trajectory_generation_b200.hpp, the reference's class names over the C ABI): a C++
program written like MrsTrajectoryGeneration::findTrajectory (tests/cpp/test_shim.cpp) is compiled, linked against the
C-ABI library and its results compared bit for bit with the oracle.  CPU runs link the host emulation of the device
code; the GPU run links libtg_b200.so."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from mrs_uav_trajectory_generation_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_shim(libdir, libname, tmp_path):
    exe = str(tmp_path / "test_shim")
    out = str(tmp_path / "shim.txt")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                           "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir])
    subprocess.check_call([exe, out])
    res = {}
    for line in open(out):
        f = line.split()
        res.setdefault(f[0], []).append(np.array([float.fromhex(x) for x in f[2:]]))
    return res


def _check(res):
    wps = W.F1B_WAYPOINTS
    V, r = len(wps), 2
    mask = np.ones(V, np.uint8)
    mask[0] = mask[-1] = 0b111
    vals = np.zeros((V, O.HALF, O.D))
    vals[:, 0] = wps
    times = np.array([1.0 + 0.25 * (i % 3) for i in range(V - 1)])
    # PolynomialOptimization<10>
    coef, cost, _, _ = O.solve_linear(mask, vals, times, r)
    assert np.array_equal(res["lin_times"][0], times)
    assert np.array_equal(res["lin_coef"][0].reshape(V - 1, 4, 10), coef)
    assert res["lin_cost"][0][0] == cost
    tq = [0.0, 2.6, float(np.add.accumulate(times)[-1]) * 0.75]
    for k, t in enumerate(tq):
        assert np.array_equal(res["lin_eval_p"][k], O.trajectory_evaluate(coef, times, t, 0)[0])
        assert np.array_equal(res["lin_eval_v"][k], O.trajectory_evaluate(coef, times, t, 1)[0])
    assert res["lin_bad_setup"][0][0] == 0.0  # wrong number of segment times: setupFromVertices returns false
    # PolynomialOptimizationNonLinear<10> + sampleWholeTrajectory
    ref = O.time_alloc(mask, vals, times, r)
    code, evals, passes, fcost = res["nl_meta"][0]
    assert (int(code), int(evals), int(passes)) == (ref["nlopt_code"], ref["n_evals"], ref["n_scale_passes"]) and fcost == ref["final_cost"]
    assert np.array_equal(res["nl_times"][0], ref["times"])
    assert np.array_equal(res["nl_coef"][0].reshape(V - 1, 4, 10), ref["coef"])
    full, t_ns = O.sample(ref["coef"], ref["times"], 0.2)  # [M][19] = p4 v4 a4 j3 s3 yaw
    got = res["nl_samples"][0].reshape(-1, 7)
    assert got.shape[0] == full.shape[0]
    assert np.array_equal(got[:, :3], full[:, :3]) and np.array_equal(got[:, 3], full[:, 18])
    assert np.array_equal(got[:, 4], full[:, 4]) and np.array_equal(got[:, 5], full[:, 10])
    assert np.array_equal(got[:, 6].astype(np.int64), t_ns)
    mx = O.segment_maxima(ref["coef"], ref["times"])
    assert np.array_equal(res["nl_max_h"][0], mx[:, 0:3].max(axis=0))
    for k in (1, 2, 3):
        t, v, i = O.max_magnitude(ref["coef"], ref["times"], k)
        assert np.array_equal(res["nl_maxmag"][k - 1], [t, v, float(i)])
    # evaluateObjectives: objectiveFunctionTime (kSquaredTime) with two soft constraints at three candidate time vectors
    xs = np.stack([times, times * 1.25, times * 0.8])
    rt, rp = O.objective(mask, vals, r, 0, xs, 500.0, True, 100.0, [1, 2], [4.0, 2.0])
    assert np.array_equal(res["obj_total"][0], rt) and np.array_equal(res["obj_parts"][0].reshape(3, 3), rp)
    # TrajectoryGeneratorBatch: optimize() over two paths
    for p, path in enumerate([W.F1B_WAYPOINTS, W.F1A_WAYPOINTS]):
        o = O.optimize_path(path)
        meta = res["batch_meta"][p]
        assert (int(meta[0]), int(meta[1]), int(meta[2]), int(meta[3]), int(meta[4]), int(meta[5])) == \
            (o["res"].success, o["res"].rounds, o["res"].n_waypoints, o["res"].n_samples, o["res"].nlopt_code, o["res"].safe)
        assert np.array_equal(res["batch_times"][p], o["times"])
        assert np.array_equal(res["batch_coef"][p].reshape(-1, 4, 10), o["coeffs"])
        assert np.array_equal(res["batch_samples"][p].reshape(-1, 4), o["samples"])


def _check_side_steps(res):
    raw = W.F1B_WAYPOINTS.copy()
    raw = np.insert(raw, 3, raw[2] + np.array([0.01, 0, 0, 0]), axis=0)
    stop = np.zeros(len(raw), np.uint8)
    stop[6] = 1
    pw, ps = O.preprocess_path(raw, stop)
    got = res["pre"][0].reshape(-1, 5)
    assert np.array_equal(got[:, :4], pw) and np.array_equal(got[:, 4].astype(np.uint8), ps) and len(pw) == len(raw) - 1
    lim = np.array(O.DEFAULT_LIMITS) * np.array([0.75, 0.75, 0.75, 0.75, 1, 1, 1, 1, 1])
    fb = O.fallback_sample(pw, ps, lim, 0.2, 2.0)
    assert np.array_equal(res["fallback"][0].reshape(-1, 4), fb)
    assert np.array_equal(res["idxs"][0].astype(np.int32), O.waypoint_idxs(fb, pw))


def test_cpp_shim_on_host_emulation(oracle, emu_lib, tmp_path):
    res = _run_shim(os.path.join(ROOT, "tests", "host_emu"), "tg_emu", tmp_path)
    _check(res)
    _check_side_steps(res)


@pytest.mark.gpu
def test_cpp_shim_on_gpu(oracle, gpu_ctx, tmp_path):
    res = _run_shim(os.path.join(ROOT, "mrs_uav_trajectory_generation_b200"), "tg_b200", tmp_path)
    _check(res)
    _check_side_steps(res)
