"""Accuracy of include/tg_detmath.h -- the deterministic libm subset shared by the CUDA kernels and the oracle's DET mode --
measured against mpmath (60 digits) on the argument ranges the path uses, next to glibc on the same arguments.

Reported per function: worst error in ulps of the detmath result, of the glibc result, and the fraction of arguments on which
the two agree bit for bit.  The numbers feed DESIGN.md's "numeric floor" paragraph: wherever detmath and glibc differ by one
ulp, a segment time differs by one ulp, and cond(Rpp) ~ 1e8..1e13 carries that into the coefficients (tests/test_numeric_floor.py).
"""
import math

import mpmath as mp
import numpy as np
import pytest

import oracle_lib as O

mp.mp.dps = 60

FN = {"log": 0, "exp": 1, "sin": 2, "cos": 3, "atan2": 4, "cbrt": 5, "pow_int": 6, "hypot": 7}


def ulp(x):
    return math.ulp(x) if x != 0 else 5e-324


def _err_ulps(got, exact):
    e = float(exact)  # correctly rounded reference value
    return abs(mp.mpf(got) - exact) / mp.mpf(ulp(e))


def _scan(name, args, exact_fn, libm_fn):
    worst_det = worst_libm = mp.mpf(0)
    agree = 0
    cr_det = 0
    for a in args:
        a = a if isinstance(a, tuple) else (a,)
        det = O.math_fn(FN[name], *a)
        lib = libm_fn(*a)
        ex = exact_fn(*a)
        worst_det = max(worst_det, _err_ulps(det, ex))
        worst_libm = max(worst_libm, _err_ulps(lib, ex))
        agree += det == lib
        cr_det += det == float(ex)
    n = len(args)
    return float(worst_det), float(worst_libm), agree / n, cr_det / n


@pytest.fixture(scope="module")
def det(oracle):
    O.set_math_mode(O.MATH_DET)
    yield
    O.set_math_mode(O.MATH_DET)


def test_detmath_accuracy(det, record_property):
    rng = np.random.default_rng(11)
    N = 4000
    rows = {}
    # log / exp: Jenkins-Traub scaling (rpoly_ak1.cpp:156,227,246) and the soft-constraint cost exp(violation * weight)
    xs = [float(x) for x in np.exp(rng.uniform(-40, 40, N))]
    rows["log"] = _scan("log", xs, lambda x: mp.log(mp.mpf(x)), math.log)
    xs = [float(x) for x in rng.uniform(-80, 80, N)]
    rows["exp"] = _scan("exp", xs, lambda x: mp.exp(mp.mpf(x)), math.exp)
    # sin / cos: inclinations in [-pi/2, pi/2] (eth/vertex.cpp:512-520), half headings for the quaternion (|yaw|/2 up to ~20)
    xs = [float(x) for x in np.concatenate([rng.uniform(-1.6, 1.6, N // 2), rng.uniform(-20, 20, N // 2)])]
    rows["sin"] = _scan("sin", xs, lambda x: mp.sin(mp.mpf(x)), math.sin)
    rows["cos"] = _scan("cos", xs, lambda x: mp.cos(mp.mpf(x)), math.cos)
    # atan2: inclination atan2(dz, hypot) and limit angle atan2(v_v, v_h), yawFromQuaternion
    pts = [(float(y), float(x)) for y, x in zip(rng.uniform(-3, 3, N), rng.uniform(-3, 3, N))]
    rows["atan2"] = _scan("atan2", pts, lambda y, x: mp.atan2(mp.mpf(y), mp.mpf(x)), math.atan2)
    # cbrt: jerk violation ratios (eth/trajectory.cpp:642)
    xs = [float(x) for x in np.exp(rng.uniform(-12, 12, N))]
    rows["cbrt"] = _scan("cbrt", xs, lambda x: mp.cbrt(mp.mpf(x)), lambda x: math.copysign(abs(x) ** (1.0 / 3.0), x) if False else np.cbrt(x))
    # pow(T, e), e = 1..15 (lin_impl.h:612-615)
    pts = [(float(t), float(e)) for t, e in zip(np.exp(rng.uniform(-4.6, 4.0, N)), rng.integers(1, 16, N))]
    rows["pow_int"] = _scan("pow_int", pts, lambda t, e: mp.mpf(t) ** int(e), lambda t, e: math.pow(t, e))
    # hypot: spacing of consecutive samples (node.cpp:1588)
    pts = [(float(a), float(b)) for a, b in zip(rng.uniform(-2, 2, N) * np.exp(rng.uniform(-8, 0, N)), rng.uniform(-2, 2, N) * np.exp(rng.uniform(-8, 0, N)))]
    rows["hypot"] = _scan("hypot", pts, lambda a, b: mp.sqrt(mp.mpf(a) ** 2 + mp.mpf(b) ** 2), math.hypot)
    for k, (wd, wl, ag, cr) in rows.items():
        print(f"detmath {k:8s}: worst {wd:.3f} ulp (glibc {wl:.3f} ulp), bit-identical to glibc on {100 * ag:.2f} %, correctly rounded on {100 * cr:.2f} %")
        record_property(f"detmath_{k}", (wd, wl, ag, cr))
    # the header's claim: correctly rounded on the path's ranges (double-double evaluation, one rounding)
    for k, (wd, _, ag, cr) in rows.items():
        assert wd <= 0.5000001 and cr >= 0.9995, (k, wd, cr)
        assert ag >= 0.99, (k, ag)  # glibc itself is correctly rounded on all but a few per mille of these arguments


def test_detmath_special_values(det):
    inf = float("inf")
    assert O.math_fn(FN["log"], 0.0) == -inf and math.isnan(O.math_fn(FN["log"], -1.0)) and O.math_fn(FN["log"], 1.0) == 0.0
    assert O.math_fn(FN["exp"], 0.0) == 1.0 and O.math_fn(FN["exp"], 800.0) == inf and O.math_fn(FN["exp"], -800.0) == 0.0
    assert O.math_fn(FN["sin"], 0.0) == 0.0 and O.math_fn(FN["cos"], 0.0) == 1.0
    assert O.math_fn(FN["atan2"], 0.0, 1.0) == 0.0 and O.math_fn(FN["atan2"], 0.0, -1.0) == math.pi
    assert O.math_fn(FN["atan2"], 1.0, 0.0) == math.pi / 2 and O.math_fn(FN["atan2"], -1.0, 0.0) == -math.pi / 2
    assert O.math_fn(FN["cbrt"], 27.0) == 3.0 and O.math_fn(FN["cbrt"], -8.0) == -2.0 and O.math_fn(FN["cbrt"], 0.0) == 0.0
    assert O.math_fn(FN["pow_int"], 2.0, 10.0) == 1024.0
