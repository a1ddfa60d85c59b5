"""Parity checks shared by the CPU-emulation tests (tests/test_emu_parity.py, `-m "not gpu"`) and the GPU tests
(tests/test_gpu_parity.py, `-m gpu`).  Each check runs the SAME inputs through the product's C ABI (`ctx`) and through
the oracle (CPU restatement of the reference, deterministic-math mode) and compares.

Tolerances (BASELINE.json north_star): coefficients 1e-9 relative, sampled positions 1e-6 m, feasibility verdicts,
subdivision counts, segment counts and sample indices exact.  Because both sides follow one numeric contract
 the observed differences are 0; the asserts use the stated tolerances and the tests also report
whether the match was bit-exact.
"""
import numpy as np

import oracle_lib as O
from mrs_uav_trajectory_generation_b200 import workloads as W

COEF_RTOL = 1e-9     # north_star: coefficients within 1e-9 relative
SAMPLE_ATOL = 1e-6   # north_star: sampled positions within 1e-6 m


def coef_rel_err(c, ref, times):
    """max_k |dc_k| T^k / max_k |c_k| T^k per (segment, dimension) -- SURVEY.md H1's definition of 'relative'."""
    c = np.asarray(c)
    ref = np.asarray(ref)
    T = np.asarray(times)[:, None, None]
    w = T ** np.arange(c.shape[-1])[None, None, :]
    num = (np.abs(c - ref) * w).max(axis=-1)
    den = (np.abs(ref) * w).max(axis=-1)
    den = np.where(den > 0, den, 1.0)
    return float((num / den).max()) if c.size else 0.0


def random_linear_problem(rng, V, kind=0):
    wp = np.cumsum(rng.uniform(-2, 2, (V, 4)), axis=0)
    m = np.ones(V, np.uint8)
    m[0] = 7
    m[-1] = 7
    v = np.zeros((V, 5, 4))
    v[:, 0, :] = wp
    if kind == 1:  # initial state present: start fixes p, v, a, j with non-zero values
        m[0] = 15
        v[0, 1:4, :] = rng.uniform(-1, 1, (3, 4))
    if kind == 2 and V > 3:  # a stop_at waypoint in the middle
        m[V // 2] = 15
    if kind == 3:  # min-snap recipe: ends fix 0..4
        m[0] = 31
        m[-1] = 31
    if kind == 4:  # generic: a vertex with velocity fixed, one with nothing but position
        m[1] = 3
        v[1, 1, :] = rng.uniform(-1, 1, 4)
    if kind == 5:  # everything fixed: n_free == 0 (lin_impl.h:344-349)
        m[:] = 31
        v[:, 1:, :] = rng.uniform(-0.2, 0.2, (V, 4, 4))
    t = rng.uniform(0.3, 3.0, V - 1)
    return m, v, t


def check_linear_batch(ctx, seed=0, B=24, r=2):
    rng = np.random.default_rng(seed)
    masks, vals, times, Vs = [], [], [], []
    for p in range(B):
        V = int(rng.integers(2, 24))
        m, v, t = random_linear_problem(rng, V, kind=p % 6)
        masks.append(m)
        vals.append(v)
        times.append(t)
        Vs.append(V)
    vtx_off = np.concatenate([[0], np.cumsum(Vs)]).astype(np.int32)
    coef, cost = ctx.solve_linear_batch(vtx_off, np.concatenate(masks), np.concatenate(vals), np.concatenate(times), r)
    worst, exact = 0.0, True
    s0 = 0
    for p in range(B):
        c, co, _, _ = O.solve_linear(masks[p], vals[p], times[p], r)
        S = Vs[p] - 1
        mine = coef[s0:s0 + S]
        worst = max(worst, coef_rel_err(mine, c, times[p]))
        assert abs(cost[p] - co) <= 1e-9 * max(1.0, abs(co)), (p, cost[p], co)
        exact = exact and np.array_equal(mine, c) and cost[p] == co
        s0 += S
    assert worst <= COEF_RTOL, worst
    return worst, exact


def check_time_alloc(ctx, seed=21, B=10, r=2):
    """tg_time_alloc_batch (PolynomialOptimizationNonLinear::optimize from vertices) against the oracle: nlopt code, evaluation
    and scaling-pass counts exactly; allocated times and coefficients bit for bit."""
    rng = np.random.default_rng(seed)
    probs = []
    for p in range(B):
        V = int(rng.integers(3, 14))
        wp = W.random_flier_path(1000 + seed * 100 + p, V)
        mask = np.zeros(V, np.uint8)
        vals = np.zeros((V, O.HALF, O.D))
        for v in range(V):
            vals[v, 0] = wp[v]
            if v == 0 or v == V - 1:
                mask[v] = (1 << (r + 1)) - 1  # makeStartOrEnd(position, r)
            else:
                mask[v] = 1
        times = O.estimate_times(wp)[0]
        probs.append((mask, vals, times))
    vtx_off = np.cumsum([0] + [len(m) for m, _, _ in probs]).astype(np.int32)
    P = ctx.L.default_params(derivative_to_optimize=r)
    got = ctx.time_alloc_batch(vtx_off, np.concatenate([m for m, _, _ in probs]), np.concatenate([v for _, v, _ in probs]),
                               np.concatenate([t for _, _, t in probs]), P)
    exact = True
    s0 = 0
    for p, (mask, vals, times) in enumerate(probs):
        ref = O.time_alloc(mask, vals, times, r, O.default_params(derivative_to_optimize=r))
        S = len(times)
        ok = (ref["nlopt_code"] == got["nlopt_code"][p] and ref["n_evals"] == got["n_evals"][p] and ref["n_scale_passes"] == got["n_scale_passes"][p]
              and np.array_equal(ref["times"], got["times"][s0:s0 + S]) and np.array_equal(ref["coef"], got["coef"][s0:s0 + S])
              and ref["final_cost"] == got["final_cost"][p])
        exact = exact and ok
        s0 += S
    return exact


def check_sampling(ctx, seed=1, B=12):
    rng = np.random.default_rng(seed)
    coefs, times, seg_off = [], [], [0]
    for p in range(B):
        V = int(rng.integers(2, 16))
        m, v, t = random_linear_problem(rng, V, kind=p % 3)
        c, _, _, _ = O.solve_linear(m, v, t, 2)
        coefs.append(c)
        times.append(t)
        seg_off.append(seg_off[-1] + V - 1)
    dt = 0.2
    counts, samples, full = ctx.sample_batch(np.array(seg_off, np.int32), np.concatenate(coefs), np.concatenate(times), dt, full=True)
    m0 = 0
    exact = True
    for p in range(B):
        ref, tns = O.sample(coefs[p], times[p], dt)
        assert counts[p] == len(ref), (p, counts[p], len(ref))  # sample indices bit-exact
        mine = full[m0:m0 + counts[p]]
        assert np.abs(mine[:, :3] - ref[:, :3]).max() <= SAMPLE_ATOL
        assert np.abs(mine - ref).max() <= 1e-6
        assert np.abs(samples[m0:m0 + counts[p], 3] - ref[:, 18]).max() <= 1e-9
        exact = exact and np.array_equal(mine, ref)
        m0 += counts[p]
    return exact


def check_evaluate(ctx, seed=2):
    rng = np.random.default_rng(seed)
    m, v, t = random_linear_problem(rng, 9, kind=1)
    c, _, _, _ = O.solve_linear(m, v, t, 2)
    tq = np.concatenate([[0.0, t[0], t.sum(), t.sum() + 1.0], rng.uniform(0, t.sum(), 50)])
    exact = True
    for k in range(5):
        out, ok = ctx.evaluate(c, t, tq, k)
        for i, tt in enumerate(tq):
            ref, rok = O.trajectory_evaluate(c, t, tt, k)
            assert ok[i] == rok
            assert np.abs(out[i] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
            exact = exact and np.array_equal(out[i], ref)
    return exact


def check_extrema_and_scaling(ctx, seed=3, B=6):
    rng = np.random.default_rng(seed)
    exact = True
    coefs, times, seg_off = [], [], [0]
    for p in range(B):
        V = int(rng.integers(3, 14))
        m, v, t = random_linear_problem(rng, V, kind=p % 3)
        t = t * 0.4  # short times -> limits are violated and scaling has work to do
        c, _, _, _ = O.solve_linear(m, v, t, 2)
        coefs.append(c)
        times.append(t)
        seg_off.append(seg_off[-1] + V - 1)
    allc, allt = np.concatenate(coefs), np.concatenate(times)
    mx = ctx.extrema(allc, allt)
    ref = np.concatenate([O.segment_maxima(coefs[p], times[p]) for p in range(B)])
    assert np.abs(mx - ref).max() <= 1e-9 * np.abs(ref).max()
    exact = exact and np.array_equal(mx, ref)
    c2, t2, passes, within = ctx.scale_times(np.array(seg_off, np.int32), allc, allt, O.DEFAULT_LIMITS)
    for p in range(B):
        rc, rt, rp, rw = O.scale_times(coefs[p], times[p])
        s0, s1 = seg_off[p], seg_off[p + 1]
        assert passes[p] == rp and bool(within[p]) == rw
        assert np.allclose(t2[s0:s1], rt, rtol=1e-12, atol=0)
        exact = exact and np.array_equal(t2[s0:s1], rt) and np.array_equal(c2[s0:s1], rc)
        # the stretched trajectory satisfies the limits (eth/trajectory.cpp:640: within 1 + 1e-3)
        post = ctx.extrema(c2[s0:s1], t2[s0:s1]).max(axis=0)
        lim = np.array(O.DEFAULT_LIMITS)[[0, 2, 4, 1, 3, 5, 6, 7, 8]]
        assert np.all(post <= lim * (1 + 1e-3) + 1e-12)
    return exact


def check_max_magnitude(ctx, seed=11, B=9):
    """computeMaximumOfMagnitude (lin_impl.h:477-508): Extremum {time, value, segment} per trajectory, derivatives 1..4, against the
    oracle; plus the identity the reference's own tests use (test_utils.h:40-50): analytic maximum >= densely sampled maximum."""
    rng = np.random.default_rng(seed)
    coefs, times, seg_off = [], [], [0]
    for p in range(B):
        V = int(rng.integers(2, 14))
        m, v, t = random_linear_problem(rng, V, kind=p % 3)
        c, _, _, _ = O.solve_linear(m, v, t, 2)
        coefs.append(c)
        times.append(t)
        seg_off.append(seg_off[-1] + V - 1)
    allc, allt = np.concatenate(coefs), np.concatenate(times)
    exact = True
    for k in (1, 2, 3, 4):
        tt, vv, ii = ctx.max_magnitude(np.array(seg_off, np.int32), allc, allt, k)
        for p in range(B):
            rt, rv, ri = O.max_magnitude(coefs[p], times[p], k)
            assert ii[p] == ri and abs(vv[p] - rv) <= 1e-9 * max(1.0, abs(rv)) and abs(tt[p] - rt) <= 1e-9 * max(1.0, abs(rt)), (k, p)
            exact = exact and vv[p] == rv and tt[p] == rt
            # sampled maximum over the whole trajectory never exceeds the analytic one (up to rounding)
            smax = 0.0
            for s in range(len(times[p])):
                ts = np.linspace(0.0, times[p][s], 41)
                val = np.zeros((41, 4))
                for d in range(4):
                    cd = coefs[p][s, d]
                    for j in range(9, k - 1, -1):
                        f = 1.0
                        for q in range(k):
                            f *= (j - q)
                        val[:, d] = val[:, d] * ts + f * cd[j]
                smax = max(smax, float(np.sqrt((val ** 2).sum(axis=1)).max()))
            assert smax <= vv[p] * (1 + 1e-9) + 1e-12, (k, p, smax, vv[p])
    return exact


def check_objectives(ctx, seed=5, K=24):
    """Objective functions of the time-allocation methods 0/1/3/4 (nl_impl.h:567-722) at K candidates, soft constraints through
    computeMaximumOfMagnitude: total and the three terms against the oracle."""
    rng = np.random.default_rng(seed)
    exact = True
    con_deriv = [1, 2, 3, 1, 2, 3]           # the node adds (dimension, derivative) pairs; only the derivative matters (lin_impl.h:407-409)
    con_value = [4.0, 2.0, 20.0, 2.0, 1.0, 20.0]
    for kind, V in ((0, 6), (1, 11), (2, 4)):
        m, v, t = random_linear_problem(rng, V, kind=kind)
        S = V - 1
        c0, _, dp0, _ = O.solve_linear(m, v, t, 2)
        n_free = dp0.shape[1]
        for method in (0, 1, 3, 4):
            xt = t[None, :] * np.exp(rng.uniform(-0.3, 0.3, size=(K, S)))
            if method >= 3:
                xf = dp0.reshape(1, -1) * (1.0 + 0.05 * rng.standard_normal((K, 4 * n_free)))
                x = np.concatenate([xt, xf], axis=1)
            else:
                x = xt
            for soft in (True, False):
                tot, parts = ctx.objective(m, v, 2, method, x, 500.0, soft, 100.0, con_deriv, con_value)
                rtot, rparts = O.objective(m, v, 2, method, x, 500.0, soft, 100.0, con_deriv, con_value)
                assert np.allclose(tot, rtot, rtol=1e-9, atol=0), (kind, method, soft)
                assert np.allclose(parts, rparts, rtol=1e-9, atol=1e-300)
                exact = exact and np.array_equal(tot, rtot) and np.array_equal(parts, rparts)
        # the unperturbed solution of solveLinear is a stationary point of the trajectory cost in the free derivatives:
        # method 3 at (t, d_p*) has the same trajectory cost as method 0 at t
        x0 = t[None, :]
        x3 = np.concatenate([x0, dp0.reshape(1, -1)], axis=1)
        a, pa = ctx.objective(m, v, 2, 0, x0, 500.0, False)
        b, pb = ctx.objective(m, v, 2, 3, x3, 500.0, False)
        assert abs(pa[0, 0] - pb[0, 0]) <= 1e-9 * abs(pa[0, 0])
    return exact


def check_derivative_free_time_allocation(ctx):
    """Methods 0/1/3/4 through the reference-shaped class: the search only ever accepts improvements of the reference's objective, stays
    inside the bounds, and returns the segments of its final point (checked against the oracle's objective at that point)."""
    import mrs_uav_trajectory_generation_b200.api as A

    wps = [(0, 0, 1, 0), (2, 0.5, 1.2, 0.1), (4, -0.5, 1.0, 0.2), (6, 0.5, 1.1, 0.1), (8, 0, 1, 0)]
    verts = []
    for i, w in enumerate(wps):
        v = A.Vertex(4)
        if i in (0, len(wps) - 1):
            v.makeStartOrEnd(np.array(w, dtype=np.float64), 2)
        else:
            v.addConstraint(0, np.array(w, dtype=np.float64))
        verts.append(v)
    t0 = np.full(len(wps) - 1, 1.5)
    for method in (0, 1, 3, 4):
        P = A.NonlinearOptimizationParameters()
        P.time_alloc_method = method
        P.max_iterations = 12
        P.f_rel = 1e-4
        opt = A.PolynomialOptimizationNonLinear(4, P, ctx=ctx)
        assert opt.setupFromVertices(verts, t0, 2)
        for dim in range(3):
            opt.addMaximumMagnitudeConstraint(dim, 1, 3.0)
            opt.addMaximumMagnitudeConstraint(dim, 2, 2.0)
        code = opt.optimize()
        assert code in (3, 4, 5)
        dfo = opt._dfo
        _, mask, vals = A.pack_vertices([verts])
        cd = [c[1] for c in dfo.constraints]
        cv = [c[2] for c in dfo.constraints]
        x0 = t0 if method < 3 else None
        rt, rp = O.objective(mask, vals, 2, method, dfo.x[None, :], P.time_penalty, True, P.soft_constraint_weight, cd, cv)
        assert abs(rt[0] - dfo.cost) <= 1e-9 * abs(rt[0])
        if x0 is not None:
            r0, _ = O.objective(mask, vals, 2, method, x0[None, :], P.time_penalty, True, P.soft_constraint_weight, cd, cv)
            assert dfo.cost <= r0[0]
        assert np.all(dfo.x[: len(t0)] >= 0.01)
        traj = opt.getTrajectory()
        assert traj.coef.shape == (len(t0), 4, 10) and np.array_equal(traj.times, dfo.x[: len(t0)])
    return True


def check_two_lanes(lib, n=96):
    """tg_optimize_batch cuts a large batch in two and runs the halves on two lanes (two host threads, two sets of streams); the
    merged results and ragged outputs must equal the single-lane ones bit for bit, and the oracle's."""
    import os

    from mrs_uav_trajectory_generation_b200 import Context
    from mrs_uav_trajectory_generation_b200 import workloads as W

    wp_off, wp = W.random_flier_paths_fast(n, first_index=21)
    outs = []
    for lanes, min_batch in ((1, 1 << 30), (2, 2)):
        old = {k: os.environ.get(k) for k in ("TG_LANES", "TG_LANE_MIN_BATCH")}
        os.environ["TG_LANES"], os.environ["TG_LANE_MIN_BATCH"] = str(lanes), str(min_batch)
        try:
            ctx = Context(lib, 0)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        res, totals = ctx.optimize_batch(wp_off, wp, None, None, ctx.L.default_params())
        outs.append((res, totals, ctx.fetch_outputs(), ctx.counters()))
    (r1, t1, o1, c1), (r2, t2, o2, c2) = outs
    assert np.array_equal(t1, t2)
    for f in r1.dtype.names:
        assert np.array_equal(r1[f], r2[f]), f
    for k in ("seg_off", "wp", "times", "coef", "smp_off", "samples"):
        assert np.array_equal(o1[k], o2[k]), k
    assert c1["solves"] == c2["solves"] and c1["root_finds"] == c2["root_finds"]
    ref = O.optimize_batch(wp_off, wp, params=O.default_params(), cap_wp=1400, cap_samples=6000)
    for p in (0, n // 2 - 1, n // 2, n - 1):  # either side of the cut
        s0, s1 = o2["seg_off"][p], o2["seg_off"][p + 1]
        nw = ref["res"][p].n_waypoints
        assert s1 - s0 == nw - 1 and np.array_equal(o2["coef"][s0:s1], ref["coeffs"][p][: nw - 1])
    return True


def compare_optimize(ctx, wp_off, wp, stop_at=None, init=None, params_kw=None, cap_wp=1400, cap_samples=6000):
    """Runs the full optimize() pipeline on both sides; asserts parity; returns (results, bit_exact, worst_coef_err)."""
    params_kw = params_kw or {}
    ref = O.optimize_batch(wp_off, wp, stop_at=stop_at, init=init, params=O.default_params(**params_kw), cap_wp=cap_wp,
                           cap_samples=cap_samples)
    res, totals = ctx.optimize_batch(wp_off, wp, stop_at, init, ctx.L.default_params(**params_kw))
    out = ctx.fetch_outputs()
    B = len(wp_off) - 1
    exact = True
    worst = 0.0
    ints = ("status", "success", "nlopt_code", "n_evals", "rounds", "safe", "n_waypoints", "n_samples", "n_scale_passes",
            "total_solves", "total_root_calls", "total_evals")
    for p in range(B):
        r, g = ref["res"][p], res[p]
        assert not r.overflow
        for k in ints:  # verdicts, subdivision counts, segment counts, sample counts: exact
            assert getattr(r, k) == g[k], (p, k, getattr(r, k), g[k])
        s0, s1 = out["seg_off"][p], out["seg_off"][p + 1]
        m0, m1 = out["smp_off"][p], out["smp_off"][p + 1]
        S = s1 - s0
        assert S == max(g["n_waypoints"] - 1, 0) and m1 - m0 == g["n_samples"]  # a path without a trajectory (status 5, 6) owns nothing
        if not g["success"]:
            continue
        rt, rc = ref["times"][p, :S], ref["coeffs"][p, :S]
        rs, rw = ref["samples"][p, :m1 - m0], ref["wp"][p, :S + 1]
        assert np.allclose(out["times"][s0:s1], rt, rtol=1e-12, atol=0)
        worst = max(worst, coef_rel_err(out["coef"][s0:s1], rc, rt))
        assert np.abs(out["samples"][m0:m1, :3] - rs[:, :3]).max() <= SAMPLE_ATOL
        assert np.abs(out["wp"][s0 + p:s1 + p + 1] - rw).max() <= 1e-12
        assert abs(g["max_dev"] - r.max_dev) <= 1e-9 and abs(g["final_cost"] - r.final_cost) <= 1e-9 * max(1.0, abs(r.final_cost))
        exact = (exact and np.array_equal(out["times"][s0:s1], rt) and np.array_equal(out["coef"][s0:s1], rc)
                 and np.array_equal(out["samples"][m0:m1], rs) and np.array_equal(out["wp"][s0 + p:s1 + p + 1], rw)
                 and g["max_dev"] == r.max_dev and g["final_cost"] == r.final_cost and g["baca_total"] == r.baca_total)
    assert worst <= COEF_RTOL, worst
    if B <= 2048:
        check_streamed_equals_fetched(ctx, wp_off, wp, stop_at, init, params_kw, res, out)
    return res, out, exact, worst


def check_streamed_equals_fetched(ctx, wp_off, wp, stop_at, init, params_kw, res, out):
    """tg_optimize_batch_streamed (samples copied out while later rounds run, completion order) delivers the same results and, path by
    path, the same sample rows as tg_optimize_batch + tg_fetch_outputs; a buffer that is too small is refused without losing the batch."""
    tot = int(res["n_samples"].sum())
    buf = np.full((tot + 7, 4), np.nan)
    res2, totals2, begin = ctx.optimize_batch_streamed(wp_off, wp, buf, stop_at, init, ctx.L.default_params(**params_kw))
    assert res2.tobytes() == res.tobytes() and totals2[1] == tot
    used = np.zeros(tot + 7, bool)
    for p in range(len(res)):
        n, m0 = int(res["n_samples"][p]), int(out["smp_off"][p])
        assert np.array_equal(buf[begin[p]:begin[p] + n], out["samples"][m0:m0 + n]), p
        assert not used[begin[p]:begin[p] + n].any()
        used[begin[p]:begin[p] + n] = True
    assert used[:tot].all() and not used[tot:].any() and np.isnan(buf[tot:]).all()
    if tot > 1:
        try:
            ctx.optimize_batch_streamed(wp_off, wp, np.empty((tot - 1, 4)), stop_at, init, ctx.L.default_params(**params_kw))
            raise AssertionError("a buffer one row short was accepted")
        except Exception as e:  # TgError: TG_ERR_CAPACITY
            assert "too small" in str(e), e
        again = ctx.fetch_outputs(want=("smp_off", "samples"))  # the batch result is complete all the same
        assert np.array_equal(again["samples"], out["samples"])


def check_random_flier(ctx, B, first_index=0, **params_kw):
    wp_off, wp = W.random_flier_paths(B, first_index=first_index)
    return compare_optimize(ctx, wp_off, wp, params_kw=params_kw)


def check_roots_against_reference_vectors(ctx):
    """The device Jenkins-Traub (tg_poly.cuh) against tests/golden/rpoly_reference.npz -- 600 polynomials whose zeros were computed by the
    REFERENCE's own rpoly_ak1.cpp (compiled unmodified, tests/golden/gen_golden.py): every zero, in the reference's order, bit for bit."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpoly_reference.npz"))
    n = len(z["n_coeffs"])
    assert z["coeffs"].shape[1] == 16
    re, im, nr = ctx.test_find_roots(z["coeffs"], z["n_coeffs"])
    bad = 0
    for i in range(n):
        k = int(z["n_roots"][i])
        if nr[i] != k or not np.array_equal(re[i, :k], z["re"][i, :k]) or not np.array_equal(im[i, :k], z["im"][i, :k]):
            bad += 1
    assert bad == 0, bad
    return n


def check_roots_adversarial(ctx, n=1500, seed=77):
    """Random polynomials of every degree 1..15 with wide coefficient ranges, exact zeros at the origin, vanishing leading coefficients,
    clustered zeros: device zeros == oracle zeros (the oracle itself is pinned to the reference's file)."""
    rng = np.random.default_rng(seed)
    C16 = np.zeros((n, 16))
    nc = np.zeros(n, np.int32)
    for i in range(n):
        k = int(rng.integers(2, 17))
        c = rng.standard_normal(k) * np.exp(rng.uniform(-5, 5, k))
        if i % 5 == 0:
            c[: int(rng.integers(1, 3))] = 0.0
        if i % 7 == 0:
            c[-int(rng.integers(1, 3)):] = 0.0
        if i % 11 == 0 and k > 5:
            roots = np.concatenate([np.full(3, rng.uniform(-2, 2)), rng.uniform(-3, 3, k - 4)])
            c = np.poly(roots)[::-1][:k]
        C16[i, : len(c)] = c
        nc[i] = len(c)
    re, im, nr = ctx.test_find_roots(C16, nc)
    for i in range(n):
        r2, i2, ok = O.find_roots(C16[i, : nc[i]])
        assert nr[i] == len(r2) and np.array_equal(re[i, : nr[i]], r2) and np.array_equal(im[i, : nr[i]], i2), i
    return True


def check_heading_override(ctx, B=16):
    """getTrajectoryReference with override_heading_atan2 (node.cpp:1586-1599): heading = direction to the next sample, previous heading
    when the step is shorter than 0.05 m (a path with a stop_at waypoint produces such steps), last sample keeps getYaw()."""
    paths, stops = [], []
    for p in range(B):
        path = W.random_flier_path(4000 + p, 6 + p % 5)
        st = np.zeros(len(path), np.uint8)
        if p % 2 == 0:
            st[len(path) // 2] = 1  # the vehicle stops there: consecutive samples closer than 0.05 m
        paths.append(path)
        stops.append(st)
    wp_off = np.concatenate([[0], np.cumsum([len(q) for q in paths])]).astype(np.int32)
    res, out, exact, worst = compare_optimize(ctx, wp_off, np.concatenate(paths), stop_at=np.concatenate(stops), params_kw=dict(override_heading_atan2=1))
    assert exact
    # the override really happened: headings follow the direction of travel, and some short steps reused the previous heading
    smp, off = out["samples"], out["smp_off"]
    reused = 0
    for p in range(B):
        s = smp[off[p]: off[p + 1]]
        d = s[1:, :2] - s[:-1, :2]
        dist = np.hypot(d[:, 1], d[:, 0])
        far = dist >= 0.05
        assert np.allclose(s[:-1, 3][far], np.arctan2(d[:, 1], d[:, 0])[far], rtol=0, atol=1e-12)
        reused += int((~far[1:]).sum())
    assert reused > 0
    return True


def check_several_groups(library, device=0, B=40):
    """A batch larger than the per-group segment budget is cut into several groups in round 0 and again in every subdivision round
    (TG_SEG_BUDGET lowers the 2 M-segment budget so that 40 paths already need ~8 groups): results must not depend on the cut."""
    import os

    from mrs_uav_trajectory_generation_b200 import Context

    old = os.environ.get("TG_SEG_BUDGET")
    os.environ["TG_SEG_BUDGET"] = "48"
    try:
        ctx = Context(library, device)
    finally:
        if old is None:
            os.environ.pop("TG_SEG_BUDGET", None)
        else:
            os.environ["TG_SEG_BUDGET"] = old
    paths, stops = [], []
    for p in range(B):
        path = W.random_flier_path(8000 + p, 3 + p % 9)
        paths.append(path)
        stops.append(np.array([(p + i) % 7 == 0 for i in range(len(path))], np.uint8))
    wp_off = np.concatenate([[0], np.cumsum([len(q) for q in paths])]).astype(np.int32)
    res, out, exact, worst = compare_optimize(ctx, wp_off, np.concatenate(paths), stop_at=np.concatenate(stops))
    assert exact and res["success"].all() and res["rounds"].max() >= 1
    return True


def check_degenerate_inputs(ctx):
    """Inputs at the edge of the contract, every one against the oracle (verdicts, counts and outputs bit for bit):
    the shortest path (two waypoints), a repeated waypoint (zero-length segment: rejected by the length filter), waypoints 1 km apart
    (six subdivision rounds), 6 cm apart, a single waypoint ("the path is empty", node.cpp:676-681 -> status 5), a non-finite waypoint
    (checkNaN, node.cpp:1896-1900 -> status 6), all mixed with ordinary paths in one ragged batch; an empty batch; a negative count."""
    two = np.array([[0, 0, 1, 0], [3, 0, 1, 0.0]])
    dup = np.array([[0, 0, 1, 0], [3, 0, 1, 0.0], [3, 0, 1, 0.0], [6, 1, 1, 0.5]])
    far = np.array([[0, 0, 1, 0], [1000, 0, 1, 0.0], [1000, 500, 30, 3.0]])
    near = np.array([[0, 0, 1, 0], [0.06, 0, 1, 0.0], [0.12, 0.01, 1, 0.0]])
    one = np.array([[0, 0, 1, 0.0]])
    nanp = np.array([[0, 0, 1, 0], [np.nan, 0, 1, 0.0], [3, 3, 1, 0]])
    infp = np.array([[0, 0, 1, 0], [1, 0, 1, np.inf], [3, 3, 1, 0]])
    usual = [W.random_flier_path(7000 + i, 4 + i) for i in range(3)]
    paths = [two, usual[0], dup, one, far, nanp, usual[1], near, infp, usual[2]]
    wp_off = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.int32)
    res, out, exact, worst = compare_optimize(ctx, wp_off, np.concatenate(paths), cap_wp=400, cap_samples=20000)
    assert exact
    assert list(res["status"]) == [0, 0, 2, 5, 0, 6, 0, 0, 6, 0], list(res["status"])
    assert list(res["success"]) == [1, 1, 0, 0, 1, 0, 1, 1, 0, 1]
    assert res["rounds"][4] == 6 and res["n_samples"][3] == 0 and res["n_samples"][5] == 0
    # an empty batch is not an error and returns nothing; a negative waypoint count flags that problem only
    r0, t0 = ctx.optimize_batch(np.array([0], np.int32), np.zeros((0, 4)), None, None, ctx.L.default_params())
    assert len(r0) == 0 and tuple(t0) == (0, 0) or len(r0) == 0
    r1, _ = ctx.optimize_batch(np.array([0, 3, 2], np.int32), np.concatenate([usual[0][:3]]), None, None, ctx.L.default_params())
    assert r1["success"][0] == 1 and r1["success"][1] == 0 and r1["status"][1] == 5
    return True


def check_acceptance_rejects(ctx, B=48):
    """The acceptance logic of findTrajectory (node.cpp:1138-1149, 1178-1199): with max_len_factor = 1.9 some of these paths are
    rejected as 'too long' (FindStatus 2), with min_len_factor = 1.9 most as 'too short' (3); a rejected findTrajectory makes
    optimize() fail for that path without touching its neighbours.  Status 1 (an NLopt code outside {>= 1 except 6, -1}) cannot
    occur on either side: every failure of the optimiser surfaces as -1 (nlopt::opt throws, nl_impl.h:190-208) and MAXTIME (6)
    needs a wall clock, which neither the oracle nor the kernels have."""
    seen = set()
    for kw in (dict(max_len_factor=1.9), dict(min_len_factor=1.8), dict(max_len_factor=1.75, min_len_factor=1.65)):
        res, out, exact, worst = check_random_flier(ctx, B, first_index=700, **kw)
        assert exact
        st = set(int(x) for x in res["status"])
        assert (res["success"][res["status"] != 0] == 0).all() and (res["success"][res["status"] == 0] == 1).all()
        seen |= st
    assert seen == {0, 2, 3}, seen
    return True


def check_fixtures(ctx):
    """F1a (the reference tests' 4-waypoint path + prepended start) and F1b (10-waypoint zig-zag)."""
    paths = [W.F1A_WAYPOINTS, W.F1B_WAYPOINTS]
    wp_off = np.array([0, len(paths[0]), len(paths[0]) + len(paths[1])], np.int32)
    wp = np.concatenate(paths)
    init = np.stack([W.init14(W.F1A_INIT_HEADING), W.init14(W.F1B_INIT_HEADING)])
    res, out, exact, worst = compare_optimize(ctx, wp_off, wp, init=init)
    assert res["success"].all()
    return res, out, exact


def check_mixed_batch(ctx, seed=5):
    """Ragged batch: different waypoint counts, stop_at flags, initial states with non-zero derivatives, heading wraps,
    a two-waypoint path (S = 1: Mellinger gradient is defined as zero, nl_impl.h:264-271)."""
    rng = np.random.default_rng(seed)
    paths, stops, inits = [], [], []
    for p in range(14):
        nwp = [2, 3, 5, 8, 11, 14, 20][p % 7]
        path = W.random_flier_path(1000 + p, nwp)
        if p % 3 == 0:
            path[:, 3] += 5.5  # headings beyond pi: exercises sradians::unwrap / radians::interp
        st = np.zeros(nwp, np.uint8)
        if p % 4 == 1 and nwp > 3:
            st[nwp // 2] = 1
        paths.append(path)
        stops.append(st)
        if p % 2 == 0:
            inits.append(W.init14(path[0, 3] + 0.3, vel=rng.uniform(-0.5, 0.5, 4), acc=rng.uniform(-0.2, 0.2, 4), jerk=rng.uniform(-0.1, 0.1, 4)))
        else:
            z = np.zeros(14)
            inits.append(z)
    wp_off = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.int32)
    return compare_optimize(ctx, wp_off, np.concatenate(paths), stop_at=np.concatenate(stops), init=np.stack(inits))


def check_config2(ctx, B=64, r=2):
    """BASELINE config 2: linear solve at the Euclidean times + sampling, no time allocation, no deviation loop."""
    wp_off, wp = W.random_flier_paths(B, first_index=5000)
    return compare_optimize(ctx, wp_off, wp, params_kw=dict(run_time_alloc=0, check_deviation=0, derivative_to_optimize=r))


def check_sweep(ctx, K=300, seed=7):
    """BASELINE config 5: K candidate time vectors of one problem; best index and cost must match the oracle."""
    rng = np.random.default_rng(seed)
    path = W.random_flier_path(0)
    V = len(path)
    m = np.ones(V, np.uint8)
    m[0] = 7
    m[-1] = 7
    v = np.zeros((V, 5, 4))
    v[:, 0, :] = path
    T0, _ = O.estimate_times(path)
    cand = np.maximum(0.01, T0[None, :] * np.exp(rng.uniform(-0.5, 0.5, (K, V - 1))))
    costs, bi, bc = ctx.sweep_costs(m, v, cand, r=2)
    ref = O.sweep_costs(m, v, 2, cand)
    assert np.allclose(costs, ref, rtol=1e-9, atol=0)
    assert bi == int(np.argmin(ref)) and bc == ref.min()
    return np.array_equal(costs, ref)


def check_sweep_best(ctx, K=1001, seed=8, n_ctx=3):
    """tg_sweep_best: the candidates sharded over several contexts of one process (here: contexts on the same device); the first
    minimum of the whole list, its cost bit for bit and its time vector must be what a serial scan of the oracle's costs gives, also
    when the minimum is duplicated in a later shard."""
    from mrs_uav_trajectory_generation_b200 import Context

    rng = np.random.default_rng(seed)
    path = W.random_flier_path(1)
    V = len(path)
    m = np.ones(V, np.uint8)
    m[0] = m[-1] = 7
    v = np.zeros((V, 5, 4))
    v[:, 0, :] = path
    T0, _ = O.estimate_times(path)
    cand = np.maximum(0.01, T0[None, :] * np.exp(rng.uniform(-0.5, 0.5, (K, V - 1))))
    ref = O.sweep_costs(m, v, 2, cand)
    first = int(np.argmin(ref))
    cand[(first + K // 2) % K] = cand[first]  # a tie in another shard: the lower index must win
    ref = O.sweep_costs(m, v, 2, cand)
    ctxs = [ctx] + [Context(ctx.L, ctx.device if hasattr(ctx, "device") else 0) for _ in range(n_ctx - 1)]
    bc, bi, bt = Context.sweep_best(ctxs, m, v, cand, r=2)
    assert bi == int(np.argmin(ref)) and bc == ref.min() and np.array_equal(bt, cand[bi])
    bc1, bi1, _ = Context.sweep_best(ctxs[:1], m, v, cand, r=2)
    assert (bc1, bi1) == (bc, bi)
    return True


def check_scaling_multi_pass(ctx, seed=17, B=24):
    """The global check's certificates (tg_bound.cuh) and their completion before a further pass: with the reference's
    tolerance a second pass practically never happens, so the tolerance is lowered (test hooks on both sides) to force up
    to 20 passes; times, coefficients, pass counts and verdicts must still equal the oracle's bit for bit."""
    rng = np.random.default_rng(seed)
    ok = True
    passes_seen = set()
    try:
        for tol in (-0.2, -0.6):
            O.set_scale_tolerance(tol)
            ctx.test_set_scale_tolerance(tol)
            segs = [int(rng.integers(1, 7)) for _ in range(B)]
            seg_off = np.cumsum([0] + segs).astype(np.int32)
            coef = rng.standard_normal((seg_off[-1], 4, 10)) * np.exp(rng.uniform(-3, 1, (seg_off[-1], 1, 10)))
            times = np.exp(rng.uniform(-1, 1.5, seg_off[-1]))
            lim = np.array(O.DEFAULT_LIMITS) * np.exp(rng.uniform(-1, 1, 9))
            c2, t2, passes, within = ctx.scale_times(seg_off, coef, times, lim)
            for p in range(B):
                s0, s1 = seg_off[p], seg_off[p + 1]
                rc, rt, rp, rw = O.scale_times(coef[s0:s1], times[s0:s1], lim)
                ok = ok and np.array_equal(c2[s0:s1], rc) and np.array_equal(t2[s0:s1], rt) and passes[p] == rp and bool(within[p]) == bool(rw)
                passes_seen.add(int(rp))
    finally:
        O.set_scale_tolerance(1e-3)
        ctx.test_set_scale_tolerance(1e-3)
    return ok and max(passes_seen) > 1


def check_path_side_steps(ctx, seed=31, B=40):
    """preprocessPath, findTrajectoryFallback and getWaypointInTrajectoryIdxs against the oracle, bit for bit."""
    rng = np.random.default_rng(seed)
    paths, stops = [], []
    for p in range(B):
        V = int(rng.integers(2, 24))
        wp = W.random_flier_path(3000 + seed * 100 + p, V)
        if p % 3 == 0:  # near-duplicate and collinear waypoints: work for the min-distance filter and the straightener
            k = int(rng.integers(1, V)) if V > 2 else 1
            wp[k] = wp[k - 1] + np.array([0.01, 0.0, 0.0, 0.0]) * (p % 2)
            if V > 4:
                wp[2] = 0.5 * (wp[1] + wp[3])
        if p % 5 == 0:
            wp[:, 3] = rng.uniform(-7, 7, V)  # headings that need wrapping / unwrapping
        st = (rng.uniform(size=V) < 0.15).astype(np.uint8)
        paths.append(wp)
        stops.append(st)
    wp_off = np.cumsum([0] + [len(w) for w in paths]).astype(np.int32)
    wp = np.concatenate(paths)
    stop = np.concatenate(stops)
    ok = True
    for straight in (False, True):
        off, owp, ostop = ctx.preprocess_paths(wp_off, wp, stop, 0.05, straight, 0.05, 0.1)
        for p in range(B):
            rw, rs = O.preprocess_path(paths[p], stops[p], 0.05, straight, 0.05, 0.1)
            ok = ok and np.array_equal(owp[off[p]:off[p + 1]], rw) and np.array_equal(ostop[off[p]:off[p + 1]], rs)
    lim = np.array(O.DEFAULT_LIMITS) * np.array([0.75, 0.75, 0.75, 0.75, 1, 1, 1, 1, 1])  # speed / acceleration factors applied by the caller
    soff, smp = ctx.fallback_sample_batch(wp_off, wp, stop, lim, 0.2, 2.0)
    for p in range(B):
        ref = O.fallback_sample(paths[p], stops[p], lim, 0.2, 2.0)
        ok = ok and np.array_equal(smp[soff[p]:soff[p + 1]], ref)
    idx = ctx.waypoint_idxs_batch(soff, smp, wp_off, wp)
    for p in range(B):
        ok = ok and np.array_equal(idx[p], O.waypoint_idxs(smp[soff[p]:soff[p + 1]], paths[p]))
    return ok


def geometric_predicate(samples, waypoints, pos_tol=0.5, hdg_tol=0.2):
    """The reference tests' acceptance check (test/include/get_path_test.h:45-68): every input waypoint is approached
    within 0.5 m / 0.2 rad by some sample, in order."""
    idx = 0
    for w in waypoints:
        found = False
        while idx < len(samples):
            s = samples[idx]
            dh = abs((s[3] - w[3] + np.pi) % (2 * np.pi) - np.pi)
            if np.linalg.norm(s[:3] - w[:3]) < pos_tol and dh < hdg_tol:
                found = True
                break
            idx += 1
        if not found:
            return False
    return True
