"""Independent checks of the oracle's linear QP (the reference's tests pin no numbers): a dense numpy
mirror of lin_impl.h written from the paper's formulas (not from the oracle's code), and the identities of the upstream
unit-test helpers (eth/test_utils.h:40-59)."""
import math

import numpy as np
import pytest

import oracle_lib as O
import parity_checks as PC

N, HALF, D = 10, 5, 4


def base(k, i):  # i!/(i-k)!  (eth/polynomial.cpp:155-170)
    return float(math.prod(range(i - k + 1, i + 1))) if i >= k else 0.0


def numpy_mirror(mask, vals, times, r):
    """Unconstrained QP of Richter et al. with dense matrices: returns coefficients [S,4,10] and the cost."""
    S, V = len(times), len(times) + 1
    Hs, Ainvs, Qs = [], [], []
    for T in times:
        A = np.zeros((N, N))
        for k in range(HALF):
            A[k, k] = base(k, k)
            for j in range(k, N):
                A[HALF + k, j] = base(k, j) * T ** (j - k)
        Q = np.zeros((N, N))
        for i in range(r, N):
            for j in range(r, N):
                e = i + j - 2 * r + 1
                Q[i, j] = 2.0 * base(r, i) * base(r, j) * T ** e / e
        Ai = np.linalg.inv(A)
        Hs.append(Ai.T @ Q @ Ai)
        Ainvs.append(Ai)
        Qs.append(Q)
    fixed = [(v, k) for v in range(V) for k in range(HALF) if (mask[v] >> k) & 1]
    free = [(v, k) for v in range(V) for k in range(HALF) if not (mask[v] >> k) & 1]
    col = {s: i for i, s in enumerate(fixed + free)}
    n, nf = len(col), len(fixed)
    R = np.zeros((n, n))
    for s in range(S):
        idx = [col[(s + a // HALF, a % HALF)] for a in range(N)]
        R[np.ix_(idx, idx)] += Hs[s]
    coef = np.zeros((S, D, N))
    cost = 0.0
    for d in range(D):
        df = np.array([vals[v, k, d] for v, k in fixed])
        dp = np.linalg.solve(R[nf:, nf:], -R[nf:, :nf] @ df) if len(free) else np.zeros(0)
        full = np.concatenate([df, dp])
        for s in range(S):
            idx = [col[(s + a // HALF, a % HALF)] for a in range(N)]
            coef[s, d] = Ainvs[s] @ full[idx]
            cost += 0.5 * coef[s, d] @ Qs[s] @ coef[s, d]
    return coef, cost, R


@pytest.mark.parametrize("r", [2, 3, 4])
def test_oracle_agrees_with_dense_numpy_mirror(oracle, r):
    rng = np.random.default_rng(40 + r)
    for kind in range(3):
        V = int(rng.integers(3, 12))
        mask, vals, times = PC.random_linear_problem(rng, V, kind)
        mask = np.asarray(mask)
        mask[0] |= (1 << (r + 1)) - 1  # well posed for every r: both ends fix derivatives 0..r
        mask[-1] |= (1 << (r + 1)) - 1
        coef, cost, _, _ = O.solve_linear(mask, vals, times, r)
        ref, ref_cost, R = numpy_mirror(mask, vals, times, r)
        # the formulation is ill conditioned (cond(Rpp) up to 1e10): compare at the self-noise level
        assert PC.coef_rel_err(coef, ref, times) < 1e-5
        assert abs(cost - ref_cost) <= 1e-6 * max(1.0, abs(ref_cost))
        Ro = O.dense_R(mask, vals, times, r)
        assert np.abs(Ro - R).max() <= 1e-6 * np.abs(R).max()


def test_segment_matrices_identities(oracle):
    for T, tol in ((0.05, 1e-4), (0.7, 1e-8), (3.0, 1e-8)):  # A^-1 has entries ~ T^-9: 1e13 at T = 0.05
        A, Ainv, Q = O.segment_matrices(T, 2)
        assert np.abs(A @ Ainv - np.eye(N)).max() < tol
        assert np.allclose(Q, Q.T)


def test_constraints_continuity_and_cost_identity(oracle):
    rng = np.random.default_rng(77)
    V, r = 9, 2
    mask, vals, times = PC.random_linear_problem(rng, V, 0)
    coef, cost, _, _ = O.solve_linear(mask, vals, times, r)
    fact = lambda k: np.array([base(k, j) for j in range(N)])
    pw = np.arange(N)
    for s in range(V - 1):
        for k in range(HALF):
            end = (coef[s] * fact(k) * times[s] ** np.clip(pw - k, 0, None) * (pw >= k)).sum(axis=1)
            start = coef[s][:, k] * base(k, k)
            if (mask[s] >> k) & 1:
                assert np.abs(start - vals[s, k]).max() < 1e-9 * max(1.0, np.abs(vals[s, k]).max())  # fixed constraints are reproduced
            if s + 1 < V - 1:
                nxt = coef[s + 1][:, k] * base(k, k)
                assert np.abs(end - nxt).max() <= 1e-6 * max(1.0, np.abs(end).max())  # C^4 continuity (lin_impl.h:202-220)
    # analytic cost vs numeric integral of |p^(r)|^2 (eth/test_utils.h:52-59)
    num = 0.0
    for s in range(V - 1):
        t = np.linspace(0.0, times[s], 4001)
        for d in range(D):
            dr = sum(coef[s, d, j] * base(r, j) * t ** (j - r) for j in range(r, N))
            num += np.trapezoid(dr * dr, t)
    assert abs(cost - num) <= 2e-4 * max(1.0, abs(num))  # computeCost carries the reference's factor: 0.5 * c^T Q c with Q = 2 * integral


def test_analytic_maximum_dominates_sampled_maximum(oracle):
    """eth/test_utils.h:40-50: the Jenkins-Traub extremum is never below a dense sampling of the same magnitude."""
    r = O.optimize_path(np.array([[0, 0, 5, 0], [1, 1, 5, 0.3], [3, 0, 5.2, 0.1], [4, 2, 5, 0.5], [6, 2, 4.9, 0.2]], float),
                        params=O.default_params(check_deviation=0))
    coef, times = r["coeffs"], r["times"]
    mx = O.segment_maxima(coef, times)
    for s in range(len(times)):
        t = np.linspace(0.0, times[s], 2001)
        for k in (1, 2, 3):
            der = [sum(coef[s, d, j] * base(k, j) * t ** (j - k) for j in range(k, N)) for d in range(D)]
            assert mx[s, k - 1] >= np.sqrt(der[0] ** 2 + der[1] ** 2).max() * (1 - 1e-9)
            assert mx[s, 3 + k - 1] >= np.abs(der[2]).max() * (1 - 1e-9)
            assert mx[s, 6 + k - 1] >= np.abs(der[3]).max() * (1 - 1e-9)
