// tg_bound.cuh -- rigorous upper bounds on the per-segment maxima, used to skip Jenkins-Traub runs whose result cannot
// change a decision.
//
// scaleSegmentTimesToMeetConstraints (eth/trajectory.cpp:598-692) ends every pass with a GLOBAL check: the maxima of all
// nine quantities over the whole (just stretched) trajectory are compared with 1.001 x their limits; only the boolean is
// used unless another pass follows.  Half of all root finding of the pipeline goes into that check.  For a stretched
// segment the quantities that did not bind are now well below their limit, and a cheap certificate is enough:
//
//     max_{t in [0,T]} |p^(k)(t)|  <=  max_i |b_i|      (b = Bernstein coefficients of p^(k) on [0,T], convex hull property;
//                                                         one de Casteljau subdivision makes it tight to ~1 % on this data)
//
// If bound / limit <= 1.001 the exact maximum -- which is a value of the same polynomial at a point of [0,T] and therefore
// <= bound -- also passes, whatever zeros Jenkins-Traub would have found, so the (segment, quantity) needs no root finding
// for the check.  The bound is stored in place of the maximum and flagged; if the problem needs another pass after all
// (rare), flagged entries are recomputed exactly before they are read (ExtremaCompleteFn).  Decisions, pass counts and
// every number that leaves the routine are unchanged; measured on the round-1 workload 85 % of the check's root finding
// disappears (bench counters root_finds_reference_equivalent vs root_finds_executed, profiles/r01_final_bench_65536.json).
//
// Rounding: the power->Bernstein conversion and the Horner evaluation behind the exact maximum each err by at most a few
// n*eps*sum_j |c_j T^j|; the bound adds 1024 eps times that sum and a relative 1e-12, orders of magnitude above both.
#ifndef TG_BOUND_CUH_
#define TG_BOUND_CUH_

#include "tg_common.cuh"

namespace tg {

TG_HD constexpr double tg_binom(int n, int k) {
  double r = 1.0;
  for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
  return r;
}

// bound on max |p^(DERIV)| over [0, T] for one dimension; c: the 10 coefficients (increasing powers)
template <int DERIV>
TG_HD double bernstein_bound_1d(const double* __restrict__ c, double T) {
  constexpr int n = TG_N - 1 - DERIV;  // degree of the derivative
  double a[n + 1];
  double tp = 1.0, A = 0.0;
#pragma unroll
  for (int j = 0; j <= n; ++j) {
    a[j] = c[j + DERIV] * bcoef(DERIV, j + DERIV) * tp;
    A = A + dabs(a[j]);
    tp = tp * T;
  }
  double b[n + 1];
#pragma unroll
  for (int i = 0; i <= n; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j <= i; ++j) s = s + (tg_binom(i, j) / tg_binom(n, j)) * a[j];
    b[i] = s;
  }
  // one de Casteljau subdivision at 1/2: the control points of both halves are the first / last entries of every level
  double m = dmax(dabs(b[0]), dabs(b[n]));
#pragma unroll
  for (int level = 1; level <= n; ++level) {
#pragma unroll
    for (int i = 0; i <= n - level; ++i) b[i] = 0.5 * (b[i] + b[i + 1]);
    m = dmax(m, dmax(dabs(b[0]), dabs(b[n - level])));
  }
  return (m + 1024.0 * TG_DBL_EPSILON * A) * (1.0 + 1e-12);
}

// index into the nine limits (v_h v_v a_h a_v j_h j_v v_hdg a_hdg j_hdg) of quantity q (hor v,a,j ; ver v,a,j ; heading v,a,j)
TG_HD int limit_index(int q) {
  const int group = q / 3, d = q - 3 * group;
  return group == 2 ? 6 + d : 2 * d + group;
}

// nine bounds of one segment, in quantity order
TG_HD void segment_bounds(const double* __restrict__ coef, double T, double* __restrict__ out9) {
  const double* x = coef;
  const double* y = coef + TG_N;
  const double* z = coef + 2 * TG_N;
  const double* h = coef + 3 * TG_N;
  {
    const double bx = bernstein_bound_1d<1>(x, T), by = bernstein_bound_1d<1>(y, T);
    out9[0] = dsqrt(bx * bx + by * by) * (1.0 + 1e-12);
    out9[3] = bernstein_bound_1d<1>(z, T);
    out9[6] = bernstein_bound_1d<1>(h, T);
  }
  {
    const double bx = bernstein_bound_1d<2>(x, T), by = bernstein_bound_1d<2>(y, T);
    out9[1] = dsqrt(bx * bx + by * by) * (1.0 + 1e-12);
    out9[4] = bernstein_bound_1d<2>(z, T);
    out9[7] = bernstein_bound_1d<2>(h, T);
  }
  {
    const double bx = bernstein_bound_1d<3>(x, T), by = bernstein_bound_1d<3>(y, T);
    out9[2] = dsqrt(bx * bx + by * by) * (1.0 + 1e-12);
    out9[5] = bernstein_bound_1d<3>(z, T);
    out9[8] = bernstein_bound_1d<3>(h, T);
  }
}


// ---- adaptive certificate: is max_{t in [0,T]} || (p_d^(DERIV)(t))_{d < ND} || <= thr ? ------------------------------------
// Depth-first subdivision of the Bernstein control polygons.  An interval is discarded as soon as its hull bound is below
// thr; the search stops with `false` when a curve point (the polygon's end points ARE curve values) exceeds thr -- a certain
// violation -- or when the depth limit is reached.  `true` is a proof (up to the rounding margins, as above); `false` only
// means "run the exact method".  Near a maximum that touches thr from below the hull converges quadratically, so depth 8
// (intervals of T/256) certifies maxima within ~1e-4 of the threshold.
constexpr int kCertifyMaxDepth = 8;

template <int DERIV, int ND>
TG_HD bool certify_max_le(const double* __restrict__ c0, const double* __restrict__ c1, double T, double thr) {
  constexpr int n = TG_N - 1 - DERIV;
  if (!(thr >= 0.0)) return false;
  double cur[ND][n + 1];
  double err[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double* c = (d == 0) ? c0 : c1;
    double a[n + 1];
    double tp = 1.0, A = 0.0;
#pragma unroll
    for (int j = 0; j <= n; ++j) {
      a[j] = c[j + DERIV] * bcoef(DERIV, j + DERIV) * tp;
      A = A + dabs(a[j]);
      tp = tp * T;
    }
#pragma unroll
    for (int i = 0; i <= n; ++i) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j <= i; ++j) s = s + (tg_binom(i, j) / tg_binom(n, j)) * a[j];
      cur[d][i] = s;
    }
    err[d] = 1024.0 * TG_DBL_EPSILON * A;
    if (!dfinite(A)) return false;
  }
  double stack[kCertifyMaxDepth][ND][n + 1];
  int stack_depth[kCertifyMaxDepth];
  int sp = 0, depth = 0;
  for (;;) {
    // hull bound and end-point values of the current interval
    // The curve (p_d(t))_d lies in the convex hull of its control POINTS (all dimensions share the interval and the degree),
    // and the Euclidean norm is convex, so its maximum over the hull is attained at a control point: max_i ||b_i||.  (The
    // box bound sqrt(sum_d max_i b_di^2) loses a term of FIRST order in the angle by which the vector turns inside the
    // interval; with it, depth 8 could not certify a horizontal acceleration that touches its limit while turning.)
    double hull2 = 0.0, lo2a = 0.0, lo2b = 0.0;
#pragma unroll
    for (int i = 0; i <= n; ++i) {
      double s2 = 0.0;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double m = dabs(cur[d][i]) + err[d];
        s2 = s2 + m * m;
      }
      hull2 = dmax(hull2, s2);
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      lo2a = lo2a + cur[d][0] * cur[d][0];
      lo2b = lo2b + cur[d][n] * cur[d][n];
    }
    const double hull = dsqrt(hull2) * (1.0 + 1e-12);
    bool descend = false;
    if (!(hull <= thr)) {
      const double lo = dsqrt(dmax(lo2a, lo2b)) * (1.0 - 1e-9);
      if (lo > thr || depth >= kCertifyMaxDepth || !dfinite(hull)) return false;  // certain violation, or undecided
      descend = true;
    }
    if (descend) {
      // de Casteljau at 1/2: left half stays in `cur`, right half goes to the stack
      double right[ND][n + 1];
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        double w[n + 1];
#pragma unroll
        for (int i = 0; i <= n; ++i) w[i] = cur[d][i];
        right[d][n] = w[n];
#pragma unroll
        for (int level = 1; level <= n; ++level) {
#pragma unroll
          for (int i = 0; i <= n - level; ++i) w[i] = 0.5 * (w[i] + w[i + 1]);
          cur[d][level] = w[0];
          right[d][n - level] = w[n - level];
        }
      }
      depth = depth + 1;
#pragma unroll
      for (int d = 0; d < ND; ++d)
#pragma unroll
        for (int i = 0; i <= n; ++i) stack[sp][d][i] = right[d][i];
      stack_depth[sp] = depth;
      sp = sp + 1;
    } else {
      if (sp == 0) return true;
      sp = sp - 1;
#pragma unroll
      for (int d = 0; d < ND; ++d)
#pragma unroll
        for (int i = 0; i <= n; ++i) cur[d][i] = stack[sp][d][i];
      depth = stack_depth[sp];
    }
  }
}

// certificate for quantity q of one segment (coef: [4][10])
TG_HD bool certify_quantity_le(const double* __restrict__ coef, double T, int q, double thr) {
  const double* x = coef;
  const double* y = coef + TG_N;
  const double* z = coef + 2 * TG_N;
  const double* h = coef + 3 * TG_N;
  switch (q) {
    case 0: return certify_max_le<1, 2>(x, y, T, thr);
    case 1: return certify_max_le<2, 2>(x, y, T, thr);
    case 2: return certify_max_le<3, 2>(x, y, T, thr);
    case 3: return certify_max_le<1, 1>(z, z, T, thr);
    case 4: return certify_max_le<2, 1>(z, z, T, thr);
    case 5: return certify_max_le<3, 1>(z, z, T, thr);
    case 6: return certify_max_le<1, 1>(h, h, T, thr);
    case 7: return certify_max_le<2, 1>(h, h, T, thr);
    default: return certify_max_le<3, 1>(h, h, T, thr);
  }
}

}  // namespace tg

#endif  // TG_BOUND_CUH_
