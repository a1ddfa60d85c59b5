// tg_lbfgs.cuh -- segment-time allocation driver: Mellinger outer loop with a deterministic projected L-BFGS
// ("TG-LBFGS").  Replaces PolynomialOptimizationNonLinear<10>::optimizeTimeMellingerOuterLoop and the
// NLopt LD_LBFGS call inside it (reference: nl_impl.h:159-234, 616-649; NLopt itself is a third-party library that
// the reference does not vendor -- its objective/gradient are reproduced exactly, its iterate sequence is replaced
// by the documented algorithm below, with NLopt's stopping semantics and result codes).
//
// The optimiser is a per-problem STATE MACHINE advanced once per objective evaluation: the expensive part of an
// evaluation (S+1 linear solves: the base point and the S perturbed points of nl_impl.h:282-323) runs as batched
// warp solves; this file only consumes the S+1 costs.  One thread per problem.
#ifndef TG_LBFGS_CUH_
#define TG_LBFGS_CUH_

#include "tg_common.cuh"

namespace tg {

constexpr int kLbfgsMem = 10;
constexpr double kTimeLowerBound = 0.01;  // nl.h:32

struct LbfgsScalars {
  double f, gd, step;
  double rho[kLbfgsMem + 1];
  int npairs, n_evals, iter, ls_it, stage, code, done;
  int n_grads;  // evaluations whose gradient (the S perturbed solves) was actually computed
};

struct LbfgsVectors {  // all of length S for this problem; hist_* hold kLbfgsMem + 1 pairs, stride `hstride`
  double *x, *g, *d, *xeval, *hist_s, *hist_y;
  size_t hstride;
};

// NLopt's relstop() (util/stop.c)
TG_HD bool relstop(double vold, double vnew, double reltol, double abstol) {
  if (tgdm::disinf(vold)) return false;
  return (dabs(vnew - vold) < abstol || dabs(vnew - vold) < reltol * (dabs(vnew) + dabs(vold)) * 0.5 || (reltol > 0 && vnew == vold));
}

// gradient of the Mellinger objective from the S+1 costs of one evaluation (nl_impl.h:319-322)
TG_HD double mellinger_grad(int S, const double* __restrict__ costs, int n) {
  if (S == 1) return 0.0;  // nl_impl.h:264-271
  return (costs[1 + n] - costs[0]) / 0.1;
}

// Computes the search direction from (x, g, history) and the first trial point into xeval.
// Returns false when the projected gradient vanishes (-> NLOPT_SUCCESS).
TG_HD bool lbfgs_new_direction(int S, LbfgsScalars& st, const LbfgsVectors& v) {
  const double lb = kTimeLowerBound;
  double alpha[kLbfgsMem];
  double* q = v.d;  // build q in place, negate at the end
  for (int i = 0; i < S; ++i) {
    const bool act = (v.x[i] <= lb && v.g[i] > 0.0);
    q[i] = act ? 0.0 : v.g[i];
  }
  for (int k = st.npairs - 1; k >= 0; --k) {
    const double* hs = v.hist_s + (size_t)k * v.hstride;
    const double* hy = v.hist_y + (size_t)k * v.hstride;
    double sq = 0.0;
    for (int i = 0; i < S; ++i) sq = sq + hs[i] * q[i];
    alpha[k] = st.rho[k] * sq;
    for (int i = 0; i < S; ++i) q[i] = q[i] - alpha[k] * hy[i];
  }
  if (st.npairs > 0) {
    const double* hs = v.hist_s + (size_t)(st.npairs - 1) * v.hstride;
    const double* hy = v.hist_y + (size_t)(st.npairs - 1) * v.hstride;
    double sy = 0.0, yy = 0.0;
    for (int i = 0; i < S; ++i) {
      sy = sy + hs[i] * hy[i];
      yy = yy + hy[i] * hy[i];
    }
    const double gamma = sy / yy;
    for (int i = 0; i < S; ++i) q[i] = gamma * q[i];
  }
  for (int k = 0; k < st.npairs; ++k) {
    const double* hs = v.hist_s + (size_t)k * v.hstride;
    const double* hy = v.hist_y + (size_t)k * v.hstride;
    double yq = 0.0;
    for (int i = 0; i < S; ++i) yq = yq + hy[i] * q[i];
    const double beta = st.rho[k] * yq;
    for (int i = 0; i < S; ++i) q[i] = q[i] + (alpha[k] - beta) * hs[i];
  }
  double gd = 0.0;
  for (int i = 0; i < S; ++i) {
    const bool act = (v.x[i] <= lb && v.g[i] > 0.0);
    v.d[i] = act ? 0.0 : -q[i];
    gd = gd + v.g[i] * v.d[i];
  }
  if (!(gd < 0.0)) {
    st.npairs = 0;
    gd = 0.0;
    for (int i = 0; i < S; ++i) {
      const bool act = (v.x[i] <= lb && v.g[i] > 0.0);
      v.d[i] = act ? 0.0 : -v.g[i];
      gd = gd + v.g[i] * v.d[i];
    }
    if (!(gd < 0.0)) return false;
  }
  st.gd = gd;
  const double frac = (st.npairs == 0) ? 0.2 : 0.5;
  double step = (st.npairs == 0) ? TG_DBL_MAX : 1.0;
  for (int i = 0; i < S; ++i)
    if (v.d[i] != 0.0) {
      const double cap = frac * v.x[i] / dabs(v.d[i]);
      if (cap < step) step = cap;
    }
  st.step = step;
  st.ls_it = 0;
  for (int i = 0; i < S; ++i) v.xeval[i] = dmax(lb, v.x[i] + step * v.d[i]);
  return true;
}

// Start: clamp the initial times into xeval (the first evaluation point).
TG_HD void lbfgs_begin(int S, LbfgsScalars& st, const LbfgsVectors& v, const double* __restrict__ times0) {
  for (int i = 0; i < S; ++i) v.xeval[i] = dmax(kTimeLowerBound, times0[i]);
  st.f = 0.0;
  st.gd = 0.0;
  st.step = 0.0;
  st.npairs = 0;
  st.n_evals = 0;
  st.n_grads = 0;
  st.iter = 0;
  st.ls_it = 0;
  st.stage = 0;
  st.code = -1;
  st.done = 0;
}

// Does the evaluation whose base cost `fe` has just arrived need its gradient?  The S perturbed solves of an evaluation
// (nl_impl.h:282-323) are consumed only by the first evaluation and by an ACCEPTED line-search trial after which the
// iteration continues: a rejected trial reads the base cost alone, and so does an accepted one that ends the run (ftol,
// xtol, maxeval).  On the round-1 workload two evaluations out of three need no gradient, so the caller solves the base
// point first, asks this function, and runs the perturbed solves only where the answer is yes.  Same tests, in the same
// order, as lbfgs_advance below; the state is not touched.
TG_HD bool lbfgs_needs_gradient(int S, const LbfgsScalars& st, const LbfgsVectors& v, double fe, int max_evals, double f_rel, double x_rel,
                                double f_abs, double x_abs) {
  const int n_evals = st.n_evals + 1;
  if (st.stage == 0) return n_evals < max_evals;
  if (!(dfinite(fe) && fe <= st.f + 1e-4 * st.step * st.gd)) return false;  // rejected trial
  if (relstop(st.f, fe, f_rel, f_abs)) return false;
  bool x_stop = true;
  for (int i = 0; i < S; ++i)
    if (!relstop(v.x[i], v.xeval[i], x_rel, x_abs)) x_stop = false;
  if (x_stop) return false;
  return n_evals < max_evals;
}

// Advance after one evaluation at xeval.  have_grad: all S+1 costs are in `costs` (base first); otherwise only costs[0] is
// valid, which lbfgs_needs_gradient has shown to be enough for this step.
TG_HD_NOINLINE void lbfgs_advance(int S, LbfgsScalars& st, const LbfgsVectors& v, const double* __restrict__ costs, int max_evals,
                         double f_rel, double x_rel, double f_abs, double x_abs, bool have_grad) {
  if (st.done) return;
  st.n_evals += 1;
  if (have_grad) st.n_grads += 1;
  const double fe = costs[0];
  if (st.stage == 0) {
    st.f = fe;
    for (int i = 0; i < S; ++i) {
      v.x[i] = v.xeval[i];
      if (have_grad) v.g[i] = mellinger_grad(S, costs, i);
    }
    if (st.n_evals >= max_evals) { st.code = 5; st.done = 1; return; }
    st.stage = 1;
    if (!lbfgs_new_direction(S, st, v)) { st.code = 1; st.done = 1; }
    return;
  }
  // a line-search trial came back
  const bool finite = dfinite(fe);
  if (finite && fe <= st.f + 1e-4 * st.step * st.gd) {
    // accepted: curvature pair, stopping tests, next direction
    // the candidate pair is written into slot `npairs` (the history holds kLbfgsMem + 1 slots)
    double sy = 0.0, ss = 0.0, yy = 0.0;
    const int slot = st.npairs;
    double* hs = v.hist_s + (size_t)slot * v.hstride;
    double* hy = v.hist_y + (size_t)slot * v.hstride;
    bool x_stop = true;
    for (int i = 0; i < S; ++i) {
      if (!relstop(v.x[i], v.xeval[i], x_rel, x_abs)) x_stop = false;
      if (have_grad) {
        const double gn = mellinger_grad(S, costs, i);
        const double sv = v.xeval[i] - v.x[i];
        const double yv = gn - v.g[i];
        hs[i] = sv;
        hy[i] = yv;
        sy = sy + sv * yv;
        ss = ss + sv * sv;
        yy = yy + yv * yv;
        v.g[i] = gn;
      }
      v.x[i] = v.xeval[i];
    }
    if (have_grad && sy > 1e-10 * dsqrt(ss) * dsqrt(yy)) {
      st.rho[slot] = 1.0 / sy;
      if (slot == kLbfgsMem) {  // memory full: drop the oldest pair
        for (int k = 1; k <= kLbfgsMem; ++k) {
          for (int i = 0; i < S; ++i) {
            v.hist_s[(size_t)(k - 1) * v.hstride + i] = v.hist_s[(size_t)k * v.hstride + i];
            v.hist_y[(size_t)(k - 1) * v.hstride + i] = v.hist_y[(size_t)k * v.hstride + i];
          }
          st.rho[k - 1] = st.rho[k];
        }
      } else {
        st.npairs = slot + 1;
      }
    }
    const bool f_stop = relstop(st.f, fe, f_rel, f_abs);
    st.f = fe;
    st.iter += 1;
    if (f_stop) { st.code = 3; st.done = 1; return; }
    if (x_stop) { st.code = 4; st.done = 1; return; }
    if (st.n_evals >= max_evals) { st.code = 5; st.done = 1; return; }
    if (!lbfgs_new_direction(S, st, v)) { st.code = 1; st.done = 1; }
    return;
  }
  if (st.n_evals >= max_evals) { st.code = 5; st.done = 1; return; }
  st.ls_it += 1;
  if (st.ls_it >= 30) { st.code = -1; st.done = 1; return; }
  double next = 0.5 * st.step;
  if (finite) {
    const double denom = 2.0 * (fe - st.f - st.gd * st.step);
    if (denom > 0.0) {
      const double cand = -(st.gd * st.step * st.step) / denom;
      next = dmin(0.5 * st.step, dmax(0.1 * st.step, cand));
    }
  }
  st.step = next;
  for (int i = 0; i < S; ++i) v.xeval[i] = dmax(kTimeLowerBound, v.x[i] + next * v.d[i]);
}

}  // namespace tg

#endif  // TG_LBFGS_CUH_
