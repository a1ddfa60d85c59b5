// oracle/nonlinear.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates impl/polynomial_optimization_nonlinear_impl.h for time_alloc_method == kMellingerOuterLoop
// (the node's production setting, config/private/trajectory_generation.yaml:7).
//
// NLopt is a third-party dependency that is absent here (package.xml:24 pins nlopt >= 2.4.2; LD_LBFGS is
// Luksan's PLIS).  Its objective and gradient are fully pinned by nl_impl.h:256-333,616-649 and are restated
// exactly; its ITERATE SEQUENCE is not reproducible without the source, so it is replaced by the
// deterministic projected L-BFGS written down in DESIGN.md ("TG-LBFGS").  Stopping tests follow NLopt's
// documented semantics (ftol_rel, xtol_rel, maxeval) and return NLopt's result codes.  PARITY UNPINNED for
// the post-optimisation segment times versus the real NLopt; pinned GPU-vs-oracle.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "oracle.h"

namespace orc {

constexpr double kTimeLowerBound = 0.01;  // nl.h:32

// nl_impl.h:256-333.  On entry the solver holds the solution at `times`; on exit too.
double mellinger_cost_and_grad(LinearSolver& ls, std::vector<double>* grad, int* n_solves) {
  const std::vector<double> times = ls.times;
  const double J_d = ls.cost();
  const int S = ls.S;
  if (S == 1) {
    if (grad) grad->assign(S, 0.0);
    return J_d;
  }
  if (grad) {
    grad->assign(S, 0.0);
    std::vector<double> bigger(S);
    const double increment_time = 0.1;
    for (int n = 0; n < S; ++n) {
      bigger = times;
      const double corr = increment_time / ((double)S - 1.0);
      for (int i = 0; i < S; ++i) {
        if (i == n) bigger[i] += increment_time;
        else bigger[i] -= corr;
      }
      for (double& t : bigger) t = std::max(kTimeLowerBound, t);
      ls.update_times(bigger);
      ls.solve();
      if (n_solves) ++*n_solves;
      const double J_bigger = ls.cost();
      (*grad)[n] = (J_bigger - J_d) / increment_time;
    }
    ls.update_times(times);
    ls.solve();
    if (n_solves) ++*n_solves;
  }
  return J_d;
}

// nl_impl.h:616-649
static double objective(LinearSolver& ls, const std::vector<double>& x, std::vector<double>* grad, NlInfo* info) {
  ls.update_times(x);
  ls.solve();
  info->n_solves++;
  const double c = mellinger_cost_and_grad(ls, grad, &info->n_solves);
  info->n_evals++;
  return c;
}

// NLopt's relstop() (util/stop.c) -- the documented ftol/xtol semantics.
static bool relstop(double vold, double vnew, double reltol, double abstol) {
  if (std::isinf(vold)) return false;
  return (std::fabs(vnew - vold) < abstol || std::fabs(vnew - vold) < reltol * (std::fabs(vnew) + std::fabs(vold)) * 0.5 ||
          (reltol > 0 && vnew == vold));
}

namespace {
constexpr int kMem = 10;  // history pairs; >= the 9 iterations that maxeval = 10 allows
struct Lbfgs {
  int S, npairs = 0;
  std::vector<double> s[kMem], y[kMem];
  double rho[kMem];
};
}  // namespace

// TG-LBFGS: deterministic projected L-BFGS with Armijo back-tracking (quadratic interpolation).
// Returns the NLopt-style code.  x is in/out; the LinearSolver is left at the LAST EVALUATED point, which is what
// the reference reads back after nlopt returns (nl_impl.h:210-215 use poly_opt_, not nlopt's x).
static int tg_lbfgs(LinearSolver& ls, std::vector<double>& x, const NlParams& P, NlInfo* info, double* fbest) {
  const int S = (int)x.size();
  const double lb = kTimeLowerBound;
  for (double& t : x) t = std::max(lb, t);
  std::vector<double> g(S), gn(S), d(S), xn(S), q(S);
  Lbfgs H;
  H.S = S;
  static const bool trace = std::getenv("ORC_TRACE") != nullptr;
  double f = objective(ls, x, &g, info);
  *fbest = f;
  if (info->n_evals >= P.max_evals) return 5;  // NLOPT_MAXEVAL_REACHED
  for (int iter = 0;; ++iter) {
    // active set: at the lower bound with the gradient pushing outwards
    std::vector<uint8_t> act(S);
    for (int i = 0; i < S; ++i) act[i] = (x[i] <= lb && g[i] > 0.0) ? 1 : 0;
    // two-loop recursion on the free variables
    double alpha[kMem];
    for (int i = 0; i < S; ++i) q[i] = act[i] ? 0.0 : g[i];
    for (int k = H.npairs - 1; k >= 0; --k) {
      double sq = 0.0;
      for (int i = 0; i < S; ++i) sq += H.s[k][i] * q[i];
      alpha[k] = H.rho[k] * sq;
      for (int i = 0; i < S; ++i) q[i] = q[i] - alpha[k] * H.y[k][i];
    }
    if (H.npairs > 0) {
      const int k = H.npairs - 1;
      double sy = 0.0, yy = 0.0;
      for (int i = 0; i < S; ++i) { sy += H.s[k][i] * H.y[k][i]; yy += H.y[k][i] * H.y[k][i]; }
      const double gamma = sy / yy;
      for (int i = 0; i < S; ++i) q[i] = gamma * q[i];
    }
    for (int k = 0; k < H.npairs; ++k) {
      double yq = 0.0;
      for (int i = 0; i < S; ++i) yq += H.y[k][i] * q[i];
      const double beta = H.rho[k] * yq;
      for (int i = 0; i < S; ++i) q[i] = q[i] + (alpha[k] - beta) * H.s[k][i];
    }
    double gd = 0.0;
    for (int i = 0; i < S; ++i) { d[i] = act[i] ? 0.0 : -q[i]; gd += g[i] * d[i]; }
    if (!(gd < 0.0)) {  // not a descent direction: drop the memory, steepest descent
      H.npairs = 0;
      gd = 0.0;
      for (int i = 0; i < S; ++i) { d[i] = act[i] ? 0.0 : -g[i]; gd += g[i] * d[i]; }
      if (!(gd < 0.0)) return 1;  // projected gradient is zero: NLOPT_SUCCESS
    }
    // first trial step: unit quasi-Newton step, capped so that no segment time changes by more than 50 %
    // (first iteration, no curvature information yet: 20 %)
    const double frac = (H.npairs == 0) ? 0.2 : 0.5;
    double step = (H.npairs == 0) ? DBL_MAX : 1.0;
    for (int i = 0; i < S; ++i)
      if (d[i] != 0.0) {
        const double cap = frac * x[i] / std::fabs(d[i]);
        if (cap < step) step = cap;
      }
    // Armijo back-tracking
    bool accepted = false;
    double fn = f;
    for (int ls_it = 0; ls_it < 30; ++ls_it) {
      for (int i = 0; i < S; ++i) xn[i] = std::max(lb, x[i] + step * d[i]);
      fn = objective(ls, xn, &gn, info);
      if (trace) std::fprintf(stderr, "[orc lbfgs] iter %d ls %d step %.4g f %.9g -> fn %.9g gd %.4g evals %d\n", iter, ls_it, step, f, fn, gd, info->n_evals);
      const bool finite = std::isfinite(fn);
      if (finite && fn <= f + 1e-4 * step * gd) {
        accepted = true;
        break;
      }
      if (info->n_evals >= P.max_evals) return 5;
      // quadratic interpolation through f, gd, fn ; clamped to [0.1, 0.5] * step
      double next = 0.5 * step;
      if (finite) {
        const double denom = 2.0 * (fn - f - gd * step);
        if (denom > 0.0) {
          const double cand = -(gd * step * step) / denom;
          next = std::min(0.5 * step, std::max(0.1 * step, cand));
        }
      }
      step = next;
    }
    if (!accepted) return -1;  // NLOPT_FAILURE (tolerated by the node, node.cpp:1141-1143)
    // curvature pair
    {
      std::vector<double> sv(S), yv(S);
      double sy = 0.0, ss = 0.0, yy = 0.0;
      for (int i = 0; i < S; ++i) {
        sv[i] = xn[i] - x[i];
        yv[i] = gn[i] - g[i];
        sy += sv[i] * yv[i];
        ss += sv[i] * sv[i];
        yy += yv[i] * yv[i];
      }
      if (sy > 1e-10 * std::sqrt(ss) * std::sqrt(yy)) {
        if (H.npairs == kMem) {
          for (int k = 1; k < kMem; ++k) { H.s[k - 1] = H.s[k]; H.y[k - 1] = H.y[k]; H.rho[k - 1] = H.rho[k]; }
          H.npairs--;
        }
        H.s[H.npairs] = sv;
        H.y[H.npairs] = yv;
        H.rho[H.npairs] = 1.0 / sy;
        H.npairs++;
      }
    }
    // stopping tests on the accepted step: ftol, then xtol, then maxeval
    const bool f_stop = relstop(f, fn, P.f_rel, P.f_abs);
    bool x_stop = true;
    for (int i = 0; i < S; ++i)
      if (!relstop(x[i], xn[i], P.x_rel, P.x_abs)) { x_stop = false; break; }
    x = xn;
    g = gn;
    f = fn;
    *fbest = f;
    if (f_stop) return 3;  // NLOPT_FTOL_REACHED
    if (x_stop) return 4;  // NLOPT_XTOL_REACHED
    if (info->n_evals >= P.max_evals) return 5;
  }
}

// nl_impl.h:335-427
static void scale_with_violation(LinearSolver& ls, const Limits& L, NlInfo* info) {
  std::vector<Segment> seg = ls.seg;  // poly_opt_.getTrajectory(&traj)
  bool within = false;
  info->n_scale_passes = scale_times_to_meet_constraints(seg, L, &within, &info->n_root_calls);
  std::vector<double> t(seg.size());
  for (size_t i = 0; i < seg.size(); ++i) t[i] = seg[i].T;
  ls.update_times(t);
  ls.solve();
  info->n_solves++;
}

// nl_impl.h:159-234
int optimize_time_mellinger(LinearSolver& ls, const NlParams& P, const Limits& L, NlInfo* info) {
  std::vector<double> x = ls.times;
  double fbest = DBL_MAX;
  const int code = tg_lbfgs(ls, x, P, info, &fbest);
  info->final_cost = fbest;
  scale_with_violation(ls, L, info);
  info->code = code;
  return code;
}

}  // namespace orc
