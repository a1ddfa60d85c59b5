// oracle/nonlinear.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates impl/polynomial_optimization_nonlinear_impl.h for time_alloc_method == kMellingerOuterLoop
// (the node's production setting, config/private/trajectory_generation.yaml:7).
//
// The objective and its forward-difference gradient follow nl_impl.h:256-333,616-649 line by line.  The optimiser the
// reference calls, nlopt::LD_LBFGS (nl_impl.h:68-74,178-191), is Luksan's PLIS; NLopt is a third-party dependency that is
// neither vendored under /root/reference nor installed here, so oracle/plis.cpp restates the published algorithm.
// PARITY UNPINNED for the post-optimisation segment times versus a real NLopt build; pinned GPU-vs-oracle.
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

namespace {
struct ObjCtx {
  LinearSolver* ls;
  NlInfo* info;
  std::vector<double> x, g;
  double last_f = 0.0;  // OptimizationInfo::cost_trajectory: written by every objective call (nl_impl.h:646)
};
// the C callback nlopt::opt hands to the algorithm (nlopt.hpp myvfunc): vectors in, vectors out
double plis_objective(int n, const double* x, double* grad, void* data) {
  ObjCtx* c = static_cast<ObjCtx*>(data);
  c->x.assign(x, x + n);
  const double f = objective(*c->ls, c->x, &c->g, c->info);
  c->last_f = f;
  for (int i = 0; i < n; ++i) grad[i] = c->g[i];
  return f;
}
}  // namespace

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

// nl_impl.h:159-234.  The LinearSolver is left at the LAST EVALUATED point, which is what the reference reads back after
// nlopt returns (nl_impl.h:210-215 use poly_opt_, not nlopt's x); nlopt's C++ wrapper throws on negative codes, the
// reference catches that, keeps result = FAILURE (-1) and goes on to the scaling because final_cost has been written.
int optimize_time_mellinger(LinearSolver& ls, const NlParams& P, const Limits& L, NlInfo* info) {
  std::vector<double> x = ls.times;
  const int S = (int)x.size();
  // nlopt_optimize_(): the start point must lie inside the bounds, else NLOPT_INVALID_ARGS (-2) before any evaluation;
  // the wrapper throws, final_cost is still DBL_MAX and the reference returns FAILURE without scaling (nl_impl.h:192-194).
  std::vector<double> lb(S, kTimeLowerBound), ub(S, DBL_MAX);
  for (int i = 0; i < S; ++i)
    if (x[i] < lb[i] || x[i] > ub[i]) {
      // (the reference would hand back an optimiser that never solved; one solve at the given times keeps the outputs defined)
      ls.solve();
      info->n_solves++;
      info->code = -1;
      info->final_cost = 0.0;  // OptimizationInfo::cost_trajectory keeps its initial value
      return -1;
    }
  ObjCtx ctx{&ls, info, {}, {}, 0.0};
  PlisStop stop;
  stop.maxeval = P.max_evals;
  stop.xtol_rel = P.x_rel;
  stop.ftol_rel = P.f_rel;
  stop.xtol_abs = P.x_abs;
  double fbest = DBL_MAX;
  int code = luksan_plis(S, plis_objective, &ctx, lb.data(), ub.data(), x.data(), &fbest, &stop);
  if (code < 0) code = -1;  // every exception lands in the same catch block (nl_impl.h:190-208)
  // what a caller of the reference can observe is getOptimizationInfo().cost_trajectory = the cost at the LAST EVALUATED point
  // (nlopt's own optimum `final_cost` is a local of optimizeTimeMellingerOuterLoop, printed in debug mode only)
  info->final_cost = ctx.last_f;
  scale_with_violation(ls, L, info);
  info->code = code;
  return code;
}

}  // namespace orc
