// oracle/ref_shim/nlopt.hpp -- TEST INFRASTRUCTURE ONLY.
// Stand-in for the slice of NLopt's C++ wrapper that the reference's polynomial_optimization_nonlinear(_impl).h uses
// (nlopt::opt with set_*tol*, set_maxeval, set_maxtime, set_lower/upper_bounds, set_initial_step, set_min_objective,
// add_inequality_constraint, optimize; nlopt_srand; the result / algorithm enums).  NLopt is absent from this image.
//   * LD_LBFGS runs oracle/plis.cpp (the restatement of Luksan's PLIS, see there) through the same vfunc trampoline
//     nlopt.hpp uses, with nlopt_optimize_'s bound check in front and nlopt::opt::optimize's exception mapping behind.
//   * every other algorithm (the reference's default LN_BOBYQA of the soft-constraint modes) is an EVALUATION HARNESS:
//     it calls the objective at the start point and at every candidate registered with ref_shim_set_candidates(), records
//     the values, and returns MAXEVAL_REACHED with the best candidate.  BOBYQA itself is not restated; the harness exists
//     so that the reference's own objective functions (nl_impl.h:567-722) can be compared with the restatement.
#ifndef ORACLE_REF_SHIM_NLOPT_HPP_
#define ORACLE_REF_SHIM_NLOPT_HPP_
#include <cfloat>
#include <cmath>
#include <new>
#include <stdexcept>
#include <vector>

namespace orc {  // oracle/oracle.h
typedef double (*PlisObjective)(int n, const double* x, double* grad, void* data);
struct PlisStop;
int ref_shim_run_plis(int n, PlisObjective f, void* data, const double* lb, const double* ub, double* x, double* minf, int maxeval, double xtol_rel,
                      double ftol_rel, double xtol_abs);
}  // namespace orc

inline void nlopt_srand(unsigned long) {}
inline void nlopt_srand_time() {}

namespace nlopt {
enum algorithm {
  GN_DIRECT = 0, GN_DIRECT_L, GN_DIRECT_L_RAND, GN_DIRECT_NOSCAL, GN_DIRECT_L_NOSCAL, GN_DIRECT_L_RAND_NOSCAL, GN_ORIG_DIRECT, GN_ORIG_DIRECT_L,
  GD_STOGO, GD_STOGO_RAND, LD_LBFGS_NOCEDAL, LD_LBFGS, LN_PRAXIS, LD_VAR1, LD_VAR2, LD_TNEWTON, LD_TNEWTON_RESTART, LD_TNEWTON_PRECOND,
  LD_TNEWTON_PRECOND_RESTART, GN_CRS2_LM, GN_MLSL, GD_MLSL, GN_MLSL_LDS, GD_MLSL_LDS, LD_MMA, LN_COBYLA, LN_NEWUOA, LN_NEWUOA_BOUND, LN_NELDERMEAD,
  LN_SBPLX, LN_AUGLAG, LD_AUGLAG, LN_AUGLAG_EQ, LD_AUGLAG_EQ, LN_BOBYQA, GN_ISRES, AUGLAG, AUGLAG_EQ, G_MLSL, G_MLSL_LDS, LD_SLSQP, LD_CCSAQ, GN_ESCH,
  NUM_ALGORITHMS
};
enum result {
  FAILURE = -1, INVALID_ARGS = -2, OUT_OF_MEMORY = -3, ROUNDOFF_LIMITED = -4, FORCED_STOP = -5,
  SUCCESS = 1, STOPVAL_REACHED = 2, FTOL_REACHED = 3, XTOL_REACHED = 4, MAXEVAL_REACHED = 5, MAXTIME_REACHED = 6
};
typedef double (*vfunc)(const std::vector<double>& x, std::vector<double>& grad, void* data);

struct ShimHarness {  // candidates for the evaluation harness (tests only)
  static std::vector<std::vector<double>>& candidates() {
    static std::vector<std::vector<double>> c;
    return c;
  }
  static std::vector<double>& values() {
    static std::vector<double> v;
    return v;
  }
};

class opt {
 public:
  opt() : alg_(LN_BOBYQA), n_(0) { init(); }
  opt(algorithm a, unsigned n) : alg_(a), n_(n) { init(); }
  void set_ftol_rel(double v) { ftol_rel_ = v; }
  void set_ftol_abs(double v) { ftol_abs_ = v; }
  void set_xtol_rel(double v) { xtol_rel_ = v; }
  void set_xtol_abs(double v) { xtol_abs_ = v; }
  void set_maxeval(int v) { maxeval_ = v; }
  void set_maxtime(double v) { maxtime_ = v; }
  void set_lower_bounds(double v) { lb_.assign(n_, v); }
  void set_upper_bounds(double v) { ub_.assign(n_, v); }
  void set_lower_bounds(const std::vector<double>& v) { check(v); lb_ = v; }
  void set_upper_bounds(const std::vector<double>& v) { check(v); ub_ = v; }
  void set_initial_step(const std::vector<double>& v) {
    check(v);
    for (double s : v)
      if (s == 0.0) throw std::invalid_argument("nlopt invalid argument");
    dx_ = v;
  }
  void set_min_objective(vfunc f, void* data) { f_ = f; f_data_ = data; }
  void add_inequality_constraint(vfunc f, void* data, double tol) { ineq_.push_back(Con{f, data, tol}); }
  unsigned get_dimension() const { return n_; }
  int get_numevals() const { return nevals_; }

  result optimize(std::vector<double>& x, double& opt_f) {
    if (x.size() != n_) throw std::invalid_argument("dimension mismatch");
    const result ret = run(x, opt_f);
    switch (ret) {  // nlopt.hpp mythrow()
      case FAILURE: throw std::runtime_error("nlopt failure");
      case OUT_OF_MEMORY: throw std::bad_alloc();
      case INVALID_ARGS: throw std::invalid_argument("nlopt invalid argument");
      case ROUNDOFF_LIMITED: throw std::runtime_error("nlopt roundoff-limited");
      case FORCED_STOP: throw std::runtime_error("nlopt forced stop");
      default: break;
    }
    return ret;
  }

 private:
  struct Con { vfunc f; void* data; double tol; };
  struct Tramp { opt* self; std::vector<double> xv, gv; };
  void init() {
    ftol_rel_ = ftol_abs_ = xtol_rel_ = xtol_abs_ = 0.0;
    maxeval_ = 0;
    maxtime_ = 0.0;
    f_ = nullptr;
    f_data_ = nullptr;
    nevals_ = 0;
    lb_.assign(n_, -HUGE_VAL);
    ub_.assign(n_, HUGE_VAL);
  }
  void check(const std::vector<double>& v) const {
    if (v.size() != n_) throw std::invalid_argument("dimension mismatch");
  }
  static double tramp(int n, const double* x, double* grad, void* data) {  // nlopt.hpp myvfunc
    Tramp* t = static_cast<Tramp*>(data);
    t->xv.assign(x, x + n);
    t->gv.assign(grad ? n : 0, 0.0);
    const double f = t->self->f_(t->xv, t->gv, t->self->f_data_);
    if (grad)
      for (int i = 0; i < n; ++i) grad[i] = t->gv[i];
    ++t->self->nevals_;
    return f;
  }
  result run(std::vector<double>& x, double& opt_f) {
    if (!f_) return INVALID_ARGS;
    for (unsigned i = 0; i < n_; ++i)  // nlopt_optimize_: the start point must satisfy the bounds
      if (lb_[i] > ub_[i] || x[i] < lb_[i] || x[i] > ub_[i]) return INVALID_ARGS;
    Tramp t{this, {}, {}};
    if (alg_ == LD_LBFGS) {
      const int code = orc::ref_shim_run_plis((int)n_, &opt::tramp, &t, lb_.data(), ub_.data(), x.data(), &opt_f, maxeval_, xtol_rel_, ftol_rel_, xtol_abs_);
      return (result)code;
    }
    // evaluation harness for the derivative-free algorithms
    std::vector<double>& vals = ShimHarness::values();
    vals.clear();
    double best = tramp((int)n_, x.data(), nullptr, &t);
    vals.push_back(best);
    std::vector<double> bx = x;
    for (const std::vector<double>& c : ShimHarness::candidates()) {
      if (c.size() != n_) continue;
      const double v = tramp((int)n_, c.data(), nullptr, &t);
      vals.push_back(v);
      if (v < best) {
        best = v;
        bx = c;
      }
    }
    x = bx;
    opt_f = best;
    return MAXEVAL_REACHED;
  }
  algorithm alg_;
  unsigned n_;
  double ftol_rel_, ftol_abs_, xtol_rel_, xtol_abs_, maxtime_;
  int maxeval_, nevals_;
  std::vector<double> lb_, ub_, dx_;
  vfunc f_;
  void* f_data_;
  std::vector<Con> ineq_;
};
}  // namespace nlopt
#endif
