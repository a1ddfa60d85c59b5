// oracle/ref_wrap_eth.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the REFERENCE's own eth_trajectory_generation classes.  oracle/Makefile compiles
//   /root/reference/src/eth_trajectory_generation/{motion_defines,polynomial,segment,timing,trajectory,trajectory_sampling,
//   vertex}.cpp, rpoly/rpoly_ak1.cpp and the header templates impl/polynomial_optimization_{linear,nonlinear}_impl.h
// UNMODIFIED, where they lie, against the stand-ins in oracle/ref_shim/ (Eigen, nlopt.hpp, mrs_lib cyclic, geometry_msgs,
// boost clamp -- none of which exists in this image), together with this file, into oracle/_ref/libref_eth.so.
// Every function below does nothing but marshal plain arrays into the reference's types, call the reference, and
// marshal back; the signatures mirror the orc_* functions of oracle/capi.cpp so that tests compare them one to one.
#include <cstdint>
#include <cstring>
#include <vector>

#include <eth_trajectory_generation/polynomial_optimization_nonlinear.h>
#include <eth_trajectory_generation/trajectory_sampling.h>

#include "oracle.h"

namespace etg = eth_trajectory_generation;
#ifndef REF_N
#define REF_N 10  // _ref/libref_eth_n{6,8,12}.so instantiate the reference templates for the other even N (lin.h:46-55)
#endif
static const int kN = REF_N, kHalf = REF_N / 2, kD = 4;

namespace orc {
// bridge used by ref_shim/nlopt.hpp: LD_LBFGS -> oracle/plis.cpp
int ref_shim_run_plis(int n, PlisObjective f, void* data, const double* lb, const double* ub, double* x, double* minf, int maxeval, double xtol_rel,
                      double ftol_rel, double xtol_abs) {
  PlisStop stop;
  stop.maxeval = maxeval;
  stop.xtol_rel = xtol_rel;
  stop.ftol_rel = ftol_rel;
  stop.xtol_abs = xtol_abs;
  return luksan_plis(n, f, data, lb, ub, x, minf, &stop);
}
}  // namespace orc

static etg::Vertex::Vector make_vertices(int V, const uint8_t* mask, const double* vals) {
  etg::Vertex::Vector vs;
  for (int v = 0; v < V; ++v) {
    etg::Vertex vx(kD);
    for (int k = 0; k < kHalf; ++k) {
      if (!((mask[v] >> k) & 1u)) continue;
      Eigen::VectorXd c(kD);
      for (int d = 0; d < kD; ++d) c[d] = vals[((size_t)v * kHalf + k) * kD + d];
      vx.addConstraint(k, c);
    }
    vs.push_back(vx);
  }
  return vs;
}

static etg::Segment::Vector make_segments(int S, const double* coeffs, const double* times) {
  etg::Segment::Vector segs;
  for (int i = 0; i < S; ++i) {
    etg::Segment s(kN, kD);
    s.setTime(times[i]);
    for (int d = 0; d < kD; ++d) {
      Eigen::VectorXd c(kN);
      for (int k = 0; k < kN; ++k) c[k] = coeffs[((size_t)i * kD + d) * kN + k];
      s[d] = etg::Polynomial(kN, c);
    }
    segs.push_back(s);
  }
  return segs;
}

static void store_segments(const etg::Segment::Vector& segs, double* coeffs, double* times) {
  for (size_t i = 0; i < segs.size(); ++i) {
    if (times) times[i] = segs[i].getTime();
    for (int d = 0; d < kD; ++d) {
      const Eigen::VectorXd c = segs[i][d].getCoefficients(0);
      for (int k = 0; k < kN; ++k) coeffs[(i * kD + d) * kN + k] = c[k];
    }
  }
}

extern "C" {

// PolynomialOptimization<REF_N>::setupMappingMatrix / invertMappingMatrix / computeQuadraticCostJacobian (row-major 10x10 out)
void ref_segment_matrices(double T, int r, double* A, double* Ainv, double* Q) {
  typedef etg::PolynomialOptimization<REF_N> PO;
  PO::SquareMatrix a, ai, q;
  a.setZero();
  ai.setZero();
  PO::setupMappingMatrix(T, &a);
  PO::invertMappingMatrix(a, &ai);
  PO::computeQuadraticCostJacobian(r, T, &q);
  for (int i = 0; i < kN; ++i)
    for (int j = 0; j < kN; ++j) {
      A[i * kN + j] = a(i, j);
      Ainv[i * kN + j] = ai(i, j);
      Q[i * kN + j] = q(i, j);
    }
}

// setupFromVertices + solveLinear + getSegments + computeCost + getFreeConstraints
int ref_solve_linear(int V, const uint8_t* mask, const double* vals, const double* times, int r, double* coeffs, double* cost, double* dp, int* dims) {
  etg::PolynomialOptimization<REF_N> opt(kD);
  const std::vector<double> t(times, times + (V - 1));
  if (!opt.setupFromVertices(make_vertices(V, mask, vals), t, r)) return 1;
  if (!opt.solveLinear()) return 2;
  etg::Segment::Vector segs;
  opt.getSegments(&segs);
  store_segments(segs, coeffs, nullptr);
  *cost = opt.computeCost();
  if (dims) {
    dims[0] = (int)opt.getNumberFixedConstraints();
    dims[1] = (int)opt.getNumberFreeConstraints();
  }
  if (dp) {
    std::vector<Eigen::VectorXd> fc;
    opt.getFreeConstraints(&fc);
    const size_t nf = opt.getNumberFreeConstraints();
    for (int d = 0; d < kD; ++d)
      for (size_t i = 0; i < nf; ++i) dp[d * nf + i] = fc[d][i];
  }
  return 0;
}

// getR: (n_fixed + n_free)^2 row-major
int ref_dense_R(int V, const uint8_t* mask, const double* vals, const double* times, int r, double* R) {
  etg::PolynomialOptimization<REF_N> opt(kD);
  const std::vector<double> t(times, times + (V - 1));
  if (!opt.setupFromVertices(make_vertices(V, mask, vals), t, r)) return 1;
  Eigen::MatrixXd Rm;
  opt.getR(&Rm);
  for (Eigen::Index i = 0; i < Rm.rows(); ++i)
    for (Eigen::Index j = 0; j < Rm.cols(); ++j) R[i * Rm.cols() + j] = Rm(i, j);
  return 0;
}

// computeMaximumOfMagnitude(derivative) after a linear solve at `times`
int ref_solve_max_magnitude(int V, const uint8_t* mask, const double* vals, const double* times, int r, int derivative, double* time, double* value,
                            int* segment_idx) {
  etg::PolynomialOptimization<REF_N> opt(kD);
  const std::vector<double> t(times, times + (V - 1));
  if (!opt.setupFromVertices(make_vertices(V, mask, vals), t, r)) return 1;
  opt.solveLinear();
  const etg::Extremum e = opt.computeMaximumOfMagnitude(derivative, nullptr);
  *time = e.time;
  *value = e.value;
  *segment_idx = e.segment_idx;
  return 0;
}

// estimateSegmentTimes (= Euclidean) and estimateSegmentTimesBaca from position-only vertices
void ref_estimate_times(int V, const double* pos4, const double* L, double* euclid, double* baca) {
  etg::Vertex::Vector vs;
  for (int v = 0; v < V; ++v) {
    etg::Vertex vx(kD);
    Eigen::VectorXd c(kD);
    for (int d = 0; d < kD; ++d) c[d] = pos4[4 * v + d];
    vx.addConstraint(etg::derivative_order::POSITION, c);
    vs.push_back(vx);
  }
  // limits: v_h, v_v, a_h, a_v, j_h, j_v, v_hdg, a_hdg, (j_hdg unused)
  const std::vector<double> e = etg::estimateSegmentTimes(vs, L[0], L[1], L[2], L[3], L[4], L[5], L[6], L[7]);
  const std::vector<double> b = etg::estimateSegmentTimesBaca(vs, L[0], L[1], L[2], L[3], L[4], L[5], L[6], L[7]);
  for (int i = 0; i < V - 1; ++i) {
    euclid[i] = e[i];
    baca[i] = b[i];
  }
}

// nine maxima per segment through Trajectory::computeMaxDerivatives{Horizontal,Vertical,Heading}(…, seg): hv ha hj vv va vj yv ya yj
void ref_segment_maxima(int S, const double* coeffs, const double* times, double* out) {
  etg::Trajectory traj;
  traj.setSegments(make_segments(S, coeffs, times));
  for (int i = 0; i < S; ++i) {
    traj.computeMaxDerivativesHorizontal(&out[i * 9 + 0], &out[i * 9 + 1], &out[i * 9 + 2], i);
    traj.computeMaxDerivativesVertical(&out[i * 9 + 3], &out[i * 9 + 4], &out[i * 9 + 5], i);
    traj.computeMaxDerivativesHeading(&out[i * 9 + 6], &out[i * 9 + 7], &out[i * 9 + 8], i);
  }
}

// Trajectory::scaleSegmentTimesToMeetConstraints in place (the pass count is not observable from outside: returns -1)
int ref_scale_times(int S, double* coeffs, double* times, const double* L, int* within) {
  etg::Trajectory traj;
  traj.setSegments(make_segments(S, coeffs, times));
  const bool w = traj.scaleSegmentTimesToMeetConstraints(L[0], L[1], L[2], L[3], L[4], L[5], L[6], L[7], L[8]);
  *within = w ? 1 : 0;
  etg::Segment::Vector segs;
  traj.getSegments(&segs);
  store_segments(segs, coeffs, times);
  return -1;
}

// sampleWholeTrajectory: rows of 19 doubles [p4 v4 a4 j3 s3 yaw_out]; heading entries of p, v, a come from the quaternion /
// angular-rate fields the way the node reads them back (getYaw(), getYawRate(), getYawAcc())
int ref_sample(int S, const double* coeffs, const double* times, double dt, int cap, double* out19, int64_t* t_ns) {
  etg::Trajectory traj;
  traj.setSegments(make_segments(S, coeffs, times));
  eth_mav_msgs::EigenTrajectoryPoint::Vector states;
  if (!etg::sampleWholeTrajectory(traj, dt, &states)) return 0;
  if ((int)states.size() > cap) return -(int)states.size();
  // the 4th components of position / velocity / acceleration are not stored in EigenTrajectoryPoint; recompute them the way
  // sampleTrajectoryInRange obtained them (evaluateRange at the same accumulated times)
  std::vector<Eigen::VectorXd> p, v, a;
  traj.evaluateRange(traj.getMinTime(), traj.getMaxTime(), dt, etg::derivative_order::POSITION, &p);
  traj.evaluateRange(traj.getMinTime(), traj.getMaxTime(), dt, etg::derivative_order::VELOCITY, &v);
  traj.evaluateRange(traj.getMinTime(), traj.getMaxTime(), dt, etg::derivative_order::ACCELERATION, &a);
  for (size_t i = 0; i < states.size(); ++i) {
    const eth_mav_msgs::EigenTrajectoryPoint& s = states[i];
    double* o = out19 + 19 * i;
    for (int d = 0; d < 3; ++d) {
      o[d] = s.position_W[d];
      o[4 + d] = s.velocity_W[d];
      o[8 + d] = s.acceleration_W[d];
      o[12 + d] = s.jerk_W[d];
      o[15 + d] = s.snap_W[d];
    }
    o[3] = p[i][3];
    o[7] = s.getYawRate();
    o[11] = s.getYawAcc();
    (void)v;
    (void)a;
    o[18] = s.getYaw();
    if (t_ns) t_ns[i] = s.time_from_start_ns;
  }
  return (int)states.size();
}

int ref_trajectory_evaluate(int S, const double* coeffs, const double* times, double t, int deriv, double* out4) {
  etg::Trajectory traj;
  traj.setSegments(make_segments(S, coeffs, times));
  double total = 0.0;
  for (int i = 0; i < S; ++i) total += times[i];
  const Eigen::VectorXd v = traj.evaluate(t, deriv);
  for (int d = 0; d < kD; ++d) out4[d] = v[d];
  // the reference signals "out of range" only by a log line and a zero vector (eth/trajectory.cpp:72-75)
  double acc = 0.0;
  for (int i = 0; i < S; ++i) acc += times[i];
  return (t > acc) ? 0 : 1;
}

// PolynomialOptimizationNonLinear<REF_N>: setupFromVertices + 12 x addMaximumMagnitudeConstraint + optimize() + getSegments,
// with the node's parameter recipe (node.cpp:880-905, 1063-1083); limits9 = v_h v_v a_h a_v j_h j_v v_hdg a_hdg j_hdg
int ref_time_alloc(int V, const uint8_t* mask, const double* vals, double* times, int r, int max_evals, double f_rel, double x_rel, const double* L,
                   double* coeffs, int* code, int* n_evals, double* final_cost) {
  etg::NonlinearOptimizationParameters parameters;
  parameters.f_rel = f_rel;
  parameters.x_rel = x_rel;
  parameters.time_alloc_method = etg::NonlinearOptimizationParameters::kMellingerOuterLoop;
  parameters.algorithm = nlopt::LD_LBFGS;
  parameters.initial_stepsize_rel = 0.1;
  parameters.max_iterations = max_evals;
  parameters.max_time = 1e9;
  etg::PolynomialOptimizationNonLinear<REF_N> opt(kD, parameters);
  const std::vector<double> t(times, times + (V - 1));
  opt.setupFromVertices(make_vertices(V, mask, vals), t, r);
  using namespace etg::derivative_order;
  opt.addMaximumMagnitudeConstraint(0, VELOCITY, L[0]);
  opt.addMaximumMagnitudeConstraint(0, ACCELERATION, L[2]);
  opt.addMaximumMagnitudeConstraint(0, JERK, L[4]);
  opt.addMaximumMagnitudeConstraint(1, VELOCITY, L[0]);
  opt.addMaximumMagnitudeConstraint(1, ACCELERATION, L[2]);
  opt.addMaximumMagnitudeConstraint(1, JERK, L[4]);
  opt.addMaximumMagnitudeConstraint(2, VELOCITY, L[1]);
  opt.addMaximumMagnitudeConstraint(2, ACCELERATION, L[3]);
  opt.addMaximumMagnitudeConstraint(2, JERK, L[5]);
  opt.addMaximumMagnitudeConstraint(3, VELOCITY, L[6]);
  opt.addMaximumMagnitudeConstraint(3, ACCELERATION, L[7]);
  opt.addMaximumMagnitudeConstraint(3, JERK, L[8]);
  opt.optimize();
  const etg::OptimizationInfo info = opt.getOptimizationInfo();
  *code = info.stopping_reason;
  *n_evals = info.n_iterations;
  *final_cost = info.cost_trajectory;
  etg::Segment::Vector segs;
  opt.getPolynomialOptimizationRef().getSegments(&segs);
  store_segments(segs, coeffs, times);
  return 0;
}

// objectiveFunctionTime (methods 0, 1) / objectiveFunctionTimeAndConstraints (3, 4) at K candidate vectors, through
// optimize() with the evaluation harness of ref_shim/nlopt.hpp.  total[k]; parts[k] = {trajectory, time, soft constraints}.
// The harness evaluates the start point first (x[0]) and then every candidate; constraints are (dimension 0, derivative, value).
int ref_objective(int V, const uint8_t* mask, const double* vals, int r, int method, int K, const double* x, int nvar, double time_penalty, int use_soft,
                  double soft_weight, int ncon, const int* con_deriv, const double* con_value, double* total, double* parts) {
  const int S = V - 1;
  for (int k = 0; k < K; ++k) {
    const double* xk = x + (size_t)k * nvar;
    etg::NonlinearOptimizationParameters parameters;
    parameters.time_alloc_method = static_cast<etg::NonlinearOptimizationParameters::TimeAllocMethod>(method);
    parameters.algorithm = nlopt::LN_BOBYQA;
    parameters.time_penalty = time_penalty;
    parameters.use_soft_constraints = use_soft != 0;
    parameters.soft_constraint_weight = soft_weight;
    parameters.random_seed = 0;
    etg::PolynomialOptimizationNonLinear<REF_N> opt(kD, parameters);
    const std::vector<double> t(xk, xk + S);
    opt.setupFromVertices(make_vertices(V, mask, vals), t, r);
    for (int c = 0; c < ncon; ++c) opt.addMaximumMagnitudeConstraint(0, con_deriv[c], con_value[c]);
    nlopt::ShimHarness::candidates().clear();
    if (method >= 3) nlopt::ShimHarness::candidates().push_back(std::vector<double>(xk, xk + nvar));
    opt.optimize();
    const etg::OptimizationInfo info = opt.getOptimizationInfo();
    // methods 0/1: the start point IS the candidate (first harness evaluation); 3/4: the start point is the linear solution
    // at the candidate's times, the candidate itself is evaluated second -- either way `info` holds the last evaluation
    if (parts) {
      parts[(size_t)k * 3 + 0] = info.cost_trajectory;
      parts[(size_t)k * 3 + 1] = info.cost_time;
      parts[(size_t)k * 3 + 2] = info.cost_soft_constraints;
    }
    total[k] = nlopt::ShimHarness::values().back();
  }
  return 0;
}

}  // extern "C"
