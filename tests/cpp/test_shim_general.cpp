// tests/cpp/test_shim_general.cpp -- the drop-in header for the other shapes the reference's templates allow
// (lin.h:46-55: PolynomialOptimization<_N>, _N even; eth/polynomial.h:45-48: kMaxN = 12; Vertex(D), Trajectory::D()):
// PolynomialOptimization<8>(3), <12>(4), <6>(1) and <10>(3) -- setupFromVertices / solveLinear / computeCost / getTrajectory /
// Trajectory::evaluate / evaluateRange / sampleWholeTrajectory.  Dumps hex floats for tests/test_cpp_shim.py, which compares
// them bit for bit with the oracle compiled for that N.
#include <cstdio>
#include <cstdlib>

#include "../../include/eth_trajectory_generation_b200.hpp"

using namespace eth_trajectory_generation;

static void dump(FILE* f, const std::string& key, const double* v, size_t n) {
  std::fprintf(f, "%s %zu", key.c_str(), n);
  for (size_t i = 0; i < n; ++i) std::fprintf(f, " %a", v[i]);
  std::fprintf(f, "\n");
}

// waypoint i of the test path in `dims` dimensions
static Vector waypoint(int i, int dims) {
  const double full[4] = {1.7 * i, (i % 2 == 0) ? 0.6 : -0.4, 3.0 + 0.25 * i, 0.15 * i};
  Vector v = b200::make_vector((size_t)dims, 0.0);
  for (int d = 0; d < dims; ++d) v[d] = full[d];
  return v;
}

template <int N>
static int run(FILE* f, const std::string& key, int dims, int derivative_to_optimize) {
  const int n_wp = 6;
  Vertex::Vector vertices;
  for (int i = 0; i < n_wp; ++i) {
    Vertex v(dims);
    if (i == 0 || i == n_wp - 1) v.makeStartOrEnd(waypoint(i, dims), derivative_to_optimize);
    else v.addConstraint(derivative_order::POSITION, waypoint(i, dims));
    if (i == 2) v.addConstraint(derivative_order::VELOCITY, b200::make_vector((size_t)dims, 0.3));
    vertices.push_back(v);
  }
  std::vector<double> times;
  for (int i = 0; i + 1 < n_wp; ++i) times.push_back(0.8 + 0.3 * (double)(i % 3));
  PolynomialOptimization<N> opt(dims);
  if (!opt.setupFromVertices(vertices, times, derivative_to_optimize)) return 3;
  if (!opt.solveLinear()) return 4;
  Trajectory t;
  opt.getTrajectory(&t);
  if (t.N() != N || t.D() != dims || t.K() != n_wp - 1) return 5;
  std::vector<double> coef, tt;
  t.pack(&coef, &tt);
  dump(f, key + "_coef", coef.data(), coef.size());
  const double cost = opt.computeCost();
  dump(f, key + "_cost", &cost, 1);
  const double tq[3] = {0.0, 1.3, t.getMaxTime() * 0.8};
  for (int k = 0; k < 3; ++k)
    for (int deriv = 0; deriv <= 2; ++deriv) {
      const Vector p = t.evaluate(tq[k], deriv);
      if ((int)p.size() != dims) return 6;
      dump(f, key + "_eval", p.data(), (size_t)dims);
    }
  std::vector<Vector> range;
  t.evaluateRange(0.4, t.getMaxTime(), 0.35, derivative_order::VELOCITY, &range);
  std::vector<double> flat;
  for (const Vector& v : range)
    for (int d = 0; d < dims; ++d) flat.push_back(v[d]);
  dump(f, key + "_range", flat.data(), flat.size());
  TrajectoryPoint::Vector states;
  const bool sampled = sampleWholeTrajectory(t, 0.2, &states);
  if (sampled != (dims >= 3)) return 7;  // eth/trajectory_sampling.cpp:58-61
  std::vector<double> smp;
  for (const TrajectoryPoint& s : states)
    for (int k = 0; k < 3; ++k) {
      smp.push_back(s.position_W[k]);
      smp.push_back(s.velocity_W[k]);
      smp.push_back(s.snap_W[k]);
    }
  dump(f, key + "_samples", smp.data(), smp.size());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "w");
  if (!f) return 2;
  int rc = 0;
  if ((rc = run<8>(f, "n8d3", 3, derivative_order::JERK))) return 10 + rc;
  if ((rc = run<12>(f, "n12d4", 4, derivative_order::SNAP))) return 20 + rc;
  if ((rc = run<6>(f, "n6d1", 1, derivative_order::ACCELERATION))) return 30 + rc;
  if ((rc = run<10>(f, "n10d3", 3, derivative_order::ACCELERATION))) return 40 + rc;
  // N = 10 on three dimensions: maxima and time scaling run on the tuned kernels with a zero fourth dimension
  {
    Vertex::Vector vertices;
    for (int i = 0; i < 5; ++i) {
      Vertex v(3);
      if (i == 0 || i == 4) v.makeStartOrEnd(waypoint(i, 3), derivative_order::ACCELERATION);
      else v.addConstraint(derivative_order::POSITION, waypoint(i, 3));
      vertices.push_back(v);
    }
    PolynomialOptimization<10> opt(3);
    if (!opt.setupFromVertices(vertices, std::vector<double>(4, 0.7), derivative_order::ACCELERATION) || !opt.solveLinear()) return 50;
    Trajectory t;
    opt.getTrajectory(&t);
    double m[6];
    t.computeMaxDerivativesHorizontal(&m[0], &m[1], &m[2]);
    t.computeMaxDerivativesVertical(&m[3], &m[4], &m[5]);
    dump(f, "n10d3_max", m, 6);
    const Extremum e = opt.computeMaximumOfMagnitude(derivative_order::VELOCITY);
    const double ev[3] = {e.time, e.value, (double)e.segment_idx};
    dump(f, "n10d3_maxmag", ev, 3);
    const bool within = t.scaleSegmentTimesToMeetConstraints(4.0, 2.0, 2.0, 1.0, 20.0, 20.0, 1.0, 2.0, 10.0);
    std::vector<double> coef, tt;
    t.pack(&coef, &tt);
    if (t.D() != 3) return 51;
    tt.push_back(within ? 1.0 : 0.0);
    dump(f, "n10d3_scaled_times", tt.data(), tt.size());
    dump(f, "n10d3_scaled_coef", coef.data(), coef.size());
    // PolynomialOptimizationNonLinear<10> on three dimensions (Mellinger): optimize() + getTrajectory; a derivative-free method is refused
    {
      NonlinearOptimizationParameters prm;
      prm.time_alloc_method = NonlinearOptimizationParameters::kMellingerOuterLoop;
      PolynomialOptimizationNonLinear<10> nl(3, prm);
      if (!nl.setupFromVertices(vertices, std::vector<double>(4, 0.7), derivative_order::ACCELERATION)) return 53;
      nl.addMaximumMagnitudeConstraint(0, derivative_order::VELOCITY, 4.0);
      nl.addMaximumMagnitudeConstraint(2, derivative_order::VELOCITY, 2.0);
      nl.addMaximumMagnitudeConstraint(0, derivative_order::ACCELERATION, 2.0);
      nl.addMaximumMagnitudeConstraint(2, derivative_order::ACCELERATION, 1.0);
      const int code = nl.optimize();
      Trajectory tn;
      nl.getTrajectory(&tn);
      if (tn.D() != 3 || tn.N() != 10) return 54;
      std::vector<double> cn, ttn;
      tn.pack(&cn, &ttn);
      ttn.push_back((double)code);
      dump(f, "nl3_times_code", ttn.data(), ttn.size());
      dump(f, "nl3_coef", cn.data(), cn.size());
      prm.time_alloc_method = NonlinearOptimizationParameters::kSquaredTime;
      PolynomialOptimizationNonLinear<10> nl0(3, prm);
      const double refused0 = nl0.setupFromVertices(vertices, std::vector<double>(4, 0.7), derivative_order::ACCELERATION) ? 0.0 : 1.0;
      dump(f, "nl3_dfo_refused", &refused0, 1);
    }
    // a shape outside the template's range of the B200 path is refused
    PolynomialOptimization<8> o8(5);
    const double refused = o8.setupFromVertices(vertices, std::vector<double>(4, 0.7), 2) ? 0.0 : 1.0;
    dump(f, "d5_refused", &refused, 1);
    // N != 10: maxima are refused loudly, not fabricated
    PolynomialOptimization<8> p8(3);
    if (!p8.setupFromVertices(vertices, std::vector<double>(4, 0.7), 2) || !p8.solveLinear()) return 52;
    Trajectory t8;
    p8.getTrajectory(&t8);
    const double sc = t8.scaleSegmentTimesToMeetConstraints(4.0, 2.0, 2.0, 1.0, 20.0, 20.0, 1.0, 2.0, 10.0) ? 1.0 : 0.0;
    dump(f, "n8_scale_refused", &sc, 1);
  }
  std::fclose(f);
  return 0;
}
