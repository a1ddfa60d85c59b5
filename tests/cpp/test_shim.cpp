// tests/cpp/test_shim.cpp -- drives include/eth_trajectory_generation_b200.hpp the way MrsTrajectoryGeneration::findTrajectory
// drives the reference classes (src/mrs_trajectory_generation.cpp:923-1169) and dumps every result as hex floats; the
// pytest wrapper (tests/test_cpp_shim.py) compares the dump bit for bit with the oracle.
// Usage: test_shim <out.txt>   (links against libtg_b200.so on the GPU box, or the host emulation in CPU-only runs)
#include <cstdio>
#include <cstdlib>

#include "../../include/eth_trajectory_generation_b200.hpp"

using namespace eth_trajectory_generation;

static void dump(FILE* f, const char* key, const double* v, size_t n) {
  std::fprintf(f, "%s %zu", key, n);
  for (size_t i = 0; i < n; ++i) std::fprintf(f, " %a", v[i]);
  std::fprintf(f, "\n");
}
static void dump_traj(FILE* f, const char* key, const Trajectory& t) {
  std::vector<double> coef, times;
  t.pack(&coef, &times);
  dump(f, (std::string(key) + "_times").c_str(), times.data(), times.size());
  dump(f, (std::string(key) + "_coef").c_str(), coef.data(), coef.size());
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "w");
  if (!f) return 2;
  // the 10-waypoint zig-zag of tmux/after_takeoff/plan_path.py:40-50 behind a hovering start (fixture F1b)
  std::vector<Vector> wps;
  wps.push_back(Vector{0, 0, 3, 0});
  for (int i = 0; i < 10; ++i) wps.push_back(Vector{2.0 * i, (i % 2 == 0) ? 0.5 : -0.5, 5, 0});
  const int derivative_to_optimize = derivative_order::ACCELERATION;
  Vertex::Vector vertices;
  for (size_t i = 0; i < wps.size(); ++i) {
    Vertex v(4);
    if (i == 0 || i + 1 == wps.size()) v.makeStartOrEnd(wps[i], derivative_to_optimize);  // node.cpp:938-945, 959-962
    else v.addConstraint(derivative_order::POSITION, wps[i]);                              // node.cpp:966-975
    vertices.push_back(v);
  }
  std::vector<double> times;
  for (size_t i = 0; i + 1 < wps.size(); ++i) times.push_back(1.0 + 0.25 * (double)(i % 3));

  // --- PolynomialOptimization<10>: setupFromVertices / solveLinear / computeCost / getTrajectory / Trajectory::evaluate
  PolynomialOptimization<10> lin(4);
  if (!lin.setupFromVertices(vertices, times, derivative_to_optimize)) return 3;
  if (!lin.solveLinear()) return 3;
  Trajectory tl;
  lin.getTrajectory(&tl);
  dump_traj(f, "lin", tl);
  const double cost = lin.computeCost();
  dump(f, "lin_cost", &cost, 1);
  const double tq[3] = {0.0, 2.6, tl.getMaxTime() * 0.75};
  for (int k = 0; k < 3; ++k) {
    const Vector p = tl.evaluate(tq[k], derivative_order::POSITION), v = tl.evaluate(tq[k], derivative_order::VELOCITY);
    dump(f, "lin_eval_p", p.data(), 4);
    dump(f, "lin_eval_v", v.data(), 4);
  }
  bool bad = lin.setupFromVertices(vertices, std::vector<double>(3, 1.0), derivative_to_optimize);  // wrong size: prints, returns false
  const double badv = bad ? 1.0 : 0.0;
  dump(f, "lin_bad_setup", &badv, 1);

  // --- PolynomialOptimizationNonLinear<10> as findTrajectory uses it (node.cpp:1063-1169)
  NonlinearOptimizationParameters parameters;
  parameters.f_rel = 0.05;
  parameters.x_rel = 0.1;
  parameters.max_iterations = 10;
  parameters.time_alloc_method = NonlinearOptimizationParameters::kMellingerOuterLoop;
  PolynomialOptimizationNonLinear<10> opt(4, parameters);
  lin.setupFromVertices(vertices, times, derivative_to_optimize);
  opt.setupFromVertices(vertices, times, derivative_to_optimize);
  const double vh = 4.0, vv = 2.0, ah = 2.0, av = 1.0, jh = 20.0, jv = 20.0, vy = 1.0, ay = 2.0, jy = 10.0;
  opt.addMaximumMagnitudeConstraint(0, derivative_order::VELOCITY, vh);
  opt.addMaximumMagnitudeConstraint(0, derivative_order::ACCELERATION, ah);
  opt.addMaximumMagnitudeConstraint(0, derivative_order::JERK, jh);
  opt.addMaximumMagnitudeConstraint(1, derivative_order::VELOCITY, vh);
  opt.addMaximumMagnitudeConstraint(1, derivative_order::ACCELERATION, ah);
  opt.addMaximumMagnitudeConstraint(1, derivative_order::JERK, jh);
  opt.addMaximumMagnitudeConstraint(2, derivative_order::VELOCITY, vv);
  opt.addMaximumMagnitudeConstraint(2, derivative_order::ACCELERATION, av);
  opt.addMaximumMagnitudeConstraint(2, derivative_order::JERK, jv);
  opt.addMaximumMagnitudeConstraint(3, derivative_order::VELOCITY, vy);
  opt.addMaximumMagnitudeConstraint(3, derivative_order::ACCELERATION, ay);
  opt.addMaximumMagnitudeConstraint(3, derivative_order::JERK, jy);
  const int code = opt.optimize();
  const OptimizationInfo info = opt.getOptimizationInfo();
  const double meta[4] = {(double)code, (double)info.n_iterations, (double)info.n_scale_passes, info.cost_trajectory};
  dump(f, "nl_meta", meta, 4);
  Trajectory tn;
  opt.getTrajectory(&tn);
  dump_traj(f, "nl", tn);
  TrajectoryPoint::Vector states;
  if (!sampleWholeTrajectory(tn, 0.2, &states)) return 4;
  std::vector<double> flat;
  for (const TrajectoryPoint& s : states) {
    flat.push_back(s.position_W[0]); flat.push_back(s.position_W[1]); flat.push_back(s.position_W[2]); flat.push_back(s.yaw);
    flat.push_back(s.velocity_W[0]); flat.push_back(s.acceleration_W[2]); flat.push_back((double)s.time_from_start_ns);
  }
  dump(f, "nl_samples", flat.data(), flat.size());
  double vmax, amax, jmax;
  tn.computeMaxDerivativesHorizontal(&vmax, &amax, &jmax);
  const double mh[3] = {vmax, amax, jmax};
  dump(f, "nl_max_h", mh, 3);
  // computeMaximumOfMagnitude (lin_impl.h:477-508) through the optimiser's linear part, as evaluateMaximumMagnitudeConstraint calls it
  for (int k = 1; k <= 3; ++k) {
    const Extremum e = opt.getPolynomialOptimizationRef().computeMaximumOfMagnitude(k, nullptr);
    const double me[3] = {e.time, e.value, (double)e.segment_idx};
    dump(f, "nl_maxmag", me, 3);
  }

  // --- objective of the derivative-free methods at three candidate time vectors (kSquaredTime, soft constraints on)
  {
    NonlinearOptimizationParameters po;
    po.time_alloc_method = NonlinearOptimizationParameters::kSquaredTime;
    PolynomialOptimizationNonLinear<10> oo(4, po);
    oo.setupFromVertices(vertices, times, derivative_to_optimize);
    oo.addMaximumMagnitudeConstraint(0, derivative_order::VELOCITY, 4.0);
    oo.addMaximumMagnitudeConstraint(0, derivative_order::ACCELERATION, 2.0);
    std::vector<std::vector<double>> xs(3, times);
    for (size_t i = 0; i < times.size(); ++i) { xs[1][i] *= 1.25; xs[2][i] *= 0.8; }
    std::vector<double> total, parts;
    if (!oo.evaluateObjectives(xs, &total, &parts)) return 6;
    dump(f, "obj_total", total.data(), total.size());
    dump(f, "obj_parts", parts.data(), parts.size());
  }

  // --- batch entry: two paths through optimize() (findTrajectory + validation + subdivision)
  TrajectoryGeneratorBatch gen;
  std::vector<std::vector<Waypoint>> paths(2);
  for (const Vector& w : wps) paths[0].push_back(Waypoint{w[0], w[1], w[2], w[3], false});
  const double p1[5][4] = {{10, 20, 3.5, 1.2}, {-5, -5, 5, 1}, {-5, 5, 5, 2}, {5, -5, 5, 3}, {5, 5, 5, 4}};  // test/get_path_before_takeoff/test.cpp:29-32
  for (int i = 0; i < 5; ++i) paths[1].push_back(Waypoint{p1[i][0], p1[i][1], p1[i][2], p1[i][3], false});
  std::vector<PathResult> res;
  if (!gen.optimize(paths, std::vector<InitialState>(), &res)) return 5;
  for (int p = 0; p < 2; ++p) {
    const double m[6] = {(double)res[p].info.success, (double)res[p].info.rounds, (double)res[p].info.n_waypoints, (double)res[p].info.n_samples,
                         (double)res[p].info.nlopt_code, (double)res[p].info.safe};
    dump(f, "batch_meta", m, 6);
    dump_traj(f, "batch", res[p].trajectory);
    dump(f, "batch_samples", res[p].samples_xyzh.data(), res[p].samples_xyzh.size());
  }
  // --- the steps either side of the path: preprocessPath / findTrajectoryFallback / getWaypointInTrajectoryIdxs
  std::vector<Waypoint> raw = paths[0];
  raw.insert(raw.begin() + 3, Waypoint{raw[2].x + 0.01, raw[2].y, raw[2].z, raw[2].heading, false});  // closer than min_waypoint_distance
  raw[6].stop_at = true;
  const std::vector<Waypoint> pre = gen.preprocessPath(raw);
  std::vector<double> flat_pre;
  for (const Waypoint& w : pre) { flat_pre.push_back(w.x); flat_pre.push_back(w.y); flat_pre.push_back(w.z); flat_pre.push_back(w.heading); flat_pre.push_back(w.stop_at ? 1.0 : 0.0); }
  dump(f, "pre", flat_pre.data(), flat_pre.size());
  const std::vector<double> fb = gen.findTrajectoryFallback(pre, 0.75, 0.75, 2.0);
  dump(f, "fallback", fb.data(), fb.size());
  const std::vector<int> ix = gen.getWaypointInTrajectoryIdxs(fb, pre);
  std::vector<double> ixd(ix.begin(), ix.end());
  dump(f, "idxs", ixd.data(), ixd.size());
  // --- path messages with override fields (node.cpp:1847-1902): three requests, two parameter groups + one refused override
  {
    DynamicsConstraints dc{4.0, 2.5, 2.0, 3.0, 2.0, 1.5, 30.0, 25.0, 20.0, 1.0, 2.0, 10.0};
    std::vector<PathRequest> reqs(3);
    for (int q = 0; q < 3; ++q)
      for (int i = 0; i < 5; ++i) reqs[q].points.push_back(PathRequest::Point{p1[i][0] + q, p1[i][1], p1[i][2], p1[i][3]});
    reqs[0].loop = true;
    reqs[0].relax_heading = true;
    reqs[0].max_deviation_from_path = 0.4;
    for (int q = 1; q < 3; ++q) {
      reqs[q].override_constraints = true;
      reqs[q].override_max_velocity_horizontal = 6.0; reqs[q].override_max_acceleration_horizontal = 3.5; reqs[q].override_max_jerk_horizontal = 35.0;
      reqs[q].override_max_velocity_vertical = 3.0;   reqs[q].override_max_acceleration_vertical = 2.5;   reqs[q].override_max_jerk_vertical = 99.0;
    }
    reqs[1].stop_at_waypoints = true;
    std::vector<InitialState> st(3);
    for (int q = 0; q < 3; ++q) {
      st[q] = InitialState{p1[0][3], {0.4, -0.3, 0.1, 0.0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    }
    st[2].velocity[0] = 7.0;  // faster than the override allows: refused, the tracker constraints stay
    std::vector<PathResult> rr;
    std::vector<ResolvedRequest> rs;
    if (!gen.optimizeRequests(reqs, dc, st, &rr, &rs)) return 6;
    for (int q = 0; q < 3; ++q) {
      std::vector<double> m(rs[q].params.limits, rs[q].params.limits + 9);
      m.push_back(rs[q].params.max_deviation); m.push_back(rs[q].prepend_state); m.push_back(rs[q].constraints_overridden); m.push_back((double)rs[q].waypoints.size());
      m.push_back(rs[q].waypoints.back().stop_at);
      m.push_back((double)rr[q].info.success); m.push_back((double)rr[q].info.rounds); m.push_back((double)rr[q].info.n_waypoints); m.push_back((double)rr[q].info.n_samples);
      dump(f, "req_meta", m.data(), m.size());
      dump(f, "req_samples", rr[q].samples_xyzh.data(), rr[q].samples_xyzh.size());
    }
  }
  std::fclose(f);
  std::printf("shim test wrote %s (%s)\n", argv[1], tg_version());
  return 0;
}
