// oracle/node.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates the numeric part of src/mrs_trajectory_generation.cpp: findTrajectory (857-1209),
// validateTrajectorySpatial (1401-1455), distFromSegment (1533-1554), interpolatePoint (1612-1625) and the
// validation / midpoint-subdivision loop of optimize() (729-785).  ROS plumbing, wall-clock budgets
// (overtime()) and the fallback sampler are outside the hot path.
#include <cmath>

#include "oracle.h"

namespace orc {

// node.cpp:857-1209
FindResult find_trajectory(const std::vector<Waypoint>& wp, const InitialState& init, const NodeParams& P) {
  FindResult R;
  const int r = P.derivative_to_optimize;
  // --- vertices (node.cpp:923-977)
  std::vector<Vertex> vertices;
  double last_heading = init.present ? init.heading : wp.at(0).c[3];
  for (size_t i = 0; i < wp.size(); ++i) {
    const double heading = srad_unwrap(wp[i].c[3], last_heading);
    last_heading = heading;
    const double pos[4] = {wp[i].c[0], wp[i].c[1], wp[i].c[2], heading};
    Vertex v;
    if (i == 0) {
      v.make_start_or_end(pos, r);
      v.add(0, pos);
      if (init.present) {
        v.add(1, init.vel);
        v.add(2, init.acc);
        v.add(3, init.jerk);
      }
    } else if (i == wp.size() - 1) {
      v.make_start_or_end(pos, r);
      v.add(0, pos);
    } else {
      v.add(0, pos);
      if (wp[i].stop_at) {
        const double z[4] = {0, 0, 0, 0};
        v.add(1, z);
        v.add(2, z);
        v.add(3, z);
      }
    }
    vertices.push_back(v);
  }
  // --- initial segment times (node.cpp:1045-1056)
  std::vector<double> times = estimate_times_euclidean(vertices, P.lim);
  const std::vector<double> baca = estimate_times_baca(vertices, P.lim);
  double total_baca = 0;
  for (size_t i = 0; i < baca.size(); ++i) total_baca += baca[i];
  R.baca_total = total_baca;
  // --- optimiser (node.cpp:1063-1083)
  LinearSolver ls;
  ls.setup(vertices, times, r);
  if (P.run_time_alloc) {
    optimize_time_mellinger(ls, P.nl, P.lim, &R.nl);
    // node.cpp:1138-1149 : accept >= 1 except 6 (MAXTIME), accept -1, reject the rest
    const int code = R.nl.code;
    if (!((code >= 1 && code != 6) || code == -1)) {
      R.status = kFindNloptRejected;
      return R;
    }
  } else {
    ls.solve();
    R.nl.n_solves = 1;
    R.nl.code = 1;
    R.nl.final_cost = ls.cost();
  }
  R.seg = ls.seg;
  R.times = ls.times;
  // --- sampling (node.cpp:1162-1169)
  const bool ok = sample_whole(R.seg, P.dt, &R.samples);
  // --- length sanity check (node.cpp:1178-1199)
  const double len = (double)R.samples.size() * P.dt;
  if (len > 1.0 && len > (P.max_len_factor * total_baca)) {
    R.status = kFindTooLong;
    return R;
  } else if (len > 1.0 && len < (P.min_len_factor * total_baca)) {
    R.status = kFindTooShort;
    return R;
  }
  if (!ok) R.status = kFindSampleFail;
  return R;
}

static double norm3(const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// node.cpp:1533-1554
double dist_from_segment(const double* p, const double* s1, const double* s2) {
  const double sv[3] = {s2[0] - s1[0], s2[1] - s1[1], s2[2] - s1[2]};
  const double len = norm3(sv);
  double u[3] = {sv[0], sv[1], sv[2]};
  if (len > 0.0) {  // Eigen normalize(): no-op on a zero vector
    u[0] /= len;
    u[1] /= len;
    u[2] /= len;
  }
  const double w[3] = {p[0] - s1[0], p[1] - s1[1], p[2] - s1[2]};
  const double coord = u[0] * w[0] + u[1] * w[1] + u[2] * w[2];
  if (coord < 0) return norm3(w);
  if (coord > len) {
    const double q[3] = {p[0] - s2[0], p[1] - s2[1], p[2] - s2[2]};
    return norm3(q);
  }
  // projector = u u^T (3x3), projection = s1 + projector * w, each row accumulated over ascending column
  double proj[3];
  for (int i = 0; i < 3; ++i) {
    const double m0 = u[i] * u[0], m1 = u[i] * u[1], m2 = u[i] * u[2];
    proj[i] = s1[i] + ((m0 * w[0] + m1 * w[1]) + m2 * w[2]);
  }
  const double dd[3] = {p[0] - proj[0], p[1] - proj[1], p[2] - proj[2]};
  return norm3(dd);
}

// node.cpp:1401-1455
Validation validate_spatial(const std::vector<Sample>& traj, const std::vector<Waypoint>& wp, const NodeParams& P) {
  Validation V;
  V.seg_ok.assign(wp.size() - 1, 1);
  int widx = 0;
  if (traj.empty()) return V;  // the reference underflows size()-1 here (node.cpp:1418); never reached with M >= 1
  for (size_t i = 0; i + 1 < traj.size(); ++i) {
    const double* sample = traj[i].p;
    const double* next = traj[i + 1].p;
    const double* s0 = wp[widx].c;
    const double* s1 = wp[widx + 1].c;
    const double dist = dist_from_segment(sample, s0, s1);
    const double end_dist = dist_from_segment(s1, sample, next);
    if (widx > 0 || P.first_segment_checked || (int)wp.size() <= 2) {
      if (dist > V.max_dev) V.max_dev = dist;
      if (dist > P.max_deviation) {
        V.seg_ok[widx] = 0;
        V.safe = false;
      }
    }
    if (end_dist < 0.05 && widx < ((int)wp.size() - 2)) widx++;
  }
  return V;
}

// node.cpp:1612-1625
Waypoint interpolate_point(const Waypoint& a, const Waypoint& b, double coeff) {
  Waypoint o;
  const double diff[3] = {b.c[0] - a.c[0], b.c[1] - a.c[1], b.c[2] - a.c[2]};
  o.c[0] = a.c[0] + coeff * diff[0];
  o.c[1] = a.c[1] + coeff * diff[1];
  o.c[2] = a.c[2] + coeff * diff[2];
  o.c[3] = rad_interp(a.c[3], b.c[3], coeff);
  o.stop_at = false;
  return o;
}

// node.cpp:620-851, numeric part
OptimizeResult optimize_path(const std::vector<Waypoint>& wp_in, const InitialState& init, const NodeParams& P) {
  OptimizeResult O;
  O.wp = wp_in;
  if (O.wp.size() <= 1) return O;  // "the path is empty (after postprocessing)"
  O.find = find_trajectory(O.wp, init, P);
  auto tally = [&]() {
    O.total_solves += O.find.nl.n_solves;
    O.total_root_calls += O.find.nl.n_root_calls;
    O.total_evals += O.find.nl.n_evals;
  };
  tally();
  if (O.find.status != kFindOk) return O;
  for (int k = 0; k < P.max_deviation_iters; ++k) {
    const Validation V = validate_spatial(O.find.samples, O.wp, P);
    O.max_dev = V.max_dev;
    if (P.check_deviation && !V.safe) {
      std::vector<Waypoint> nw;
      nw.reserve(O.wp.size() * 2);
      for (size_t i = 0; i + 1 < O.wp.size(); ++i) {
        nw.push_back(O.wp[i]);
        if (!V.seg_ok[i]) {
          if (i > 0 || P.first_segment_checked || (int)O.wp.size() <= 2) nw.push_back(interpolate_point(O.wp[i], O.wp[i + 1], 0.5));
        }
      }
      nw.push_back(O.wp.back());
      O.wp = nw;
      O.find = find_trajectory(O.wp, init, P);
      O.rounds++;
      tally();
      if (O.find.status != kFindOk) return O;
    } else {
      O.safe = true;
      break;
    }
  }
  O.success = true;
  return O;
}

}  // namespace orc
