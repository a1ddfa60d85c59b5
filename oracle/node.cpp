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

// node.cpp:1560-1606
std::vector<std::array<double, 4>> trajectory_reference(const std::vector<Sample>& traj, bool override_heading_atan2) {
  std::vector<std::array<double, 4>> out;
  for (size_t it = 0; it < traj.size(); it++) {
    std::array<double, 4> pt{traj[it].p[0], traj[it].p[1], traj[it].p[2], 0.0};
    if (override_heading_atan2 && it < (traj.size() - 1)) {
      const double points_dist = m_hypot(traj[it + 1].p[1] - pt[1], traj[it + 1].p[0] - pt[0]);
      if (points_dist < 0.05 && it > 0) pt[3] = out[it - 1][3];
      else pt[3] = m_atan2(traj[it + 1].p[1] - pt[1], traj[it + 1].p[0] - pt[0]);
    } else {
      pt[3] = traj[it].yaw_out;
    }
    out.push_back(pt);
  }
  return out;
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

// node.cpp:431-500
std::vector<Waypoint> preprocess_path(const std::vector<Waypoint>& in, double min_waypoint_distance, bool straightener_enabled,
                                      double straightener_max_deviation, double straightener_max_hdg_deviation) {
  std::vector<Waypoint> out;
  size_t last_added_idx = 0;
  for (size_t i = 0; i < in.size(); ++i) {
    if (straightener_enabled && in.size() >= 3 && i > 0 && i < (in.size() - 1)) {
      const double* first = in[last_added_idx].c;
      const double* last = in[i + 1].c;
      const double first_hdg = first[3], last_hdg = last[3];
      bool segment_is_ok = true;
      for (size_t j = last_added_idx + 1; j < i + 1; ++j) {
        const double* mid = in[j].c;
        const double mid_hdg = mid[3];
        const double d = dist_from_segment(mid, first, last);
        // as written in the reference: fabs() of the comparison's bool
        if (d > straightener_max_deviation || std::fabs((double)(rad_diff(first_hdg, mid_hdg) > straightener_max_hdg_deviation)) != 0.0 ||
            std::fabs((double)(rad_diff(last_hdg, mid_hdg) > straightener_max_hdg_deviation)) != 0.0) {
          segment_is_ok = false;
          break;
        }
      }
      if (segment_is_ok) continue;
    }
    if (i > 0 && i < (in.size() - 1)) {
      const double* first = in[last_added_idx].c;
      const double* last = in[i].c;
      const double dx = first[0] - last[0], dy = first[1] - last[1], dz = first[2] - last[2];
      if (std::sqrt(dx * dx + dy * dy + dz * dz) < min_waypoint_distance) continue;  // mrs_lib::geometry::dist
    }
    out.push_back(in[i]);
    last_added_idx = i;
  }
  return out;
}

// node.cpp:1215-1395
void fallback_sample(const std::vector<Waypoint>& wp, const Limits& L, double dt, double stopping_time, std::vector<std::array<double, 4>>* out) {
  out->clear();
  if (wp.size() < 2) return;
  std::vector<Vertex> vertices(wp.size());
  double last_heading = wp[0].c[3];
  for (size_t i = 0; i < wp.size(); ++i) {
    const double heading = srad_unwrap(wp[i].c[3], last_heading);
    last_heading = heading;
    vertices[i].mask = 1;
    vertices[i].val[0][0] = wp[i].c[0];
    vertices[i].val[0][1] = wp[i].c[1];
    vertices[i].val[0][2] = wp[i].c[2];
    vertices[i].val[0][3] = heading;
  }
  const std::vector<double> baca = estimate_times_baca(vertices, L);
  for (size_t i = 0; i + 1 < wp.size(); ++i) {
    const double segment_time = baca[i];
    int n_samples;
    double interp_step;
    if (segment_time > 1e-1) {
      n_samples = (int)std::ceil(segment_time / dt);
      interp_step = (n_samples > 0) ? 1.0 / (double)n_samples : 0.5;
    } else {
      n_samples = 0;
      interp_step = 0;
    }
    if (n_samples > 0 && i == wp.size() - 2) n_samples++;  // the last segment hits the last waypoint
    for (int j = 0; j < n_samples; ++j) {
      const Waypoint pt = interpolate_point(wp[i], wp[i + 1], (double)j * interp_step);
      // setFromYaw -> quaternion -> the heading getTrajectoryReference reads back (eth_mav_msgs/common.h:130-140)
      const double ha = 0.5 * pt.c[3];
      const double qw = m_cos_k(ha), qz = m_sin_k(ha);
      const double yaw = m_atan2_k(2.0 * (qw * qz + 0.0 * 0.0), 1.0 - 2.0 * (0.0 * 0.0 + qz * qz));
      const std::array<double, 4> smp = {pt.c[0], pt.c[1], pt.c[2], yaw};
      out->push_back(smp);
      if (j == 0 && i > 0 && wp[i].stop_at) {
        const int insert_samples = (int)std::round(stopping_time / dt);
        for (int k = 0; k < insert_samples; ++k) out->push_back(smp);
      }
    }
  }
}

// node.cpp:1461-1499
std::vector<int> waypoint_trajectory_idxs(const std::vector<std::array<double, 4>>& samples, const std::vector<Waypoint>& wp) {
  std::vector<int> idxs;
  int waypoint_idx = 0;
  if (samples.empty() || wp.empty()) return idxs;
  for (size_t i = 0; i + 1 < samples.size(); ++i) {
    const double d = dist_from_segment(wp[waypoint_idx].c, samples[i].data(), samples[i + 1].data());
    if (d < 0.1) {
      idxs.push_back((int)i);
      waypoint_idx++;
    }
    if (waypoint_idx == (int)wp.size()) break;
  }
  return idxs;
}

// node.cpp:620-851, numeric part
OptimizeResult optimize_path(const std::vector<Waypoint>& wp_in, const InitialState& init, const NodeParams& P) {
  OptimizeResult O;
  O.wp = wp_in;
  // checkNaN (node.cpp:1896-1900, isnan / isinf on the four coordinates): the callbacks drop such a message before optimize() runs
  bool finite = true;
  if (O.wp.size() > 1) {
    for (const Waypoint& w : O.wp)
      for (int k = 0; k < 4; ++k) finite = finite && std::isfinite(w.c[k]);
    if (init.present) {
      finite = finite && std::isfinite(init.heading);
      for (int k = 0; k < 4; ++k) finite = finite && std::isfinite(init.vel[k]) && std::isfinite(init.acc[k]) && std::isfinite(init.jerk[k]);
    }
  }
  if (!finite) {
    O.find.status = kFindNotFinite;
    return O;
  }
  if (O.wp.size() <= 1) {  // "the path is empty (after postprocessing)" (node.cpp:676-681)
    O.find.status = kFindEmptyPath;
    return O;
  }
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
