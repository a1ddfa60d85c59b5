// oracle/capi.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).  C entry points for ctypes (tests, bench cpu_baseline).
#include <atomic>
#include <cstring>
#include <thread>

#include "oracle.h"

using namespace orc;

extern "C" {

// Mirrors include/tg_b200.h : tg_params (same field order so tests can share one ctypes.Structure).
struct orc_params {
  int derivative_to_optimize;
  int max_evals;
  double f_rel, x_rel;
  double limits[9];  // v_h, v_v, a_h, a_v, j_h, j_v, v_hdg, a_hdg, j_hdg
  double dt;
  int check_deviation;
  double max_deviation;
  int max_deviation_iters;
  int first_segment_checked;
  double max_len_factor, min_len_factor;
  int run_time_alloc;
  int override_heading_atan2;
};

static NodeParams to_node(const orc_params* p) {
  NodeParams P;
  P.derivative_to_optimize = p->derivative_to_optimize;
  P.nl.max_evals = p->max_evals;
  P.nl.f_rel = p->f_rel;
  P.nl.x_rel = p->x_rel;
  const double* l = p->limits;
  P.lim = Limits{l[0], l[1], l[2], l[3], l[4], l[5], l[6], l[7], l[8]};
  P.dt = p->dt;
  P.check_deviation = p->check_deviation != 0;
  P.max_deviation = p->max_deviation;
  P.max_deviation_iters = p->max_deviation_iters;
  P.first_segment_checked = p->first_segment_checked != 0;
  P.max_len_factor = p->max_len_factor;
  P.min_len_factor = p->min_len_factor;
  P.run_time_alloc = p->run_time_alloc != 0;
  P.override_heading_atan2 = p->override_heading_atan2 != 0;
  return P;
}

void orc_set_math_mode(int mode) { set_math_mode(mode); }
void orc_set_scale_tolerance(double tol) { g_scale_tolerance = tol; }
int orc_get_math_mode() { return math_mode(); }

double orc_math(int fn, double x, double y) {
  switch (fn) {
    case 0: return m_log(x);
    case 1: return m_exp(x);
    case 2: return m_sin(x);
    case 3: return m_cos(x);
    case 4: return m_atan2(x, y);
    case 5: return m_cbrt(x);
    case 6: return m_pow_int(x, (int)y);
    case 7: return m_hypot(x, y);
  }
  return 0;
}

void orc_segment_matrices(double T, int r, double* A, double* Ainv, double* Q) {
  setup_mapping_A(T, A);
  invert_mapping(A, Ainv);
  cost_jacobian_Q(r, T, Q);
}

static std::vector<Vertex> make_vertices(int V, const uint8_t* mask, const double* vals) {
  std::vector<Vertex> vs(V);
  for (int v = 0; v < V; ++v) {
    vs[v].mask = mask[v];
    for (int k = 0; k < kHalf; ++k)
      for (int d = 0; d < kD; ++d) vs[v].val[k][d] = vals[((size_t)v * kHalf + k) * kD + d];
  }
  return vs;
}

// one linear solve; coeffs: S x 4 x 10 ; dp: 4 x n_free (may be null) ; dims[2] = {n_fixed, n_free}
int orc_solve_linear(int V, const uint8_t* mask, const double* vals, const double* times, int r, double* coeffs,
                     double* cost, double* dp, int* dims) {
  LinearSolver ls;
  std::vector<double> t(times, times + (V - 1));
  if (!ls.setup(make_vertices(V, mask, vals), t, r)) return 1;
  ls.solve();
  for (int i = 0; i < ls.S; ++i)
    for (int d = 0; d < kD; ++d) std::memcpy(coeffs + ((size_t)i * kD + d) * kN, ls.seg[i].c[d], sizeof(double) * kN);
  *cost = ls.cost();
  if (dims) {
    dims[0] = ls.n_fixed;
    dims[1] = ls.n_free;
  }
  if (dp) std::memcpy(dp, ls.d_p.data(), sizeof(double) * ls.d_p.size());
  return 0;
}

// PolynomialOptimizationNonLinear::optimize() from vertices: Mellinger loop + time scaling + final solve
int orc_time_alloc(int V, const uint8_t* mask, const double* vals, double* times, int r, const orc_params* prm, double* coeffs, int* code,
                   int* n_evals, int* n_scale_passes, double* final_cost) {
  LinearSolver ls;
  std::vector<double> t(times, times + (V - 1));
  if (!ls.setup(make_vertices(V, mask, vals), t, r)) return 1;
  const NodeParams np = to_node(prm);
  NlInfo info;
  optimize_time_mellinger(ls, np.nl, np.lim, &info);
  for (int i = 0; i < ls.S; ++i) {
    times[i] = ls.seg[i].T;
    for (int d = 0; d < kD; ++d) std::memcpy(coeffs + ((size_t)i * kD + d) * kN, ls.seg[i].c[d], sizeof(double) * kN);
  }
  *code = info.code;
  *n_evals = info.n_evals;
  *n_scale_passes = info.n_scale_passes;
  *final_cost = info.final_cost;
  return 0;
}

// the steps either side of the path
static std::vector<Waypoint> make_waypoints(int V, const double* wp, const uint8_t* stop) {
  std::vector<Waypoint> w(V);
  for (int i = 0; i < V; ++i) {
    for (int d = 0; d < 4; ++d) w[i].c[d] = wp[4 * i + d];
    w[i].stop_at = stop ? stop[i] != 0 : false;
  }
  return w;
}
int orc_preprocess_path(int V, const double* wp, const uint8_t* stop, double min_dist, int straighten, double max_dev, double max_hdg_dev, double* out_wp,
                        uint8_t* out_stop) {
  const std::vector<Waypoint> o = preprocess_path(make_waypoints(V, wp, stop), min_dist, straighten != 0, max_dev, max_hdg_dev);
  for (size_t i = 0; i < o.size(); ++i) {
    for (int d = 0; d < 4; ++d) out_wp[4 * i + d] = o[i].c[d];
    out_stop[i] = o[i].stop_at ? 1 : 0;
  }
  return (int)o.size();
}
int orc_fallback_sample(int V, const double* wp, const uint8_t* stop, const double* limits9, double dt, double stopping_time, int cap, double* out) {
  Limits L{limits9[0], limits9[1], limits9[2], limits9[3], limits9[4], limits9[5], limits9[6], limits9[7], limits9[8]};
  std::vector<std::array<double, 4>> s;
  fallback_sample(make_waypoints(V, wp, stop), L, dt, stopping_time, &s);
  if ((int)s.size() > cap) return -(int)s.size();
  for (size_t i = 0; i < s.size(); ++i)
    for (int d = 0; d < 4; ++d) out[4 * i + d] = s[i][d];
  return (int)s.size();
}
int orc_waypoint_idxs(int M, const double* samples, int V, const double* wp, int* idxs) {
  std::vector<std::array<double, 4>> s(M);
  for (int i = 0; i < M; ++i)
    for (int d = 0; d < 4; ++d) s[i][d] = samples[4 * i + d];
  const std::vector<int> o = waypoint_trajectory_idxs(s, make_waypoints(V, wp, nullptr));
  for (size_t i = 0; i < o.size(); ++i) idxs[i] = o[i];
  return (int)o.size();
}

// dense R for structure tests: R is (n_fixed+n_free)^2 row-major
int orc_dense_R(int V, const uint8_t* mask, const double* vals, const double* times, int r, double* R) {
  LinearSolver ls;
  std::vector<double> t(times, times + (V - 1));
  if (!ls.setup(make_vertices(V, mask, vals), t, r)) return 1;
  std::vector<double> Rv;
  ls.dense_R(&Rv);
  std::memcpy(R, Rv.data(), sizeof(double) * Rv.size());
  return 0;
}

int orc_find_roots(const double* coeffs_increasing, int n, double* re, double* im, int* ok) {
  bool o = true;
  const int nr = find_roots_jt(coeffs_increasing, n, re, im, &o);
  *ok = o ? 1 : 0;
  return nr;
}

static std::vector<Segment> make_segments(int S, const double* coeffs, const double* times) {
  std::vector<Segment> seg(S);
  for (int i = 0; i < S; ++i) {
    seg[i].T = times[i];
    for (int d = 0; d < kD; ++d) std::memcpy(seg[i].c[d], coeffs + ((size_t)i * kD + d) * kN, sizeof(double) * kN);
  }
  return seg;
}

// nine maxima per segment: out[S*9], order hv ha hj vv va vj yv ya yj
void orc_segment_maxima(int S, const double* coeffs, const double* times, double* out) {
  std::vector<Segment> seg = make_segments(S, coeffs, times);
  const int hor[2] = {0, 1}, ver[1] = {2}, hdg[1] = {3};
  for (int i = 0; i < S; ++i) {
    for (int k = 1; k <= 3; ++k) out[i * 9 + k - 1] = segment_max_magnitude(seg[i], k, hor, 2, nullptr);
    for (int k = 1; k <= 3; ++k) out[i * 9 + 3 + k - 1] = segment_max_magnitude(seg[i], k, ver, 1, nullptr);
    for (int k = 1; k <= 3; ++k) out[i * 9 + 6 + k - 1] = segment_max_magnitude(seg[i], k, hdg, 1, nullptr);
  }
}

// computeMaximumOfMagnitude(derivative) of one trajectory: Extremum {time, value, segment_idx}
void orc_max_magnitude(int S, const double* coeffs, const double* times, int derivative, double* time, double* value, int* segment_idx) {
  std::vector<Segment> seg = make_segments(S, coeffs, times);
  max_of_magnitude(seg, derivative, time, value, segment_idx);
}

// in-place time scaling; returns passes
int orc_scale_times(int S, double* coeffs, double* times, const double* limits, int* within) {
  std::vector<Segment> seg = make_segments(S, coeffs, times);
  const Limits L{limits[0], limits[1], limits[2], limits[3], limits[4], limits[5], limits[6], limits[7], limits[8]};
  bool w = false;
  const int passes = scale_times_to_meet_constraints(seg, L, &w, nullptr);
  *within = w ? 1 : 0;
  for (int i = 0; i < S; ++i) {
    times[i] = seg[i].T;
    for (int d = 0; d < kD; ++d) std::memcpy(coeffs + ((size_t)i * kD + d) * kN, seg[i].c[d], sizeof(double) * kN);
  }
  return passes;
}

void orc_estimate_times(int V, const double* pos4, const double* limits, double* euclid, double* baca) {
  std::vector<Vertex> vs(V);
  for (int v = 0; v < V; ++v) vs[v].add(0, pos4 + 4 * v);
  const Limits L{limits[0], limits[1], limits[2], limits[3], limits[4], limits[5], limits[6], limits[7], limits[8]};
  const std::vector<double> e = estimate_times_euclidean(vs, L), b = estimate_times_baca(vs, L);
  for (int i = 0; i < V - 1; ++i) {
    euclid[i] = e[i];
    baca[i] = b[i];
  }
}

// samples: rows of 22 doubles [p4 v4 a4 j3 s3 yaw_out t_s t_in_unused -> padded]; returns count (or -needed if cap small)
static void pack_sample(const Sample& s, double* o, int64_t* tns) {
  for (int d = 0; d < 4; ++d) { o[d] = s.p[d]; o[4 + d] = s.v[d]; o[8 + d] = s.a[d]; }
  for (int d = 0; d < 3; ++d) { o[12 + d] = s.j[d]; o[15 + d] = s.s[d]; }
  o[18] = s.yaw_out;
  if (tns) *tns = s.t_ns;
}
int orc_sample(int S, const double* coeffs, const double* times, double dt, int cap, double* out19, int64_t* t_ns) {
  std::vector<Segment> seg = make_segments(S, coeffs, times);
  std::vector<Sample> sm;
  sample_whole(seg, dt, &sm);
  if ((int)sm.size() > cap) return -(int)sm.size();
  for (size_t i = 0; i < sm.size(); ++i) pack_sample(sm[i], out19 + 19 * i, t_ns ? t_ns + i : nullptr);
  return (int)sm.size();
}

int orc_trajectory_evaluate(int S, const double* coeffs, const double* times, double t, int deriv, double* out4) {
  std::vector<Segment> seg = make_segments(S, coeffs, times);
  return trajectory_evaluate(seg, t, deriv, out4) ? 1 : 0;
}

double orc_dist_from_segment(const double* p, const double* a, const double* b) { return dist_from_segment(p, a, b); }

void orc_cyclic(double a, double b, double c, double* out4) {
  out4[0] = rad_diff(a, b);
  out4[1] = rad_interp(a, b, c);
  out4[2] = srad_unwrap(a, b);
  out4[3] = rad_wrap(a);
}

// Per-problem result record shared by the single and batch entry points.
struct orc_result {
  int status;        // FindStatus of the last findTrajectory
  int success;       // optimize() success
  int nlopt_code;
  int n_evals;       // of the last findTrajectory
  int rounds;        // subdivision rounds
  int safe;
  int n_waypoints;   // final
  int n_samples;     // final
  int n_scale_passes;
  int overflow;      // 1 if an output did not fit its capacity
  double max_dev;
  double final_cost;
  double baca_total;
  long long total_solves, total_root_calls, total_evals;
};

static InitialState to_init(const double* init14) {
  InitialState I;
  if (!init14 || init14[0] == 0.0) return I;
  I.present = true;
  I.heading = init14[1];
  for (int d = 0; d < 4; ++d) {
    I.vel[d] = init14[2 + d];
    I.acc[d] = init14[6 + d];
    I.jerk[d] = init14[10 + d];
  }
  return I;
}

// Full optimize() (findTrajectory + validation + subdivision).  wp: V x 4, stop_at: V bytes.
// Outputs (all optional except res): wp_out (cap_wp x 4), times (cap_wp-1), coeffs ((cap_wp-1) x 40),
// samples_xyzh (cap_samples x 4: x y z heading as getTrajectoryReference emits them, node.cpp:1578-1602).
int orc_optimize_path(int V, const double* wp, const uint8_t* stop_at, const double* init14, const orc_params* prm,
                      orc_result* res, int cap_wp, double* wp_out, double* times, double* coeffs, int cap_samples,
                      double* samples_xyzh) {
  std::vector<Waypoint> w(V);
  for (int i = 0; i < V; ++i) {
    for (int d = 0; d < 4; ++d) w[i].c[d] = wp[4 * i + d];
    w[i].stop_at = stop_at ? stop_at[i] != 0 : false;
  }
  const NodeParams P = to_node(prm);
  OptimizeResult O = optimize_path(w, to_init(init14), P);
  if (O.find.status == kFindEmptyPath || O.find.status == kFindNotFinite) {  // no trajectory: nothing is returned for the path
    O.wp.clear();
    O.find.nl.code = -1;
  }
  std::memset(res, 0, sizeof(*res));
  res->status = O.find.status;
  res->success = O.success;
  res->nlopt_code = O.find.nl.code;
  res->n_evals = O.find.nl.n_evals;
  res->rounds = O.rounds;
  res->safe = O.safe;
  res->n_waypoints = (int)O.wp.size();
  res->n_samples = (int)O.find.samples.size();
  res->n_scale_passes = O.find.nl.n_scale_passes;
  res->max_dev = O.max_dev;
  res->final_cost = O.find.nl.final_cost;
  res->baca_total = O.find.baca_total;
  res->total_solves = O.total_solves;
  res->total_root_calls = O.total_root_calls;
  res->total_evals = O.total_evals;
  const int Vf = (int)O.wp.size();
  if (Vf > cap_wp || (int)O.find.samples.size() > cap_samples) {
    res->overflow = 1;
    return 0;
  }
  if (wp_out)
    for (int i = 0; i < Vf; ++i)
      for (int d = 0; d < 4; ++d) wp_out[4 * i + d] = O.wp[i].c[d];
  const int Sf = (int)O.find.seg.size();
  if (times)
    for (int i = 0; i < Sf; ++i) times[i] = O.find.seg[i].T;
  if (coeffs)
    for (int i = 0; i < Sf; ++i)
      for (int d = 0; d < kD; ++d) std::memcpy(coeffs + ((size_t)i * kD + d) * kN, O.find.seg[i].c[d], sizeof(double) * kN);
  if (samples_xyzh) {
    const std::vector<std::array<double, 4>> ref = trajectory_reference(O.find.samples, to_node(prm).override_heading_atan2);
    for (size_t i = 0; i < ref.size(); ++i)
      for (int d = 0; d < 4; ++d) samples_xyzh[4 * i + d] = ref[i][d];
  }
  return 0;
}

// Batch over B problems with ragged waypoint lists (wp_off[B+1]); outputs use fixed per-problem capacities.
// nthreads <= 0 -> hardware_concurrency.  This is the CPU baseline timed by bench.py.
int orc_optimize_batch(int B, const int* wp_off, const double* wp, const uint8_t* stop_at, const double* init14,
                       const orc_params* prm, orc_result* res, int cap_wp, double* wp_out, double* times, double* coeffs,
                       int cap_samples, double* samples_xyzh, int nthreads) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      const int o = wp_off[b], V = wp_off[b + 1] - o;
      orc_optimize_path(V, wp + 4 * (size_t)o, stop_at ? stop_at + o : nullptr, init14 ? init14 + 14 * (size_t)b : nullptr, prm,
                        res + b, cap_wp, wp_out ? wp_out + (size_t)b * cap_wp * 4 : nullptr,
                        times ? times + (size_t)b * (cap_wp - 1) : nullptr,
                        coeffs ? coeffs + (size_t)b * (cap_wp - 1) * kD * kN : nullptr, cap_samples,
                        samples_xyzh ? samples_xyzh + (size_t)b * cap_samples * 4 : nullptr);
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return nthreads;
}

// Time-allocation sweep (BASELINE config 5): cost of one problem at K candidate time vectors.
void orc_sweep_costs(int V, const uint8_t* mask, const double* vals, int r, int K, const double* cand_times, double* costs,
                     int nthreads) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  const int S = V - 1;
  std::atomic<int> next(0);
  const std::vector<Vertex> vs = make_vertices(V, mask, vals);
  auto work = [&]() {
    LinearSolver ls;
    bool ready = false;
    for (;;) {
      const int k0 = next.fetch_add(64);
      if (k0 >= K) break;
      for (int k = k0; k < std::min(K, k0 + 64); ++k) {
        std::vector<double> t(cand_times + (size_t)k * S, cand_times + (size_t)(k + 1) * S);
        if (!ready) { ls.setup(vs, t, r); ready = true; } else ls.update_times(t);
        ls.solve();
        costs[k] = ls.cost();
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}


// The objective functions of the time-allocation methods other than Mellinger's (nl_impl.h:567-614 objectiveFunctionTime,
// 651-722 objectiveFunctionTimeAndConstraints, 740-762 evaluateMaximumMagnitudeAsSoftConstraint, 724-738
// evaluateMaximumMagnitudeConstraint -> computeMaximumOfMagnitude over all dimensions), for K candidate vectors x of ONE
// problem.  method: 0 kSquaredTime, 1 kRichterTime (x = S segment times: updateSegmentTimes + solveLinear), 3 / 4 the
// ...AndConstraints variants (x = S times then kD * n_free free constraints: updateSegmentTimes + setFreeConstraints).
// parts[k] = {cost_trajectory, cost_time, cost_soft_constraints}; total = their sum in that order (nl_impl.h:613, 721).
void orc_objective(int V, const uint8_t* mask, const double* vals, int r, int method, int K, const double* x, int nvar, double time_penalty,
                   int use_soft, double soft_weight, int ncon, const int* con_deriv, const double* con_value, double* total, double* parts,
                   int nthreads) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  const int S = V - 1;
  std::atomic<int> next(0);
  const std::vector<Vertex> vs = make_vertices(V, mask, vals);
  auto work = [&]() {
    LinearSolver ls;
    bool ready = false;
    for (;;) {
      const int k0 = next.fetch_add(16);
      if (k0 >= K) break;
      for (int k = k0; k < std::min(K, k0 + 16); ++k) {
        const double* xk = x + (size_t)k * nvar;
        std::vector<double> t(xk, xk + S);
        if (!ready) { ls.setup(vs, t, r); ready = true; } else ls.update_times(t);
        if (method >= 3) {
          for (int d = 0; d < kD; ++d)
            for (int i = 0; i < ls.n_free; ++i) ls.d_p[(size_t)d * ls.n_free + i] = xk[S + (size_t)d * ls.n_free + i];
          ls.segments_from_compact();
        } else {
          ls.solve();
        }
        const double cost_traj = ls.cost();
        double total_time = 0;
        for (int i = 0; i < S; ++i) total_time += t[i];
        const double cost_time = (method == 1 || method == 4) ? total_time * time_penalty : total_time * total_time * time_penalty;
        double cost_con = 0;
        if (use_soft) {
          for (int c = 0; c < ncon; ++c) {
            double tm, val;
            int idx;
            max_of_magnitude(ls.seg, con_deriv[c], &tm, &val, &idx);
            const double abs_violation = val - con_value[c];
            const double relative_violation = abs_violation / con_value[c];
            cost_con += std::min(1.0e12, m_exp(relative_violation * soft_weight));
          }
        }
        if (parts) {
          parts[(size_t)k * 3 + 0] = cost_traj;
          parts[(size_t)k * 3 + 1] = cost_time;
          parts[(size_t)k * 3 + 2] = cost_con;
        }
        total[k] = cost_traj + cost_time + cost_con;
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}

}  // extern "C"
