// oracle/trajectory.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates eth/polynomial.{h,cpp}, eth/segment.cpp, eth/trajectory.cpp, eth/trajectory_sampling.cpp and the
// segment-time heuristics of eth/vertex.cpp.
#include <cfloat>
#include <cmath>
#include <limits>

#include "oracle.h"

namespace orc {

// eth/polynomial.h:150-163 : Horner from the top, one multiply and one add per step.
double poly_eval(const double* c, double t, int deriv) {
  if (deriv >= kN) return 0.0;
  const int top = kN - 1;
  double acc = base_coeff(deriv, top) * c[top];
  for (int j = top - 1; j >= deriv; --j) {
    acc *= t;
    acc += base_coeff(deriv, j) * c[j];
  }
  return acc;
}

// eth/polynomial.h:108-119
void poly_deriv_coeffs(const double* c, int deriv, double* out) {
  for (int j = 0; j < kN; ++j) out[j] = 0.0;
  if (deriv == 0) {
    for (int j = 0; j < kN; ++j) out[j] = c[j];
    return;
  }
  for (int j = 0; j < kN - deriv; ++j) out[j] = c[j + deriv] * base_coeff(deriv, j + deriv);
}

// eth/polynomial.cpp:176-192 : convolve(data, kernel)[m] = sum data[m-n] kernel[n], reversed-kernel loop order.
static void convolve(const double* data, int nd, const double* kernel, int nk, double* out) {
  const int len = nd + nk - 1;
  for (int i = 0; i < len; ++i) {
    out[i] = 0.0;
    const int data_idx = i - nk + 1;
    const int lower = std::max(0, -data_idx);
    const int upper = std::min(nk, nd - data_idx);
    for (int kidx = lower; kidx < upper; ++kidx) out[i] += kernel[nk - 1 - kidx] * data[data_idx + kidx];
  }
}

// eth/segment.cpp:113-156 + eth/polynomial.cpp:36-85 : candidate times on [0, T]
static int candidate_times(const Segment& s, int deriv, const int* dims, int ndims, double* cand, long* root_calls) {
  double re[2 * kN], im[2 * kN];
  int nroots = 0;
  bool ok = true;
  if (ndims > 1) {
    const int n_d = kN - deriv, n_dd = n_d - 1;
    const int len = n_d + n_dd - 1;
    double conv[2 * kN], acc[2 * kN];
    for (int i = 0; i < len; ++i) acc[i] = 0.0;
    for (int q = 0; q < ndims; ++q) {
      double d[kN], dd[kN];
      poly_deriv_coeffs(s.c[dims[q]], deriv, d);
      poly_deriv_coeffs(s.c[dims[q]], deriv + 1, dd);
      convolve(d, n_d, dd, n_dd, conv);
      for (int i = 0; i < len; ++i) acc[i] += conv[i];
    }
    // computeMinMaxCandidates(t_start, t_end, -1): roots of getCoefficients(0) of the convolved polynomial
    nroots = find_roots_jt(acc, len, re, im, &ok);
  } else {
    // one dimension: roots of the (deriv+1)-th derivative, an N-vector with trailing zeros (polynomial.cpp:69-85)
    double dd[kN];
    poly_deriv_coeffs(s.c[dims[0]], deriv + 1, dd);
    nroots = find_roots_jt(dd, kN, re, im, &ok);
  }
  if (root_calls) ++*root_calls;
  // selectMinMaxCandidatesFromRoots (polynomial.cpp:36-63)
  int n = 0;
  const double t_start = 0.0, t_end = s.T;
  if (t_start > t_end) return 0;
  cand[n++] = t_start;
  cand[n++] = t_end;
  for (int i = 0; i < nroots; ++i) {
    if (std::fabs(im[i]) > DBL_EPSILON) continue;
    const double c = re[i];
    if (c < t_start || c > t_end) continue;
    cand[n++] = c;
  }
  return n;
}

// eth/segment.cpp:162-212 + trajectory.cpp:211-243 : maximum of the magnitude over the candidates
double segment_max_magnitude(const Segment& s, int deriv, const int* dims, int ndims, long* root_calls) {
  double cand[2 * kN + 2];
  const int n = candidate_times(s, deriv, dims, ndims, cand, root_calls);
  double best = std::numeric_limits<double>::lowest();
  for (int i = 0; i < n; ++i) {
    double mag = 0.0;
    for (int q = 0; q < ndims; ++q) {
      const double v = poly_eval(s.c[dims[q]], cand[i], deriv);
      mag += v * v;  // std::pow(x, 2) is exactly x*x
    }
    mag = std::sqrt(mag);
    if (cand[i] < 0.0 || cand[i] > s.T) continue;
    if (best < mag) best = mag;  // std::max(*maximum, candidate) with operator< on value
  }
  return best;
}

// PolynomialOptimization<N>::computeMaximumOfMagnitude(derivative, nullptr) (lin_impl.h:477-508): the candidate list of a
// segment is computeSegmentMaximumMagnitudeCandidates' output -- t_start, t_end, then the real zeros inside the segment
// (the 0.0 pushed at lin_impl.h:487 is cleared again by segment.cpp:117) -- over ALL dimensions (lin_impl.h:407-409);
// value = Segment::evaluate(t, k).norm(), summed here in dimension order (Eigen's reduction order for a 4-vector is
// third-party arithmetic: parity unpinned, oracle.h); the running Extremum starts at (0, 0, 0) and is replaced on a
// strictly larger value, so the first of equal maxima wins.  The closing candidate (end of the last segment,
// lin_impl.h:501-505) repeats one already seen and cannot win.
void max_of_magnitude(const std::vector<Segment>& seg, int deriv, double* time, double* value, int* segment_idx) {
  double best = 0.0, best_t = 0.0;
  int best_i = 0;
  const int dims[4] = {0, 1, 2, 3};
  for (size_t si = 0; si < seg.size(); ++si) {
    const Segment& s = seg[si];
    double cand[2 * kN + 2];
    const int n = candidate_times(s, deriv, dims, kD, cand, nullptr);
    for (int i = 0; i < n; ++i) {
      double mag = 0.0;
      for (int q = 0; q < kD; ++q) {
        const double v = poly_eval(s.c[q], cand[i], deriv);
        mag += v * v;
      }
      mag = std::sqrt(mag);
      if (best < mag) {
        best = mag;
        best_t = cand[i];
        best_i = (int)si;
      }
    }
  }
  if (!seg.empty()) {
    const Segment& s = seg.back();
    double mag = 0.0;
    for (int q = 0; q < kD; ++q) {
      const double v = poly_eval(s.c[q], s.T, deriv);
      mag += v * v;
    }
    mag = std::sqrt(mag);
    if (best < mag) {
      best = mag;
      best_t = s.T;
      best_i = (int)seg.size() - 1;
    }
  }
  *time = best_t;
  *value = best;
  *segment_idx = best_i;
}

static void nine_maxima(const Segment& s, double* m, long* rc) {
  const int hor[2] = {0, 1}, ver[1] = {2}, hdg[1] = {3};
  // order of calls: horizontal v,a,j ; vertical v,a,j ; heading v,a,j (trajectory.cpp:616-622)
  for (int k = 1; k <= 3; ++k) m[k - 1] = segment_max_magnitude(s, k, hor, 2, rc);
  for (int k = 1; k <= 3; ++k) m[3 + k - 1] = segment_max_magnitude(s, k, ver, 1, rc);
  for (int k = 1; k <= 3; ++k) m[6 + k - 1] = segment_max_magnitude(s, k, hdg, 1, rc);
}

double g_scale_tolerance = 1e-3;  // test hook (orc_set_scale_tolerance): lets tests force the rare multi-pass path

// eth/trajectory.cpp:598-692
int scale_times_to_meet_constraints(std::vector<Segment>& seg, const Limits& L, bool* within_out, long* rc) {
  constexpr int kMaxCounter = 20;
  const double kTolerance = g_scale_tolerance;  // 1e-3 in the reference (eth/trajectory.cpp:604); tests may change it
  bool within = false;
  int passes = 0;
  for (int it = 0; it < kMaxCounter; ++it) {
    ++passes;
    for (size_t si = 0; si < seg.size(); ++si) {
      double m[9];
      nine_maxima(seg[si], m, rc);
      const double vv_h = m[0] / L.v_h, vv_v = m[3] / L.v_v, av_h = m[1] / L.a_h, av_v = m[4] / L.a_v;
      const double jv_h = m[2] / L.j_h, jv_v = m[5] / L.j_v;
      const double vv_y = m[6] / L.v_hdg, av_y = m[7] / L.a_hdg, jv_y = m[8] / L.j_hdg;
      const double vviol = std::max(std::max(vv_h, vv_v), vv_y);
      const double aviol = std::max(std::max(av_h, av_v), av_y);
      const double jviol = std::max(std::max(jv_h, jv_v), jv_y);
      const double scaling = std::max(1.0, std::max(std::max(vviol, std::sqrt(aviol)), m_cbrt(jviol)));
      const double inv = 1.0 / scaling;
      const double new_time = seg[si].T * scaling;
      for (int d = 0; d < kD; ++d) {  // scalePolynomialInTime (polynomial.cpp:218-224)
        double scale = 1.0;
        for (int n = 0; n < kN; ++n) {
          seg[si].c[d][n] *= scale;
          scale *= inv;
        }
      }
      seg[si].T = new_time;
    }
    // global re-check over the whole trajectory (trajectory.cpp:660-689)
    double g[9];
    for (int q = 0; q < 9; ++q) g[q] = std::numeric_limits<double>::lowest();
    for (size_t si = 0; si < seg.size(); ++si) {
      double m[9];
      // computeMinMaxMagnitude over all segments is called once per (group, derivative); the per-segment work
      // is identical, only the loop nesting differs, so the same nine values per segment are produced.
      nine_maxima(seg[si], m, rc);
      for (int q = 0; q < 9; ++q)
        if (m[q] > g[q]) g[q] = m[q];
    }
    const double vviol = std::max(std::max(g[0] / L.v_h, g[3] / L.v_v), g[6] / L.v_hdg);
    const double aviol = std::max(std::max(g[1] / L.a_h, g[4] / L.a_v), g[7] / L.a_hdg);
    const double jviol = std::max(std::max(g[2] / L.j_h, g[5] / L.j_v), g[8] / L.j_hdg);
    within = vviol <= 1.0 + kTolerance && aviol <= 1.0 + kTolerance && jviol <= 1.0 + kTolerance;
    if (within) break;
  }
  *within_out = within;
  return passes;
}

// eth/trajectory.cpp:55-87
bool trajectory_evaluate(const std::vector<Segment>& seg, double t, int deriv, double* out) {
  double acc = 0.0;
  size_t i = 0;
  for (i = 0; i < seg.size(); ++i) {
    acc += seg[i].T;
    if (acc > t) break;
  }
  if (t > acc) {
    for (int d = 0; d < kD; ++d) out[d] = 0.0;
    return false;
  }
  if (i >= seg.size()) i = seg.size() - 1;
  acc -= seg[i].T;
  for (int d = 0; d < kD; ++d) out[d] = poly_eval(seg[i].c[d], t - acc, deriv);
  return true;
}

// eth/trajectory.cpp:93-151 for one derivative order: emits (segment index, time in segment) pairs.
// The walk is identical for all five orders, so it is done once and the five Horner evaluations follow.
bool sample_whole(const std::vector<Segment>& seg, double dt, std::vector<Sample>* out) {
  out->clear();
  double t_end = 0.0;  // Trajectory::max_time_ (eth/trajectory.h:76-83)
  for (const Segment& s : seg) t_end += s.T;
  const double t_start = 0.0;
  double acc = 0.0;
  size_t i = 0;
  for (i = 0; i < seg.size(); ++i) {
    acc += seg[i].T;
    if (acc > t_start) break;
  }
  if (t_start > acc) return true;  // evaluateRange logs and returns; sampleTrajectoryInRange still reports true
  if (i >= seg.size()) return true;
  acc -= seg[i].T;
  double tin = t_start - acc;
  while (acc < t_end) {
    if (tin > seg[i].T) {
      tin = tin - seg[i].T;
      i++;
      if (i >= seg.size()) break;
      continue;
    }
    Sample sm;
    for (int d = 0; d < kD; ++d) {
      sm.p[d] = poly_eval(seg[i].c[d], tin, 0);
      sm.v[d] = poly_eval(seg[i].c[d], tin, 1);
      sm.a[d] = poly_eval(seg[i].c[d], tin, 2);
    }
    for (int d = 0; d < 3; ++d) {
      sm.j[d] = poly_eval(seg[i].c[d], tin, 3);
      sm.s[d] = poly_eval(seg[i].c[d], tin, 4);
    }
    // quaternionFromYaw -> AngleAxis about z: w = cos(yaw/2), z = sin(yaw/2), x = y = 0;
    // yawFromQuaternion = atan2(2(wz + xy), 1 - 2(yy + zz))   (eth_mav_msgs/common.h:130-140)
    const double ha = 0.5 * sm.p[3];
    const double qw = m_cos_k(ha), qz = m_sin_k(ha);
    sm.yaw_out = m_atan2_k(2.0 * (qw * qz + 0.0 * 0.0), 1.0 - 2.0 * (0.0 * 0.0 + qz * qz));
    const size_t idx = out->size();
    sm.t_ns = (int64_t)((t_start + dt * (double)idx) * 1.e9);  // trajectory_sampling.cpp:83
    out->push_back(sm);
    tin += dt;
    acc += dt;
  }
  return true;
}

// ---- mrs_lib::geometry cyclic helpers (not vendored; restated from memory) ----------------
static const double kPi = 3.14159265358979323846;
static double wrap_range(double val, double minimum, double supremum) {
  const double range = supremum - minimum;
  if (val >= minimum) {
    if (val < supremum) return val;
    if (val < supremum + range) return val - range;
  } else {
    if (val >= minimum - range) return val + range;
  }
  const double rem = std::fmod(val - minimum, range);
  return rem + minimum + (std::signbit(rem) ? range : 0.0);
}
double rad_wrap(double a) { return wrap_range(a, 0.0, 2.0 * kPi); }
double rad_diff(double a, double b) {
  const double d = a - b;
  if (d < -kPi) return d + 2.0 * kPi;
  if (d >= kPi) return d - 2.0 * kPi;
  return d;
}
double rad_dist(double a, double b) { return std::fabs(rad_diff(a, b)); }
double rad_interp(double a, double b, double c) { return rad_wrap(a + c * rad_diff(b, a)); }
double srad_unwrap(double what, double from) { return from + rad_diff(what, from); }

// eth/vertex.cpp:491-565
std::vector<double> estimate_times_euclidean(const std::vector<Vertex>& v, const Limits& L) {
  std::vector<double> out;
  for (size_t i = 0; i + 1 < v.size(); ++i) {
    const double* s = v[i].val[0];
    const double* e = v[i + 1].val[0];
    const double dx = e[0] - s[0], dy = e[1] - s[1], dz = e[2] - s[2];
    const double incl = m_atan2(dz, std::sqrt(dx * dx + dy * dy));
    const double lim = m_atan2(L.v_v, L.v_h);
    double v_max;
    if (incl > lim || incl < -lim) v_max = std::fabs(L.v_v / m_sin(incl));
    else v_max = std::fabs(L.v_h / m_cos(incl));
    const double distance = std::sqrt(dx * dx + dy * dy + dz * dz);  // Eigen norm(): sqrt of the squared sum
    double t = distance / v_max;
    if (t < 0.01) t = 0.01;
    const double ang = std::fabs(rad_dist(s[3], e[3]));
    double hv = 0, ha = 0;
    if (L.v_hdg < std::numeric_limits<float>::max() && L.a_hdg < std::numeric_limits<float>::max()) {
      if (((ang - ((L.v_hdg * L.v_hdg) / L.a_hdg)) / L.v_hdg) < 0) hv = ang / L.v_hdg;
      else hv = (ang - ((L.v_hdg * L.v_hdg) / L.a_hdg)) / L.v_hdg;
      if (ang > kPi / 4) ha = 2 * (L.v_hdg / L.a_hdg);
    }
    const double fix = 1.5 * (hv + ha);
    if (fix > t) t = fix;
    out.push_back(t);
  }
  return out;
}

static void normalize3(double* a) {  // Eigen normalize(): divides by norm when norm > 0
  const double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  if (n > 0.0) {
    a[0] /= n;
    a[1] /= n;
    a[2] /= n;
  }
}

// eth/vertex.cpp:301-485 (including the unconditional overwrite at line 442 and the unused jerk times)
std::vector<double> estimate_times_baca(const std::vector<Vertex>& v, const Limits& L) {
  std::vector<double> out;
  const size_t nv = v.size();
  for (size_t i = 0; i + 1 < nv; ++i) {
    const double* s = v[i].val[0];
    const double* e = v[i + 1].val[0];
    const double dx = e[0] - s[0], dy = e[1] - s[1], dz = e[2] - s[2];
    double at1 = 0, at2 = 0, c1 = 0, c2 = 0;
    const double distance = std::sqrt(dx * dx + dy * dy + dz * dz);
    const double incl = m_atan2(dz, std::sqrt(dx * dx + dy * dy));
    double v_max, a_max, j_max;
    const double lv = m_atan2(L.v_v, L.v_h), la = m_atan2(L.a_v, L.a_h), lj = m_atan2(L.j_v, L.j_h);
    if (incl > lv || incl < -lv) v_max = std::fabs(L.v_v / m_sin(incl)); else v_max = std::fabs(L.v_h / m_cos(incl));
    if (incl > la || incl < -la) a_max = std::fabs(L.a_v / m_sin(incl)); else a_max = std::fabs(L.a_h / m_cos(incl));
    if (incl > lj || incl < -lj) j_max = std::fabs(L.j_v / m_sin(incl)); else j_max = std::fabs(L.j_h / m_cos(incl));
    if (i >= 1) {
      const double* p = v[i - 1].val[0];
      double v1[3] = {s[0] - p[0], s[1] - p[1], s[2] - p[2]};
      double v2[3] = {dx, dy, dz};
      normalize3(v1);
      normalize3(v2);
      const double dot = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
      const double scalar = dot < 0 ? 0.0 : dot;
      c1 = (1 - scalar);
      at1 = c1 * ((v_max / a_max) + (a_max / j_max));
    }
    if (i == 0) {
      c1 = 1.0;
      at1 = (v_max / a_max) + (a_max / j_max);
    }
    if (i == nv - 2) {
      c2 = 1.0;
      at2 = (v_max / a_max) + (a_max / j_max);
    }
    if (i < nv - 2) {
      const double* q = v[i + 2].val[0];
      double v1[3] = {dx, dy, dz};
      double v2[3] = {q[0] - e[0], q[1] - e[1], q[2] - e[2]};
      normalize3(v1);
      normalize3(v2);
      const double dot = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
      const double scalar = dot < 0 ? 0.0 : dot;
      c2 = (1 - scalar);
      at2 = c2 * ((v_max / a_max) + (a_max / j_max));
    }
    (void)c1;
    (void)c2;
    if (at1 > std::sqrt(2 * distance / a_max)) at1 = std::sqrt(2 * distance / a_max);
    if (at2 > std::sqrt(2 * distance / a_max)) at2 = std::sqrt(2 * distance / a_max);
    const double max_velocity_time = distance / v_max;  // vertex.cpp:442 overrides the branch above it
    double t = max_velocity_time + at1 + at2;
    if (t < 0.01) t = 0.01;
    const double ang = std::fabs(rad_dist(s[3], e[3]));
    double hv = 0, ha = 0;
    if (L.v_hdg < std::numeric_limits<float>::max() && L.a_hdg < std::numeric_limits<float>::max()) {
      if (((ang - (2 * (L.v_hdg * L.v_hdg) / L.a_hdg)) / L.v_hdg) < 0) hv = ang / L.v_hdg;
      else hv = (ang - (2 * (L.v_hdg * L.v_hdg) / L.a_hdg)) / L.v_hdg;
      if (ang > kPi / 4) ha = 2 * (L.v_hdg / L.a_hdg);
    }
    const double fix = 1.5 * (hv + ha);
    if (fix > t) t = fix;
    out.push_back(t);
  }
  return out;
}

}  // namespace orc
