// oracle/linear.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Restates include/eth_trajectory_generation/impl/polynomial_optimization_linear_impl.h
// (PolynomialOptimization<10>) without Eigen.  Operation order = DESIGN.md "numeric contract".
#include <cfloat>
#include <cmath>
#include <cstring>

#include "../include/tg_detmath.h"
#include "oracle.h"

namespace orc {

static int g_math_mode = kMathLibm;
void set_math_mode(int mode) { g_math_mode = mode; }
int math_mode() { return g_math_mode; }

double m_log(double x) { return g_math_mode == kMathDet ? tgdm::dlog(x) : std::log(x); }
double m_exp(double x) { return g_math_mode == kMathDet ? tgdm::dexp(x) : std::exp(x); }
double m_sin(double x) { return g_math_mode == kMathDet ? tgdm::dsin(x) : std::sin(x); }
double m_cos(double x) { return g_math_mode == kMathDet ? tgdm::dcos(x) : std::cos(x); }
double m_atan2(double y, double x) { return g_math_mode == kMathDet ? tgdm::datan2(y, x) : std::atan2(y, x); }
double m_log_k(double x) { return g_math_mode == kMathDet ? tgdm::dlog_k(x) : std::log(x); }
double m_exp_k(double x) { return g_math_mode == kMathDet ? tgdm::dexp_k(x) : std::exp(x); }
double m_sin_k(double x) { return g_math_mode == kMathDet ? tgdm::dsin_k(x) : std::sin(x); }
double m_cos_k(double x) { return g_math_mode == kMathDet ? tgdm::dcos_k(x) : std::cos(x); }
double m_atan2_k(double y, double x) { return g_math_mode == kMathDet ? tgdm::datan2_k(y, x) : std::atan2(y, x); }
double m_hypot(double x, double y) { return g_math_mode == kMathDet ? tgdm::dhypot(x, y) : std::hypot(x, y); }
double m_cbrt(double x) { return g_math_mode == kMathDet ? tgdm::dcbrt(x) : std::cbrt(x); }
double m_pow_int(double t, int e) {
  if (g_math_mode == kMathDet) {
    double pw[32];
    tgdm::powers(t, e, pw);
    return pw[e - 1];
  }
  return std::pow(t, (double)e);
}

// eth/polynomial.cpp:155-170 : table of size kMaxConvolutionSize = 22.
namespace {
struct BaseTable {
  double b[22][22];
  BaseTable() {
    const int N = 22;
    for (int k = 0; k < N; ++k)
      for (int i = 0; i < N; ++i) b[k][i] = 0.0;
    for (int i = 0; i < N; ++i) b[0][i] = 1.0;
    const int DEG = N - 1;
    int order = DEG;
    for (int n = 1; n < N; ++n) {
      for (int i = DEG - order; i < N; ++i) b[n][i] = (double)(order - DEG + i) * b[n - 1][i];
      order--;
    }
  }
};
const BaseTable kBase;
}  // namespace
double base_coeff(int k, int i) { return kBase.b[k][i]; }

// lin_impl.h:112-121 with eth/polynomial.h:208-226 (baseCoeffsWithTime)
void setup_mapping_A(double T, double* A) {
  std::memset(A, 0, sizeof(double) * kN * kN);
  for (int k = 0; k < kHalf; ++k) {
    // row k: derivative k at t = 0 -> only the first coefficient survives
    A[k * kN + k] = base_coeff(k, k);
    // row k + 5: derivative k at t = T
    double* row = A + (k + kHalf) * kN;
    row[k] = base_coeff(k, k);
    if (std::fabs(T) < DBL_EPSILON) continue;
    double tp = T;
    for (int j = k + 1; j < kN; ++j) {
      row[j] = base_coeff(k, j) * tp;
      tp = tp * T;
    }
  }
}

// 5x5 general inverse.  The reference calls Eigen's fixed-size .inverse() (lin_impl.h:171), which for
// sizes > 4 is partial-pivot LU + solve(Identity).  Restated as: right-looking LU with first-maximum
// row pivoting, multipliers by division, then per right-hand-side column forward substitution
// (ascending j) and back substitution (ascending j inside a row).
static void inverse5(const double* Din, double* out) {
  const int n = kHalf;
  double lu[kHalf][kHalf];
  int perm[kHalf];
  for (int i = 0; i < n; ++i) {
    perm[i] = i;
    for (int j = 0; j < n; ++j) lu[i][j] = Din[i * n + j];
  }
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(lu[k][k]);
    for (int i = k + 1; i < n; ++i) {
      const double a = std::fabs(lu[i][k]);
      if (a > best) { best = a; piv = i; }
    }
    if (piv != k) {
      for (int j = 0; j < n; ++j) { const double t = lu[k][j]; lu[k][j] = lu[piv][j]; lu[piv][j] = t; }
      const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
    }
    for (int i = k + 1; i < n; ++i) lu[i][k] = lu[i][k] / lu[k][k];
    for (int i = k + 1; i < n; ++i)
      for (int j = k + 1; j < n; ++j) lu[i][j] = lu[i][j] - lu[i][k] * lu[k][j];
  }
  for (int c = 0; c < n; ++c) {
    double y[kHalf];
    for (int i = 0; i < n; ++i) {
      double s = (perm[i] == c) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s = s - lu[i][j] * y[j];
      y[i] = s;
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < n; ++j) s = s - lu[i][j] * y[j];
      y[i] = s / lu[i][i];
    }
    for (int i = 0; i < n; ++i) out[i * n + c] = y[i];
  }
}

// lin_impl.h:147-177 (Schur complement form)
void invert_mapping(const double* A, double* Ainv) {
  std::memset(Ainv, 0, sizeof(double) * kN * kN);
  double a_inv[kHalf], C[kHalf * kHalf], D[kHalf * kHalf], Dinv[kHalf * kHalf];
  for (int k = 0; k < kHalf; ++k) a_inv[k] = 1.0 / A[k * kN + k];  // cwiseInverse of the diagonal
  for (int i = 0; i < kHalf; ++i)
    for (int j = 0; j < kHalf; ++j) {
      C[i * kHalf + j] = A[(i + kHalf) * kN + j];
      D[i * kHalf + j] = A[(i + kHalf) * kN + (j + kHalf)];
    }
  inverse5(D, Dinv);
  for (int k = 0; k < kHalf; ++k) Ainv[k * kN + k] = a_inv[k];
  for (int i = 0; i < kHalf; ++i)
    for (int j = 0; j < kHalf; ++j) {
      // ((-Dinv) * C) * diag(a_inv), each product accumulated over ascending inner index
      double m = (-Dinv[i * kHalf + 0]) * C[0 * kHalf + j];
      for (int k = 1; k < kHalf; ++k) m = m + (-Dinv[i * kHalf + k]) * C[k * kHalf + j];
      Ainv[(i + kHalf) * kN + j] = m * a_inv[j];
      Ainv[(i + kHalf) * kN + (j + kHalf)] = Dinv[i * kHalf + j];
    }
}

// lin_impl.h:605-618
void cost_jacobian_Q(int r, double T, double* Q) {
  std::memset(Q, 0, sizeof(double) * kN * kN);
  double pw[2 * kN];
  const int emax = (kN - 1 - r) * 2 + 1;
  if (g_math_mode == kMathDet) {
    tgdm::powers(T, emax, pw);
  } else {
    for (int e = 1; e <= emax; ++e) pw[e - 1] = std::pow(T, (double)e);
  }
  for (int col = 0; col < kN - r; ++col)
    for (int row = 0; row < kN - r; ++row) {
      const int e = (kN - 1 - r) * 2 + 1 - row - col;
      const double exponent = (double)e;
      Q[(kN - 1 - row) * kN + (kN - 1 - col)] =
          base_coeff(r, kN - 1 - row) * base_coeff(r, kN - 1 - col) * pw[e - 1] * 2.0 / exponent;
    }
}

// lin_impl.h:61-106
bool LinearSolver::setup(const std::vector<Vertex>& vertices, const std::vector<double>& t, int deriv_to_opt) {
  r = deriv_to_opt;
  vtx = vertices;
  S = (int)vertices.size() - 1;
  if (S < 1 || (int)t.size() != S) return false;
  Ainv.assign((size_t)S * kN * kN, 0.0);
  Q.assign((size_t)S * kN * kN, 0.0);
  seg.assign(S, Segment());
  update_times(t);
  // setupConstraintReorderingMatrix (lin_impl.h:183-257): fixed slots sorted by (vertex, derivative)
  // take the first n_fixed columns, free slots sorted the same way follow (lin.h:289-302).
  const int V = S + 1;
  col_of.assign((size_t)V * kHalf, 0);
  n_fixed = n_free = 0;
  for (int v = 0; v < V; ++v)
    for (int k = 0; k < kHalf; ++k) {
      if (vtx[v].has(k)) col_of[v * kHalf + k] = n_fixed++;
      else col_of[v * kHalf + k] = -(n_free++) - 1;
    }
  d_f.assign((size_t)kD * n_fixed, 0.0);
  d_p.assign((size_t)kD * n_free, 0.0);
  for (int v = 0; v < V; ++v)
    for (int k = 0; k < kHalf; ++k)
      if (vtx[v].has(k))
        for (int d = 0; d < kD; ++d) d_f[(size_t)d * n_fixed + col_of[v * kHalf + k]] = vtx[v].val[k][d];
  return true;
}

// lin_impl.h:288-304
void LinearSolver::update_times(const std::vector<double>& t) {
  times = t;
  double A[kN * kN];
  for (int i = 0; i < S; ++i) {
    cost_jacobian_Q(r, times[i], &Q[(size_t)i * kN * kN]);
    setup_mapping_A(times[i], A);
    invert_mapping(A, &Ainv[(size_t)i * kN * kN]);
  }
}

// H_i = (Ainv_i^T * Q_i) * Ainv_i   (lin_impl.h:320), inner sums over ascending k
void LinearSolver::segment_H(int i, double* H) const {
  const double* Ai = &Ainv[(size_t)i * kN * kN];
  const double* Qi = &Q[(size_t)i * kN * kN];
  double W[kN * kN];
  for (int a = 0; a < kN; ++a)
    for (int b = 0; b < kN; ++b) {
      double s = Ai[0 * kN + a] * Qi[0 * kN + b];
      for (int k = 1; k < kN; ++k) s = s + Ai[k * kN + a] * Qi[k * kN + b];
      W[a * kN + b] = s;
    }
  for (int a = 0; a < kN; ++a)
    for (int b = 0; b < kN; ++b) {
      double s = W[a * kN + 0] * Ai[0 * kN + b];
      for (int k = 1; k < kN; ++k) s = s + W[a * kN + k] * Ai[k * kN + b];
      H[a * kN + b] = s;
    }
}

// lin_impl.h:310-334 (dense, for tests only)
void LinearSolver::dense_R(std::vector<double>* R) const {
  const int n = n_fixed + n_free;
  R->assign((size_t)n * n, 0.0);
  auto col = [&](int v, int k) {
    const int c = col_of[v * kHalf + k];
    return c >= 0 ? c : n_fixed + (-c - 1);
  };
  double H[kN * kN];
  for (int i = 0; i < S; ++i) {
    segment_H(i, H);
    for (int a = 0; a < kN; ++a)
      for (int b = 0; b < kN; ++b) {
        const int ra = col(i + a / kHalf, a % kHalf), cb = col(i + b / kHalf, b % kHalf);
        (*R)[(size_t)ra * n + cb] += H[a * kN + b];
      }
  }
}

// lin_impl.h:340-373.  SparseQR(COLAMD) is replaced by an LU factorisation without pivoting of the FULL
// (non-symmetric as formed) banded Rpp -- see DESIGN.md for why not a one-triangle Cholesky.
bool LinearSolver::solve() {
  if (n_free == 0) {  // lin_impl.h:344-349
    segments_from_compact();
    return true;
  }
  const int V = S + 1;
  const int n = n_free;
  // half bandwidth: a free slot of vertex v couples to free slots of v-1..v+1
  std::vector<int> first_free(V + 1, 0);
  for (int v = 0; v < V; ++v) {
    int cnt = 0;
    for (int k = 0; k < kHalf; ++k) cnt += vtx[v].has(k) ? 0 : 1;
    first_free[v + 1] = first_free[v] + cnt;
  }
  int hbw = 0;
  for (int v = 0; v < V; ++v) {
    const int hi = first_free[(v + 2 <= V) ? v + 2 : V] - 1;
    if (first_free[v + 1] > first_free[v]) hbw = std::max(hbw, hi - first_free[v]);
  }
  const int W = 2 * hbw + 1;
  std::vector<double> band((size_t)n * W, 0.0);  // band[i*W + (j - i + hbw)]
  std::vector<double> rhs((size_t)kD * n, 0.0);
  std::vector<double> Hs((size_t)S * kN * kN);
  for (int i = 0; i < S; ++i) segment_H(i, &Hs[(size_t)i * kN * kN]);
  // R entry between slot (v,a) and slot (w,b); |v-w| <= 1.  Contribution of segment v-1 is added first.
  auto Rentry = [&](int v, int a, int w, int b) -> double {
    if (w == v) {
      double s = 0.0;
      bool have = false;
      if (v > 0) { s = Hs[(size_t)(v - 1) * kN * kN + (kHalf + a) * kN + (kHalf + b)]; have = true; }
      if (v < S) {
        const double h = Hs[(size_t)v * kN * kN + a * kN + b];
        s = have ? s + h : h;
      }
      return s;
    }
    if (w == v + 1) return Hs[(size_t)v * kN * kN + a * kN + (kHalf + b)];
    /* w == v - 1 */ return Hs[(size_t)w * kN * kN + (kHalf + a) * kN + b];
  };
  for (int v = 0; v < V; ++v)
    for (int a = 0; a < kHalf; ++a) {
      const int ci = col_of[v * kHalf + a];
      if (ci >= 0) continue;
      const int i = -ci - 1;
      for (int w = std::max(0, v - 1); w <= std::min(S, v + 1); ++w)
        for (int b = 0; b < kHalf; ++b) {
          const int cj = col_of[w * kHalf + b];
          const double rv = Rentry(v, a, w, b);
          if (cj < 0) {
            const int j = -cj - 1;
            band[(size_t)i * W + (j - i + hbw)] = rv;
          } else {
            // rhs = (-Rpf) * d_f accumulated over ascending fixed column (the (w,b) loop order IS ascending)
            for (int d = 0; d < kD; ++d) rhs[(size_t)d * n + i] = rhs[(size_t)d * n + i] + (-rv) * d_f[(size_t)d * n_fixed + cj];
          }
        }
    }
  // forward elimination, no pivoting; multipliers through the reciprocal pivot (as LAPACK dgetf2 scales by 1/pivot)
  std::vector<double> rinv(n);
  for (int k = 0; k < n; ++k) {
    rinv[k] = 1.0 / band[(size_t)k * W + hbw];
    const int iend = std::min(n - 1, k + hbw);
    for (int i = k + 1; i <= iend; ++i) {
      const double l = band[(size_t)i * W + (k - i + hbw)] * rinv[k];
      for (int j = k + 1; j <= iend; ++j)
        band[(size_t)i * W + (j - i + hbw)] = band[(size_t)i * W + (j - i + hbw)] - l * band[(size_t)k * W + (j - k + hbw)];
      for (int d = 0; d < kD; ++d) rhs[(size_t)d * n + i] = rhs[(size_t)d * n + i] - l * rhs[(size_t)d * n + k];
    }
  }
  // back substitution, far columns first (descending j), x_i = s * (1 / a_ii)
  for (int d = 0; d < kD; ++d) {
    double* x = &d_p[(size_t)d * n];
    for (int i = n - 1; i >= 0; --i) {
      double s = rhs[(size_t)d * n + i];
      for (int j = std::min(n - 1, i + hbw); j > i; --j) s = s - band[(size_t)i * W + (j - i + hbw)] * x[j];
      x[i] = s * rinv[i];
    }
  }
  segments_from_compact();
  return true;
}

// lin_impl.h:263-282
void LinearSolver::segments_from_compact() {
  for (int d = 0; d < kD; ++d)
    for (int i = 0; i < S; ++i) {
      double nd[kN];
      for (int a = 0; a < kN; ++a) {
        const int c = col_of[(i + a / kHalf) * kHalf + a % kHalf];
        nd[a] = c >= 0 ? d_f[(size_t)d * n_fixed + c] : d_p[(size_t)d * n_free + (-c - 1)];
      }
      const double* Ai = &Ainv[(size_t)i * kN * kN];
      for (int a = 0; a < kN; ++a) {
        double s = Ai[a * kN + 0] * nd[0];
        for (int k = 1; k < kN; ++k) s = s + Ai[a * kN + k] * nd[k];
        seg[i].c[d][a] = s;
      }
      seg[i].T = times[i];
    }
}

// lin_impl.h:127-141 : 0.5 * sum_i sum_d c^T Q_i c, evaluated as (c^T Q) c
double LinearSolver::cost() const {
  double total = 0;
  for (int i = 0; i < S; ++i) {
    const double* Qi = &Q[(size_t)i * kN * kN];
    for (int d = 0; d < kD; ++d) {
      const double* c = seg[i].c[d];
      double partial = 0;
      for (int b = 0; b < kN; ++b) {
        double s = c[0] * Qi[0 * kN + b];
        for (int k = 1; k < kN; ++k) s = s + c[k] * Qi[k * kN + b];
        partial = (b == 0) ? s * c[b] : partial + s * c[b];
      }
      total += partial;
    }
  }
  return 0.5 * total;
}

}  // namespace orc
