// oracle/rpoly.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// Jenkins-Traub real-polynomial root finder (TOMS 493, three-stage variable-shift iteration) restated as a
// small state object.  Follows the arithmetic of src/eth_trajectory_generation/rpoly/rpoly_ak1.cpp:148-932
// expression by expression (same association, same comparisons, same iteration caps) so that in kMathLibm
// mode the roots agree bit-for-bit with the reference file built into oracle/_ref (tests/test_rpoly_ref.py).
#include <cfloat>
#include <cmath>

#include "oracle.h"

namespace orc {
namespace {

constexpr int kMaxDeg = 32;

struct JenkinsTraub {
  // working polynomials (decreasing powers)
  double p[kMaxDeg + 1], qp[kMaxDeg + 1], K[kMaxDeg + 1], qk[kMaxDeg + 1], svk[kMaxDeg + 1];
  int N = 0, NN = 0;
  // scalars shared between the stages (calcSC outputs)
  double a = 0, b = 0, c = 0, d = 0, e = 0, f = 0, g = 0, h = 0, a1 = 0, a3 = 0, a7 = 0;
  // results of an iteration
  double szr = 0, szi = 0, lzr = 0, lzi = 0;

  // rpoly_ak1.cpp:543-559
  static void quad_sd(int nn, double u, double v, const double* pp, double* q, double* ra, double* rb) {
    double bb, aa;
    q[0] = bb = pp[0];
    q[1] = aa = -(bb * u) + pp[1];
    for (int i = 2; i < nn; i++) {
      q[i] = -(aa * u + bb * v) + pp[i];
      bb = aa;
      aa = q[i];
    }
    *ra = aa;
    *rb = bb;
  }

  // rpoly_ak1.cpp:561-602
  int calc_sc(double u, double v) {
    quad_sd(N, u, v, K, qk, &c, &d);
    if (std::fabs(c) <= (10.0 * DBL_EPSILON * std::fabs(K[N - 1]))) {
      if (std::fabs(d) <= (10.0 * DBL_EPSILON * std::fabs(K[N - 2]))) return 3;
    }
    h = v * b;
    if (std::fabs(d) >= std::fabs(c)) {
      e = a / d;
      f = c / d;
      g = u * b;
      a3 = e * (g + a) + h * (b / d);
      a1 = -a + f * b;
      a7 = h + (f + u) * a;
      return 2;
    }
    e = a / c;
    f = d / c;
    g = e * u;
    a3 = e * a + (g + h / c) * b;
    a1 = -(a * (d / c)) + b;
    a7 = g * d + h * f + a;
    return 1;
  }

  // rpoly_ak1.cpp:604-645
  void next_k(int tFlag) {
    if (tFlag == 3) {
      K[1] = K[0] = 0.0;
      for (int i = 2; i < N; i++) K[i] = qk[i - 2];
      return;
    }
    const double temp = ((tFlag == 1) ? b : a);
    if (std::fabs(a1) > (10.0 * DBL_EPSILON * std::fabs(temp))) {
      a7 /= a1;
      a3 /= a1;
      K[0] = qp[0];
      K[1] = -(a7 * qp[0]) + qp[1];
      for (int i = 2; i < N; i++) K[i] = -(a7 * qp[i - 1]) + a3 * qk[i - 2] + qp[i];
    } else {
      K[0] = 0.0;
      K[1] = -a7 * qp[0];
      for (int i = 2; i < N; i++) K[i] = -(a7 * qp[i - 1]) + a3 * qk[i - 2];
    }
  }

  // rpoly_ak1.cpp:647-683
  void newest(int tFlag, double u, double v, double* uu, double* vv) const {
    *vv = *uu = 0.0;
    if (tFlag == 3) return;
    double a4, a5;
    if (tFlag != 2) {
      a4 = a + u * b + h * f;
      a5 = c + (u + v * f) * d;
    } else {
      a4 = (a + g) * f + h;
      a5 = (f + u) * c + v * d;
    }
    const double b1 = -K[N - 1] / p[N];
    const double b2 = -(K[N - 2] + b1 * p[N - 1]) / p[N];
    const double c1 = v * b2 * a1;
    const double c2 = b1 * a7;
    const double c3 = b1 * b1 * a3;
    const double c4 = -(c2 + c3) + c1;
    const double temp = -c4 + a5 + b1 * a4;
    if (temp != 0.0) {
      *uu = -((u * (c3 + c2) + v * (b1 * a1 + b2 * a7)) / temp) + u;
      *vv = v * (1.0 + c4 / temp);
    }
  }

  // rpoly_ak1.cpp:881-932
  static void quad(double qa, double b1, double qc, double* sr, double* si, double* lr, double* li) {
    *sr = *si = *lr = *li = 0.0;
    if (qa == 0) {
      *sr = ((b1 != 0) ? -(qc / b1) : *sr);
      return;
    }
    if (qc == 0) {
      *lr = -(b1 / qa);
      return;
    }
    const double bb = b1 / 2.0;
    double dd, ee;
    if (std::fabs(bb) < std::fabs(qc)) {
      ee = ((qc >= 0) ? qa : -qa);
      ee = -ee + bb * (bb / std::fabs(qc));
      dd = std::sqrt(std::fabs(ee)) * std::sqrt(std::fabs(qc));
    } else {
      ee = -((qa / bb) * (qc / bb)) + 1.0;
      dd = std::sqrt(std::fabs(ee)) * (std::fabs(bb));
    }
    if (ee >= 0) {
      dd = ((bb >= 0) ? -dd : dd);
      *lr = (-bb + dd) / qa;
      *sr = ((*lr != 0) ? (qc / (*lr)) / qa : *sr);
    } else {
      *lr = *sr = -(bb / qa);
      *si = std::fabs(dd / qa);
      *li = -(*si);
    }
  }

  // rpoly_ak1.cpp:685-783 ; returns number of zeros found (0 or 2)
  int quad_it(double uu, double vv) {
    int j = 0, tFlag, tried = 0, nz = 0;
    double ee, mp, omp = 0, relstp = 0, t, u, ui, v, vi, zm;
    u = uu;
    v = vv;
    do {
      quad(1.0, u, v, &szr, &szi, &lzr, &lzi);
      if (std::fabs(std::fabs(szr) - std::fabs(lzr)) > 0.01 * std::fabs(lzr)) break;
      quad_sd(NN, u, v, p, qp, &a, &b);
      mp = std::fabs(-(szr * b) + a) + std::fabs(szi * b);
      zm = std::sqrt(std::fabs(v));
      ee = 2.0 * std::fabs(qp[0]);
      t = -(szr * b);
      for (int i = 1; i < N; i++) ee = ee * zm + std::fabs(qp[i]);
      ee = ee * zm + std::fabs(a + t);
      ee = (9.0 * ee + 2.0 * std::fabs(t) - 7.0 * (std::fabs(a + t) + zm * std::fabs(b))) * DBL_EPSILON;
      if (mp <= 20.0 * ee) {
        nz = 2;
        break;
      }
      j++;
      if (j > 20) break;
      if (j >= 2) {
        if ((relstp <= 0.01) && (mp >= omp) && (!tried)) {
          relstp = ((relstp < DBL_EPSILON) ? std::sqrt(DBL_EPSILON) : std::sqrt(relstp));
          u -= u * relstp;
          v += v * relstp;
          quad_sd(NN, u, v, p, qp, &a, &b);
          for (int i = 0; i < 5; i++) {
            tFlag = calc_sc(u, v);
            next_k(tFlag);
          }
          tried = 1;
          j = 0;
        }
      }
      omp = mp;
      tFlag = calc_sc(u, v);
      next_k(tFlag);
      tFlag = calc_sc(u, v);
      newest(tFlag, u, v, &ui, &vi);
      if (vi != 0) {
        relstp = std::fabs((-v + vi) / vi);
        u = ui;
        v = vi;
      }
    } while (vi != 0);
    return nz;
  }

  // rpoly_ak1.cpp:785-879 ; returns zeros found (0/1), *iflag = 1 when a near-double real zero is suspected
  int real_it(int* iflag, double* sss) {
    int j = 0;
    const int nm1 = N - 1;
    double ee, kv, mp, ms, omp = 0, pv, s, t = 0;
    *iflag = 0;
    s = *sss;
    for (;;) {
      qp[0] = pv = p[0];
      for (int i = 1; i < NN; i++) qp[i] = pv = pv * s + p[i];
      mp = std::fabs(pv);
      ms = std::fabs(s);
      ee = 0.5 * std::fabs(qp[0]);
      for (int i = 1; i < NN; i++) ee = ee * ms + std::fabs(qp[i]);
      if (mp <= 20.0 * DBL_EPSILON * (2.0 * ee - mp)) {
        szr = s;
        szi = 0.0;
        return 1;
      }
      j++;
      if (j > 10) break;
      if (j >= 2) {
        if ((std::fabs(t) <= 0.001 * std::fabs(-t + s)) && (mp > omp)) {
          *iflag = 1;
          *sss = s;
          break;
        }
      }
      omp = mp;
      qk[0] = kv = K[0];
      for (int i = 1; i < N; i++) qk[i] = kv = kv * s + K[i];
      if (std::fabs(kv) > std::fabs(K[nm1]) * 10.0 * DBL_EPSILON) {
        t = -(pv / kv);
        K[0] = qp[0];
        for (int i = 1; i < N; i++) K[i] = t * qk[i - 1] + qp[i];
      } else {
        K[0] = 0.0;
        for (int i = 1; i < N; i++) K[i] = qk[i - 1];
      }
      kv = K[0];
      for (int i = 1; i < N; i++) kv = kv * s + K[i];
      t = ((std::fabs(kv) > (std::fabs(K[nm1]) * 10.0 * DBL_EPSILON)) ? -(pv / kv) : 0.0);
      s += t;
    }
    return 0;
  }

  // rpoly_ak1.cpp:389-541 ; returns number of zeros found
  int fixed_shift(int L2, double sr, double bnd) {
    int nz = 0;
    double betas, betav, oss, ots = 0, otv = 0, ovv, s = 0, ss, ts, tss, tv, tvv, u, ui, v, vi, vv;
    betav = betas = 0.25;
    u = -(2.0 * sr);
    oss = sr;
    ovv = v = bnd;
    quad_sd(NN, u, v, p, qp, &a, &b);
    int tFlag = calc_sc(u, v);
    for (int j = 0; j < L2; j++) {
      next_k(tFlag);
      tFlag = calc_sc(u, v);
      newest(tFlag, u, v, &ui, &vi);
      vv = vi;
      ss = ((K[N - 1] != 0.0) ? -(p[N] / K[N - 1]) : 0.0);
      ts = tv = 1.0;
      if ((j != 0) && (tFlag != 3)) {
        tv = ((vv != 0.0) ? std::fabs((vv - ovv) / vv) : tv);
        ts = ((ss != 0.0) ? std::fabs((ss - oss) / ss) : ts);
        tvv = ((tv < otv) ? tv * otv : 1.0);
        tss = ((ts < ots) ? ts * ots : 1.0);
        const int vpass = ((tvv < betav) ? 1 : 0);
        const int spass = ((tss < betas) ? 1 : 0);
        if ((spass) || (vpass)) {
          for (int i = 0; i < N; i++) svk[i] = K[i];
          s = ss;
          int stry = 0, vtry = 0;
          bool first = true;
          do {
            int iFlag = 1;
            // first pass only: go straight to the linear iteration when the s sequence converges faster
            const bool shortcut = first && ((spass) && (!vpass || (tss < tvv)));
            first = false;
            if (!shortcut) {
              nz = quad_it(ui, vi);
              if (nz > 0) return nz;
              vtry = 1;
              betav *= 0.25;
              if (stry || (!spass)) {
                iFlag = 0;
              } else {
                for (int i = 0; i < N; i++) K[i] = svk[i];
              }
            }
            if (iFlag != 0) {
              nz = real_it(&iFlag, &s);
              if (nz > 0) return nz;
              stry = 1;
              betas *= 0.25;
              if (iFlag != 0) {
                ui = -(s + s);
                vi = s * s;
                continue;
              }
            }
            for (int i = 0; i < N; i++) K[i] = svk[i];
          } while (vpass && !vtry);
          quad_sd(NN, u, v, p, qp, &a, &b);
          tFlag = calc_sc(u, v);
        }
      }
      ovv = vv;
      oss = ss;
      otv = tv;
      ots = ts;
    }
    return nz;
  }
};

// rpoly_ak1.cpp:148-387.  op: decreasing powers, *degree in/out.
void rpoly(const double* op, int* degree, double* zeror, double* zeroi) {
  JenkinsTraub jt;
  double pt[kMaxDeg + 1], temp[kMaxDeg + 1];
  const double lb2 = (math_mode() == kMathDet) ? 0x1.62e42fefa39efp-1 : std::log(2.0);
  const double lo = FLT_MIN / DBL_EPSILON;
  // cos/sin of 94 degrees as glibc returns them for 94.0 * (3.14159265358979323846 / 180)
  const double RADFAC = 3.14159265358979323846 / 180;
  const double cosr = (math_mode() == kMathDet) ? -0x1.1db8f6d6a512ap-4 : std::cos(94.0 * RADFAC);
  const double sinr = (math_mode() == kMathDet) ? 0x1.fec0b7170fff6p-1 : std::sin(94.0 * RADFAC);
  if (*degree > kMaxDeg) {
    *degree = -1;
    return;
  }
  if (op[0] == 0) {
    *degree = 0;
    return;
  }
  int N = *degree;
  double xx = std::sqrt(0.5), yy = -xx;
  int j = 0;
  while (op[N] == 0) {  // zeros at the origin
    zeror[j] = zeroi[j] = 0.0;
    N--;
    j++;
  }
  int NN = N + 1;
  for (int i = 0; i < NN; i++) jt.p[i] = op[i];
  while (N >= 1) {
    if (N <= 2) {
      if (N < 2) {
        zeror[*degree - 1] = -(jt.p[1] / jt.p[0]);
        zeroi[*degree - 1] = 0.0;
      } else {
        JenkinsTraub::quad(jt.p[0], jt.p[1], jt.p[2], &zeror[*degree - 2], &zeroi[*degree - 2], &zeror[*degree - 1],
                           &zeroi[*degree - 1]);
      }
      break;
    }
    double moduli_max = 0.0, moduli_min = FLT_MAX;
    for (int i = 0; i < NN; i++) {
      const double x = std::fabs(jt.p[i]);
      if (x > moduli_max) moduli_max = x;
      if ((x != 0) && (x < moduli_min)) moduli_min = x;
    }
    double sc = lo / moduli_min;
    if (((sc <= 1.0) && (moduli_max >= 10)) || ((sc > 1.0) && (FLT_MAX / sc >= moduli_max))) {
      sc = ((sc == 0) ? FLT_MIN : sc);
      const int l = (int)(m_log_k(sc) / lb2 + 0.5);
      const double factor = std::ldexp(1.0, l);  // pow(2.0, l), exact
      if (factor != 1.0)
        for (int i = 0; i < NN; i++) jt.p[i] *= factor;
    }
    for (int i = 0; i < NN; i++) pt[i] = std::fabs(jt.p[i]);
    pt[N] = -(pt[N]);
    const int NM1 = N - 1;
    double x = m_exp_k((m_log_k(-pt[N]) - m_log_k(pt[0])) / (double)N);
    if (pt[NM1] != 0) {
      const double xm = -pt[N] / pt[NM1];
      x = ((xm < x) ? xm : x);
    }
    double xm = x, ff;
    do {
      x = xm;
      xm = 0.1 * x;
      ff = pt[0];
      for (int i = 1; i < NN; i++) ff = ff * xm + pt[i];
    } while (ff > 0);
    double dx = x, df;
    while (std::fabs(dx / x) > 0.005) {
      df = ff = pt[0];
      for (int i = 1; i < N; i++) {
        ff = x * ff + pt[i];
        df = x * df + ff;
      }
      ff = x * ff + pt[N];
      dx = ff / df;
      x -= dx;
    }
    const double bnd = x;
    for (int i = 1; i < N; i++) jt.K[i] = (double)(N - i) * jt.p[i] / ((double)N);
    jt.K[0] = jt.p[0];
    const double aa = jt.p[N], bb = jt.p[NM1];
    int zerok = ((jt.K[NM1] == 0) ? 1 : 0);
    for (int jj = 0; jj < 5; jj++) {
      const double cc = jt.K[NM1];
      if (zerok) {
        for (int i = 0; i < NM1; i++) {
          const int jx = NM1 - i;
          jt.K[jx] = jt.K[jx - 1];
        }
        jt.K[0] = 0;
        zerok = ((jt.K[NM1] == 0) ? 1 : 0);
      } else {
        const double t = -aa / cc;
        for (int i = 0; i < NM1; i++) {
          const int jx = NM1 - i;
          jt.K[jx] = t * jt.K[jx - 1] + jt.p[jx];
        }
        jt.K[0] = jt.p[0];
        zerok = ((std::fabs(jt.K[NM1]) <= std::fabs(bb) * DBL_EPSILON * 10.0) ? 1 : 0);
      }
    }
    for (int i = 0; i < N; i++) temp[i] = jt.K[i];
    int jj;
    for (jj = 1; jj <= 20; jj++) {
      const double xxx = -(sinr * yy) + cosr * xx;
      yy = sinr * xx + cosr * yy;
      xx = xxx;
      const double sr = bnd * xx;
      jt.N = N;
      jt.NN = NN;
      const int NZ = jt.fixed_shift(20 * jj, sr, bnd);
      if (NZ != 0) {
        j = *degree - N;
        zeror[j] = jt.szr;
        zeroi[j] = jt.szi;
        NN = NN - NZ;
        N = NN - 1;
        for (int i = 0; i < NN; i++) jt.p[i] = jt.qp[i];
        if (NZ != 1) {
          zeror[j + 1] = jt.lzr;
          zeroi[j + 1] = jt.lzi;
        }
        break;
      } else {
        for (int i = 0; i < N; i++) jt.K[i] = temp[i];
      }
    }
    if (jj > 20) {
      *degree -= N;
      break;
    }
  }
}

}  // namespace

// rpoly_ak1.cpp:59-120 : trims trailing |c| < DBL_MIN, reverses, calls rpoly.  Returns number of roots.
int find_roots_jt(const double* ci, int n, double* re, double* im, bool* ok) {
  int last = -1;
  for (int i = n - 1; i != -1; i--)
    if (std::fabs(ci[i]) >= DBL_MIN) {
      last = i;
      break;
    }
  if (last == -1) {
    *ok = true;
    return 0;
  }
  const int ncoef = last + 1;
  if (ncoef < 2) {
    *ok = true;
    return 0;
  }
  double dec[kMaxDeg + 1];
  for (int i = 0; i < ncoef; ++i) dec[i] = ci[last - i];
  int degree = ncoef - 1;
  rpoly(dec, &degree, re, im);
  *ok = degree > 0;
  return degree > 0 ? degree : 0;
}

}  // namespace orc
