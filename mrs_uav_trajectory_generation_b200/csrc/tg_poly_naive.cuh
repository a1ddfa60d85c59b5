// tg_poly_naive.cuh -- direct per-thread transcription of the Jenkins-Traub iteration with dynamically indexed work
// arrays (per-thread local memory).  Kept as the measured BASELINE variant of the extremum kernels (build with
// -DTG_JT_IMPL=0); the production variant is the warp-scheduled state machine in tg_poly.cuh.  Same arithmetic.
#ifndef TG_POLY_NAIVE_CUH_
#define TG_POLY_NAIVE_CUH_

#include "tg_common.cuh"

namespace tg {

TG_HD double poly_eval(const double* __restrict__ c, double t, int deriv);

constexpr int kJtMax = 2 * TG_N - 4;  // highest degree met on the path: 2(N-1)-3 = 15 (|v|^2 derivative, 2 dims)

// Three-stage Jenkins-Traub iteration (TOMS 493).  Coefficients in DECREASING powers.  The control flow and
// every arithmetic expression follow rpoly_ak1.cpp so that roots agree bit for bit with the oracle (which is
// itself checked bit for bit against the reference file, tests/test_rpoly_ref.py).
struct JenkinsTraubDyn {
  double p[kJtMax + 1], qp[kJtMax + 1], K[kJtMax + 1], qk[kJtMax + 1], svk[kJtMax + 1];
  int N, NN;
  double a, b, c, d, e, f, g, h, a1, a3, a7;
  double szr, szi, lzr, lzi;

  TG_HD static void quad_sd(int nn, double u, double v, const double* pp, double* q, double* ra, double* rb) {
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

  TG_HD int calc_sc(double u, double v) {
    quad_sd(N, u, v, K, qk, &c, &d);
    if (dabs(c) <= (10.0 * TG_DBL_EPSILON * dabs(K[N - 1]))) {
      if (dabs(d) <= (10.0 * TG_DBL_EPSILON * dabs(K[N - 2]))) return 3;
    }
    h = v * b;
    if (dabs(d) >= dabs(c)) {
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

  TG_HD void next_k(int tFlag) {
    if (tFlag == 3) {
      K[1] = K[0] = 0.0;
      for (int i = 2; i < N; i++) K[i] = qk[i - 2];
      return;
    }
    const double temp = ((tFlag == 1) ? b : a);
    if (dabs(a1) > (10.0 * TG_DBL_EPSILON * dabs(temp))) {
      a7 = a7 / a1;
      a3 = a3 / a1;
      K[0] = qp[0];
      K[1] = -(a7 * qp[0]) + qp[1];
      for (int i = 2; i < N; i++) K[i] = -(a7 * qp[i - 1]) + a3 * qk[i - 2] + qp[i];
    } else {
      K[0] = 0.0;
      K[1] = -a7 * qp[0];
      for (int i = 2; i < N; i++) K[i] = -(a7 * qp[i - 1]) + a3 * qk[i - 2];
    }
  }

  TG_HD void newest(int tFlag, double u, double v, double* uu, double* vv) const {
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

  TG_HD static void quad(double qa, double b1, double qc, double* sr, double* si, double* lr, double* li) {
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
    if (dabs(bb) < dabs(qc)) {
      ee = ((qc >= 0) ? qa : -qa);
      ee = -ee + bb * (bb / dabs(qc));
      dd = dsqrt(dabs(ee)) * dsqrt(dabs(qc));
    } else {
      ee = -((qa / bb) * (qc / bb)) + 1.0;
      dd = dsqrt(dabs(ee)) * (dabs(bb));
    }
    if (ee >= 0) {
      dd = ((bb >= 0) ? -dd : dd);
      *lr = (-bb + dd) / qa;
      *sr = ((*lr != 0) ? (qc / (*lr)) / qa : *sr);
    } else {
      *lr = *sr = -(bb / qa);
      *si = dabs(dd / qa);
      *li = -(*si);
    }
  }

  TG_HD int quad_it(double uu, double vv) {
    int j = 0, tFlag, tried = 0, nz = 0;
    double ee, mp, omp = 0, relstp = 0, t, u, ui, v, vi, zm;
    u = uu;
    v = vv;
    do {
      quad(1.0, u, v, &szr, &szi, &lzr, &lzi);
      if (dabs(dabs(szr) - dabs(lzr)) > 0.01 * dabs(lzr)) break;
      quad_sd(NN, u, v, p, qp, &a, &b);
      mp = dabs(-(szr * b) + a) + dabs(szi * b);
      zm = dsqrt(dabs(v));
      ee = 2.0 * dabs(qp[0]);
      t = -(szr * b);
      for (int i = 1; i < N; i++) ee = ee * zm + dabs(qp[i]);
      ee = ee * zm + dabs(a + t);
      ee = (9.0 * ee + 2.0 * dabs(t) - 7.0 * (dabs(a + t) + zm * dabs(b))) * TG_DBL_EPSILON;
      if (mp <= 20.0 * ee) {
        nz = 2;
        break;
      }
      j++;
      if (j > 20) break;
      if (j >= 2) {
        if ((relstp <= 0.01) && (mp >= omp) && (!tried)) {
          relstp = ((relstp < TG_DBL_EPSILON) ? dsqrt(TG_DBL_EPSILON) : dsqrt(relstp));
          u = u - u * relstp;
          v = v + v * relstp;
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
        relstp = dabs((-v + vi) / vi);
        u = ui;
        v = vi;
      }
    } while (vi != 0);
    return nz;
  }

  TG_HD int real_it(int* iflag, double* sss) {
    int j = 0;
    const int nm1 = N - 1;
    double ee, kv, mp, ms, omp = 0, pv, s, t = 0;
    *iflag = 0;
    s = *sss;
    for (;;) {
      qp[0] = pv = p[0];
      for (int i = 1; i < NN; i++) qp[i] = pv = pv * s + p[i];
      mp = dabs(pv);
      ms = dabs(s);
      ee = 0.5 * dabs(qp[0]);
      for (int i = 1; i < NN; i++) ee = ee * ms + dabs(qp[i]);
      if (mp <= 20.0 * TG_DBL_EPSILON * (2.0 * ee - mp)) {
        szr = s;
        szi = 0.0;
        return 1;
      }
      j++;
      if (j > 10) break;
      if (j >= 2) {
        if ((dabs(t) <= 0.001 * dabs(-t + s)) && (mp > omp)) {
          *iflag = 1;
          *sss = s;
          break;
        }
      }
      omp = mp;
      qk[0] = kv = K[0];
      for (int i = 1; i < N; i++) qk[i] = kv = kv * s + K[i];
      if (dabs(kv) > dabs(K[nm1]) * 10.0 * TG_DBL_EPSILON) {
        t = -(pv / kv);
        K[0] = qp[0];
        for (int i = 1; i < N; i++) K[i] = t * qk[i - 1] + qp[i];
      } else {
        K[0] = 0.0;
        for (int i = 1; i < N; i++) K[i] = qk[i - 1];
      }
      kv = K[0];
      for (int i = 1; i < N; i++) kv = kv * s + K[i];
      t = ((dabs(kv) > (dabs(K[nm1]) * 10.0 * TG_DBL_EPSILON)) ? -(pv / kv) : 0.0);
      s = s + t;
    }
    return 0;
  }

  TG_HD int fixed_shift(int L2, double sr, double bnd) {
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
        tv = ((vv != 0.0) ? dabs((vv - ovv) / vv) : tv);
        ts = ((ss != 0.0) ? dabs((ss - oss) / ss) : ts);
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
            const bool shortcut = first && ((spass) && (!vpass || (tss < tvv)));
            first = false;
            if (!shortcut) {
              nz = quad_it(ui, vi);
              if (nz > 0) return nz;
              vtry = 1;
              betav = betav * 0.25;
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
              betas = betas * 0.25;
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

// op: DECREASING powers, *degree in/out (number of roots written); iteration counters for the flop report
TG_HD_NOINLINE void rpoly_dyn(const double* op, int* degree, double* zeror, double* zeroi, int* shifts) {
  JenkinsTraubDyn jt;
  double pt[kJtMax + 1], temp[kJtMax + 1];
  const double lb2 = 0x1.62e42fefa39efp-1;          // log(2.0)
  const double lo = TG_FLT_MIN / TG_DBL_EPSILON;
  const double cosr = -0x1.1db8f6d6a512ap-4;         // cos(94 deg) as glibc returns it
  const double sinr = 0x1.fec0b7170fff6p-1;          // sin(94 deg)
  if (*degree > kJtMax) {
    *degree = -1;
    return;
  }
  if (op[0] == 0) {
    *degree = 0;
    return;
  }
  int N = *degree;
  double xx = 0x1.6a09e667f3bcdp-1, yy = -xx;        // sqrt(0.5)
  int j = 0;
  while (op[N] == 0) {
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
        JenkinsTraubDyn::quad(jt.p[0], jt.p[1], jt.p[2], &zeror[*degree - 2], &zeroi[*degree - 2], &zeror[*degree - 1],
                           &zeroi[*degree - 1]);
      }
      break;
    }
    double moduli_max = 0.0, moduli_min = TG_FLT_MAX;
    for (int i = 0; i < NN; i++) {
      const double x = dabs(jt.p[i]);
      if (x > moduli_max) moduli_max = x;
      if ((x != 0) && (x < moduli_min)) moduli_min = x;
    }
    double sc = lo / moduli_min;
    if (((sc <= 1.0) && (moduli_max >= 10)) || ((sc > 1.0) && (TG_FLT_MAX / sc >= moduli_max))) {
      sc = ((sc == 0) ? TG_FLT_MIN : sc);
      const int l = (int)(tgdm::dlog_k(sc) / lb2 + 0.5);
      const double factor = tgdm::scalb(1.0, l);
      if (factor != 1.0)
        for (int i = 0; i < NN; i++) jt.p[i] = jt.p[i] * factor;
    }
    for (int i = 0; i < NN; i++) pt[i] = dabs(jt.p[i]);
    pt[N] = -(pt[N]);
    const int NM1 = N - 1;
    double x = tgdm::dexp_k((tgdm::dlog_k(-pt[N]) - tgdm::dlog_k(pt[0])) / (double)N);
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
    while (dabs(dx / x) > 0.005) {
      df = ff = pt[0];
      for (int i = 1; i < N; i++) {
        ff = x * ff + pt[i];
        df = x * df + ff;
      }
      ff = x * ff + pt[N];
      dx = ff / df;
      x = x - dx;
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
        zerok = ((dabs(jt.K[NM1]) <= dabs(bb) * TG_DBL_EPSILON * 10.0) ? 1 : 0);
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
      if (shifts) ++*shifts;
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

// findRootsJenkinsTraub (rpoly_ak1.cpp:76-120): trim trailing |c| < DBL_MIN, reverse, solve.  Returns #roots.
TG_HD int find_roots_jt_dyn(const double* ci, int n, double* re, double* im, int* shifts) {
  int last = -1;
  for (int i = n - 1; i != -1; i--)
    if (dabs(ci[i]) >= TG_DBL_MIN) {
      last = i;
      break;
    }
  if (last < 1) return 0;  // all zero, or a constant: no roots
  double dec[kJtMax + 1];
  for (int i = 0; i <= last; ++i) dec[i] = ci[last - i];
  int degree = last;
  rpoly_dyn(dec, &degree, re, im, shifts);
  return degree > 0 ? degree : 0;
}

// Quantity q in 0..8 : (group, derivative) = (horizontal|vertical|heading, velocity|acceleration|jerk)
// in the order the reference asks for them (eth/trajectory.cpp:616-622).  coef: [4][10] of one segment.
TG_HD_NOINLINE double segment_max_magnitude_dyn(const double* __restrict__ coef, double T, int q, int* shifts) {
  const int group = q / 3, deriv = q - 3 * group + 1;
  double re[kJtMax], im[kJtMax];
  int nroots;
  if (group == 0) {
    // sum over dims {0,1} of conv(p^(k)[0:n_d], p^(k+1)[0:n_dd]) (eth/segment.cpp:122-139, polynomial.cpp:176-192)
    const int n_d = TG_N - deriv, n_dd = n_d - 1, len = n_d + n_dd - 1;
    double acc[2 * TG_N];
    for (int i = 0; i < len; ++i) acc[i] = 0.0;
    for (int dim = 0; dim < 2; ++dim) {
      const double* c = coef + dim * TG_N;
      double dc[TG_N], ddc[TG_N];
      for (int jx = 0; jx < n_d; ++jx) dc[jx] = c[jx + deriv] * bcoef(deriv, jx + deriv);
      for (int jx = 0; jx < n_dd; ++jx) ddc[jx] = c[jx + deriv + 1] * bcoef(deriv + 1, jx + deriv + 1);
      for (int i = 0; i < len; ++i) {
        double cv = 0.0;
        const int data_idx = i - n_dd + 1;
        const int lower = imax(0, -data_idx), upper = imin(n_dd, n_d - data_idx);
        for (int kidx = lower; kidx < upper; ++kidx) cv = cv + ddc[n_dd - 1 - kidx] * dc[data_idx + kidx];
        acc[i] = acc[i] + cv;
      }
    }
    nroots = find_roots_jt_dyn(acc, len, re, im, shifts);
  } else {
    // single dimension: roots of the (k+1)-th derivative, an N-vector with trailing zeros (polynomial.cpp:69-85)
    const double* c = coef + (group == 1 ? 2 : 3) * TG_N;
    double ddc[TG_N];
    for (int jx = 0; jx < TG_N; ++jx) ddc[jx] = 0.0;
    for (int jx = 0; jx < TG_N - deriv - 1; ++jx) ddc[jx] = c[jx + deriv + 1] * bcoef(deriv + 1, jx + deriv + 1);
    nroots = find_roots_jt_dyn(ddc, TG_N, re, im, shifts);
  }
  // candidates: t_start, t_end, then the real roots inside [0, T] in root order (polynomial.cpp:36-63)
  double best = TG_DBL_LOWEST;
  const int d0 = (group == 0) ? 0 : (group == 1 ? 2 : 3), nd = (group == 0) ? 2 : 1;
  if (0.0 > T) return best;
  for (int ci = 0; ci < nroots + 2; ++ci) {
    double t;
    if (ci == 0) t = 0.0;
    else if (ci == 1) t = T;
    else {
      if (dabs(im[ci - 2]) > TG_DBL_EPSILON) continue;
      t = re[ci - 2];
      if (t < 0.0 || t > T) continue;
    }
    double mag = 0.0;
    for (int dim = d0; dim < d0 + nd; ++dim) {
      const double v = poly_eval(coef + dim * TG_N, t, deriv);
      mag = mag + v * v;
    }
    mag = dsqrt(mag);
    if (best < mag) best = mag;
  }
  return best;
}

}  // namespace tg

#endif  // TG_POLY_NAIVE_CUH_
