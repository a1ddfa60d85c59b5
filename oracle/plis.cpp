// oracle/plis.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// NLopt's LD_LBFGS, i.e. Ladislav Luksan's PLIS (limited-memory BFGS with the Strang recurrence, simple bounds handled by
// an active set, and the PS1L01 safeguarded extrapolation / interpolation line search), as NLopt >= 2.4.2 ships it in
// luksan/plis.c, luksan/pssubs.c and luksan/mssubs.c.  NLopt is a third-party dependency of the reference
// (package.xml:24; call sites nl_impl.h:68-74,178-191) and is NOT vendored under /root/reference, and it is absent from
// this image, so this file RESTATES the published algorithm; it cannot be diffed against the source here.
// PARITY UNPINNED against a real NLopt build (said in DESIGN.md as well); what is pinned is GPU == this file.
//
// What is restated, routine by routine (names are NLopt's):
//   luksan_plis      driver: memory size mf, ix codes from the bounds, result-code mapping
//   plis_            main iteration: PYTRCG / PYFUT1 termination, PYRMC0 constraint release, Strang direction
//                    (MXDRCB, MXDRCF, MXDRSU), descent test, PYTRCS, PS1L01 line search, PYTRCD, PYADC0
//   luksan_ps1l01__  line search: unit first step (INITS = 2 with no estimate of the minimum), sufficient decrease
//                    TOLS = 1e-4, curvature TOLP = 0.8, at most MRED = 10 net reductions / extrapolations
//   luksan_pnint1__  MES = 4 cubic extra-/interpolation with the bisection fall-back
//   nlopt_stop_dx / relstop   NLopt's own x-tolerance test added after every accepted step (util/stop.c)
// Quirks kept on purpose: maxeval is tested only between iterations (a line search may overrun it); FTOL/XTOL need two
// consecutive successes (MTESF = MTESX = 2) and the very first test compares against f + min(sqrt|f|, |f|/10); every
// trial point evaluates the objective AND the gradient.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "oracle.h"

namespace orc {

namespace {

inline double max2(double a, double b) { return a >= b ? a : b; }
inline double min2(double a, double b) { return a <= b ? a : b; }

// ---- mssubs.c: masked vector kernels (job = kbf: > 0 -> only indices with ix >= 0 take part) -----------------------
inline bool in_play(const int* ix, int i, int job) {
  if (job == 0) return true;
  if (job > 0) return ix[i] >= 0;
  return ix[i] != -5;
}
double mxudot(int n, const double* x, const double* y, const int* ix, int job) {
  double t = 0.0;
  for (int i = 0; i < n; ++i)
    if (in_play(ix, i, job)) t += x[i] * y[i];
  return t;
}
void mxuneg(int n, const double* x, double* y, const int* ix, int job) {
  for (int i = 0; i < n; ++i) y[i] = in_play(ix, i, job) ? -x[i] : 0.0;
}
void mxudir(int n, double a, const double* x, const double* y, double* z, const int* ix, int job) {
  for (int i = 0; i < n; ++i)
    if (in_play(ix, i, job)) z[i] = y[i] + a * x[i];
}
// MXDRCB: backward part of the Strang recurrence over m pairs (a = steps, b = gradient differences), newest first
void mxdrcb(int n, int m, const double* a, const double* b, const double* u, double* v, double* x, const int* ix, int job) {
  for (int i = 0; i < m; ++i) {
    v[i] = u[i] * mxudot(n, x, a + (size_t)i * n, ix, job);
    mxudir(n, -v[i], b + (size_t)i * n, x, x, ix, job);
  }
}
// MXDRCF: forward part, oldest first
void mxdrcf(int n, int m, const double* a, const double* b, const double* u, const double* v, double* x, const int* ix, int job) {
  for (int i = m - 1; i >= 0; --i) {
    const double t = u[i] * mxudot(n, x, b + (size_t)i * n, ix, job);
    mxudir(n, v[i] - t, a + (size_t)i * n, x, x, ix, job);
  }
}
// MXDRSU: shift the m newest pairs one slot towards "older"
void mxdrsu(int n, int m, double* a, double* b, double* u) {
  for (int l = m - 1; l >= 0; --l) {
    for (int i = 0; i < n; ++i) {
      a[(size_t)(l + 1) * n + i] = a[(size_t)l * n + i];
      b[(size_t)(l + 1) * n + i] = b[(size_t)l * n + i];
    }
    u[l + 1] = u[l];
  }
}

// ---- pssubs.c -------------------------------------------------------------------------------------------------------
// PCBS04: snap to a bound that is closer than eps9 (relative)
void pcbs04(int nf, double* x, const int* ix, const double* xl, const double* xu, double eps9, int kbf) {
  if (kbf <= 0) return;
  for (int i = 0; i < nf; ++i) {
    const int ixi = std::abs(ix[i]);
    if ((ixi == 1 || ixi == 3 || ixi == 4) && x[i] <= xl[i] + eps9 * max2(std::fabs(xl[i]), 1.0)) x[i] = xl[i];
    if ((ixi == 2 || ixi == 3 || ixi == 4) && x[i] >= xu[i] - eps9 * max2(std::fabs(xu[i]), 1.0)) x[i] = xu[i];
  }
}
// PYADC0: variables sitting on a bound become active (negative ix)
void pyadc0(int nf, int* n, double* x, int* ix, const double* xl, const double* xu, int* inew) {
  *n = nf;
  *inew = 0;
  for (int i = 0; i < nf; ++i) {
    const int ii = ix[i];
    const int ixi = std::abs(ii);
    if (ixi >= 5) {
      ix[i] = -ixi;
    } else if ((ixi == 1 || ixi == 3 || ixi == 4) && x[i] <= xl[i]) {
      x[i] = xl[i];
      ix[i] = (ixi == 4) ? -3 : -ixi;
      --*n;
      if (ii > 0) ++*inew;
    } else if ((ixi == 2 || ixi == 3 || ixi == 4) && x[i] >= xu[i]) {
      x[i] = xu[i];
      ix[i] = (ixi == 3) ? -4 : -ixi;
      --*n;
      if (ii > 0) ++*inew;
    }
  }
}
// PYTRCG: gmax = largest free gradient component, umax = largest multiplier of the wrong sign among active bounds
void pytrcg(int nf, const int* ix, const double* g, double* umax, double* gmax, int kbf, int* iold) {
  if (kbf > 0) {
    *gmax = 0.0;
    *umax = 0.0;
    *iold = 0;
    for (int i = 0; i < nf; ++i) {
      const double t = g[i];
      if (ix[i] >= 0) {
        *gmax = max2(*gmax, std::fabs(t));
      } else if (ix[i] <= -5) {
      } else if ((ix[i] == -1 || ix[i] == -3) && *umax + t >= 0.0) {
      } else if ((ix[i] == -2 || ix[i] == -4) && *umax - t >= 0.0) {
      } else {
        *iold = i + 1;
        *umax = std::fabs(t);
      }
    }
  } else {
    *umax = 0.0;
    *gmax = 0.0;
    for (int i = 0; i < nf; ++i) *gmax = max2(*gmax, std::fabs(g[i]));
  }
}
// PYRMC0: release the active bounds whose multiplier has the wrong sign once umax > eps8 * gmax
void pyrmc0(int nf, int n, int* ix, const double* g, double eps8, double umax, double gmax, double rmax, int* iold, int* irest) {
  if (n == 0 || rmax > 0.0) {
    if (umax > eps8 * gmax) {
      *iold = 0;
      for (int i = 0; i < nf; ++i) {
        const int ixi = ix[i];
        if (ixi >= 0) {
        } else if (ixi <= -5) {
        } else if ((ixi == -1 || ixi == -3) && -g[i] <= 0.0) {
        } else if ((ixi == -2 || ixi == -4) && g[i] <= 0.0) {
        } else {
          ++*iold;
          ix[i] = std::min(std::abs(ix[i]), 3);
          if (rmax == 0.0) break;
        }
      }
      if (*iold > 1) *irest = std::max(*irest, 1);
    }
  }
}

struct Pyfut1 {
  int ntesx = 0, mtesx = 2, ntesf = 0, mtesf = 2, ites = 1, ires1 = 999, ires2 = 0;
};
// PYFUT1: termination tests at the top of every iteration; increments nit when the run goes on
void pyfut1(int n, double f, double* fo, double umax, double gmax, double dmax, double tolx, double tolf, double tolb, double tolg, int kd,
            int* nit, int kit, int mit, int nfv, int mfv, int nfg, int mfg, Pyfut1& c, int* irest, int iters, int* iterm) {
  if (*iterm < 0) return;
  if (c.ites > 0 && iters != 0) {
    if (*nit <= 0) *fo = f + min2(std::sqrt(std::fabs(f)), std::fabs(f) / 10.0);
    if (f <= tolb) {
      *iterm = 3;
      return;
    }
    if (kd > 0) {
      if (gmax <= tolg && umax <= tolg) {
        *iterm = 4;
        return;
      }
    }
    if (*nit <= 0) {
      c.ntesx = 0;
      c.ntesf = 0;
    }
    if (dmax <= tolx) {
      *iterm = 1;
      ++c.ntesx;
      if (c.ntesx >= c.mtesx) return;
    } else {
      c.ntesx = 0;
    }
    const double temp = std::fabs(*fo - f) / max2(std::fabs(f), 1.0);
    if (temp <= tolf) {
      *iterm = 2;
      ++c.ntesf;
      if (c.ntesf >= c.mtesf) return;
    } else {
      c.ntesf = 0;
    }
  }
  if (*nit >= mit) {
    *iterm = 11;
    return;
  }
  if (nfv >= mfv) {
    *iterm = 12;
    return;
  }
  if (nfg >= mfg) {
    *iterm = 13;
    return;
  }
  *iterm = 0;
  if (n > 0 && *nit - kit >= c.ires1 * n + c.ires2) *irest = std::max(*irest, 1);
  ++*nit;
}

// PYTRCS: remember the starting point of the line search and clip rmax so that no bound is crossed
void pytrcs(int nf, const double* x, const int* ix, double* xo, const double* xl, const double* xu, const double* g, double* go,
            const double* s, double* ro, double* fp, double* fo, double f, double* po, double p, double* rmax, int kbf) {
  *fp = *fo;
  *ro = 0.0;
  *fo = f;
  *po = p;
  for (int i = 0; i < nf; ++i) {
    xo[i] = x[i];
    go[i] = g[i];
  }
  if (kbf > 0) {
    for (int i = 0; i < nf; ++i) {
      if (s[i] < 0.0) {
        if (ix[i] == 1 || ix[i] >= 3) *rmax = min2(*rmax, (xl[i] - x[i]) / s[i]);
      } else if (s[i] > 0.0) {
        if (ix[i] == 2 || ix[i] >= 3) *rmax = min2(*rmax, (xu[i] - x[i]) / s[i]);
      }
    }
  }
}

// PYTRCD: after the line search xo, go become the step and the gradient difference; dmax = relative step size
void pytrcd(int nf, double* x, const int* ix, double* xo, double* g, double* go, double r, double* f, double fo, double* p, double* po,
            double* dmax, int kbf, int kd, int* ld, int iters) {
  if (iters > 0) {
    for (int i = 0; i < nf; ++i) xo[i] = x[i] - xo[i];
    for (int i = 0; i < nf; ++i) go[i] = g[i] - go[i];
    *po = r * *po;
    *p = r * *p;
  } else {
    *f = fo;
    *p = *po;
    for (int i = 0; i < nf; ++i) {  // MXVSAV: xo := x - xo, x := old xo
      const double t = xo[i];
      xo[i] = x[i] - xo[i];
      x[i] = t;
    }
    for (int i = 0; i < nf; ++i) {
      const double t = go[i];
      go[i] = g[i] - go[i];
      g[i] = t;
    }
    *ld = kd;
  }
  *dmax = 0.0;
  for (int i = 0; i < nf; ++i) {
    if (kbf > 0 && ix[i] < 0) {
      xo[i] = 0.0;
      go[i] = 0.0;
      continue;
    }
    *dmax = max2(*dmax, std::fabs(xo[i]) / max2(std::fabs(x[i]), 1.0));
  }
}

// PNINT1: new trial step by extrapolation (mode 1) or interpolation (mode 2)
void pnint1(double rl, double ru, double fl, double fu, double pl, double pu, double* r, int mode, int mtyp, int* merr) {
  *merr = 0;
  if (mode <= 0) return;
  if (pl >= 0.0) {
    *merr = 2;
    return;
  } else if (ru <= rl) {
    *merr = 3;
    return;
  }
  double a = 0.0, b = 0.0;
  for (int ntyp = mtyp; ntyp >= 1; --ntyp) {
    double den = 0.0;
    if (ntyp == 1) {  // bisection / fixed extrapolation
      *r = (mode == 1) ? 4.0 * ru : 0.5 * (rl + ru);
      return;
    } else if (ntyp == mtyp) {
      a = (fu - fl) / (pl * (ru - rl));
      b = pu / pl;
    }
    if (ntyp == 2) {  // quadratic with one directional derivative
      den = 2.0 * (1.0 - a);
    } else if (ntyp == 3) {  // quadratic with two directional derivatives
      den = 1.0 - b;
    } else if (ntyp == 4) {  // cubic
      const double c = b - 2.0 * a + 1.0;
      const double d = b - 3.0 * a + 2.0;
      const double dis = d * d - 3.0 * c;
      if (dis < 0.0) continue;
      den = d + std::sqrt(dis);
    } else if (ntyp == 5) {  // conic
      const double dis = a * a - b;
      if (dis < 0.0) continue;
      den = a + std::sqrt(dis);
      if (den <= 0.0) continue;
      const double q = 1.0 / den;
      den = 1.0 - b * (q * (q * q));
    }
    if (mode == 1 && den > 0.0 && den < 1.0) {  // extrapolation accepted
      *r = rl + (ru - rl) / den;
      *r = max2(*r, 1.1 * ru);
      *r = min2(*r, 1e3 * ru);
      return;
    } else if (mode == 2 && den > 1.0) {  // interpolation accepted
      *r = rl + (ru - rl) / den;
      if (rl == 0.0) *r = max2(*r, rl + (ru - rl) * 0.01);
      else *r = max2(*r, rl + (ru - rl) * 0.1);
      *r = min2(*r, rl + (ru - rl) * 0.9);
      return;
    }
  }
}

struct Ps1l01State {
  double fl, fu, pl, pu, rl, ru;
  int mes1, mes2, mes3, mode, mtyp;
};

// PS1L01 with reverse communication: isys == 0 on entry starts a line search, isys == 1 on return asks for f, p at r.
void ps1l01(double* r, double* rp, double f, double fo, double* fp, double p, double po, double* pp, double minf, double maxf, double rmin,
            double rmax, double tols, double tolp, double* par1, double* par2, int* kd, int* ld, int nit, int kit, int* nred, int mred,
            int* maxst, int iest, int inits, int* iters, int kters, int mes, int* isys, Ps1l01State& st) {
  bool l1, l2, l3, l5, l7, m1, m2, m3;
  int merr;
  if (*isys != 1) {
    st.mes1 = 2;
    st.mes2 = 2;
    st.mes3 = 2;
    *iters = 0;
    if (po >= 0.0) {
      *r = 0.0;
      *iters = -2;
      *isys = 0;
      return;
    }
    if (rmax <= 0.0) {
      *iters = 0;
      *isys = 0;
      return;
    }
    // initial step size
    double rtemp;
    if (inits > 0) rtemp = minf - f;
    else if (iest == 0) rtemp = f - *fp;
    else rtemp = max2(f - *fp, minf - f);
    const int init1 = std::abs(inits);
    *rp = 0.0;
    *fp = fo;
    *pp = po;
    if (init1 == 0) {
    } else if (init1 == 1 || (inits >= 1 && iest == 0)) {
      *r = 1.0;
    } else if (init1 == 2) {
      *r = min2(1.0, 4.0 * rtemp / po);
    } else if (init1 == 3) {
      *r = min2(1.0, 2.0 * rtemp / po);
    } else if (init1 == 4) {
      *r = 2.0 * rtemp / po;
    }
    *r = max2(*r, rmin);
    *r = min2(*r, rmax);
    st.mode = 0;
    st.ru = 0.0;
    st.fu = fo;
    st.pu = po;
  } else {
    if (st.mode == 0) {
      *par1 = p / po;
      *par2 = f - fo;
    }
    if (*iters != 0) {
      *isys = 0;
      return;
    }
    if (f <= minf) {
      *iters = 7;
      *isys = 0;
      return;
    }
    l1 = *r <= rmin && nit != kit;
    l2 = *r >= rmax;
    l3 = f - fo <= tols * *r * po;
    l5 = p >= tolp * po || (st.mes2 == 2 && st.mode == 2);
    l7 = st.mes2 <= 2 || st.mode != 0;
    m1 = false;
    m2 = false;
    m3 = l3;
    if (st.mes3 >= 1) {
      m1 = std::fabs(p) <= 0.01 * std::fabs(po) && fo - f >= 9.9999999999999994e-12 * std::fabs(fo);
      l3 = l3 || m1;
    }
    if (st.mes3 >= 2) {
      m2 = std::fabs(p) <= 0.5 * std::fabs(po) && std::fabs(fo - f) <= 2.0000000000000001e-13 * std::fabs(fo);
      l3 = l3 || m2;
    }
    *maxst = l2 ? 1 : 0;
    // termination tests
    if (l1 && !l3) {
      *iters = 0;
      *isys = 0;
      return;
    } else if (l2 && l3 && !l5) {
      *iters = 7;
      *isys = 0;
      return;
    } else if (m3 && st.mes1 == 3) {
      *iters = 5;
      *isys = 0;
      return;
    } else if (l3 && l5 && l7) {
      *iters = 4;
      *isys = 0;
      return;
    } else if (kters < 0 || (kters == 6 && l7)) {
      *iters = 6;
      *isys = 0;
      return;
    } else if (std::abs(*nred) >= mred) {
      *iters = -1;
      *isys = 0;
      return;
    } else {
      *rp = *r;
      *fp = f;
      *pp = p;
      st.mode = std::max(st.mode, 1);
      st.mtyp = std::abs(mes);
      if (f >= maxf) st.mtyp = 1;
    }
    if (st.mode == 1) {  // interval change after extrapolation
      st.rl = st.ru;
      st.fl = st.fu;
      st.pl = st.pu;
      st.ru = *r;
      st.fu = f;
      st.pu = p;
      if (!l3) {
        *nred = 0;
        st.mode = 2;
      } else if (st.mes1 == 1) {
        st.mtyp = 1;
      }
    } else {  // interval change after interpolation
      if (!l3) {
        st.ru = *r;
        st.fu = f;
        st.pu = p;
      } else {
        st.rl = *r;
        st.fl = f;
        st.pl = p;
      }
    }
  }
  // new step size (extrapolation or interpolation)
  pnint1(st.rl, st.ru, st.fl, st.fu, st.pl, st.pu, r, st.mode, st.mtyp, &merr);
  if (merr > 0) {
    *iters = -merr;
    *isys = 0;
    return;
  } else if (st.mode == 1) {
    --*nred;
    *r = min2(*r, rmax);
  } else if (st.mode == 2) {
    ++*nred;
  }
  *kd = 1;
  *ld = -1;
  *isys = 1;
}

// NLopt util/stop.c
bool relstop(double vold, double vnew, double reltol, double abstol) {
  if (std::isinf(vold)) return false;
  return (std::fabs(vnew - vold) < abstol || std::fabs(vnew - vold) < reltol * (std::fabs(vnew) + std::fabs(vold)) * 0.5 ||
          (reltol > 0 && vnew == vold));
}
bool nlopt_stop_dx(int n, const double* x, const double* dx, double xtol_rel, double xtol_abs) {
  for (int i = 0; i < n; ++i)
    if (!relstop(x[i] - dx[i], x[i], xtol_rel, xtol_abs)) return false;
  return true;
}

}  // namespace

// luksan_plis + plis_ (luksan/plis.c).  Returns the NLopt result code; *minf is the objective at the point x returned.
int luksan_plis(int nf, PlisObjective objgrad, void* data, const double* lb, const double* ub, double* x, double* minf, PlisStop* stop) {
  // --- luksan_plis(): memory size and bound codes
  const int kMemAvail = 1310720;
  int mf = std::max(kMemAvail / nf, 10);
  if (stop->maxeval && stop->maxeval <= mf) mf = std::max(stop->maxeval, 1);
  std::vector<int> ixv(nf);
  std::vector<double> xl(lb, lb + nf), xu(ub, ub + nf), gf(nf), s(nf);
  const size_t hist = (size_t)std::max(nf, nf * mf);
  std::vector<double> xo(hist, 0.0), go(hist, 0.0), uo(std::max(nf, mf), 0.0), vo(std::max(nf, mf), 0.0);
  int* ix = ixv.data();
  for (int i = 0; i < nf; ++i) {
    const bool lbu = lb[i] <= -0.99 * HUGE_VAL, ubu = ub[i] >= 0.99 * HUGE_VAL;
    ix[i] = lbu ? (ubu ? 0 : 2) : (ubu ? 1 : (lb[i] == ub[i] ? 5 : 3));
  }
  // --- plis_(): initiation
  const int nb = 1;
  int kbf = nb > 0 ? 2 : 0;
  int nit = 0, nfg = 0, nres = 0;
  int isys = 0;
  Pyfut1 fut;
  const int inits = 2;
  int iterm = 0, iterd = 0, iters = 2;
  const int kters = 3;
  int irest = 0;
  const int mred = 10, mes = 4;
  const double eta9 = 1e120, eps8 = 1.0, eps9 = 1e-8, alf1 = 1e-10, alf2 = 1e10;
  double rmax = eta9, dmax = eta9;
  const double maxf = 1e20;
  const int iest = 0;
  const double minf_est = -HUGE_VAL;
  const double xmax = 1e16;
  double tolx = stop->xtol_rel, tolf = stop->ftol_rel;
  const double tolb = stop->minf_max;
  double tolg = 0.0;
  if (tolx <= 0.0) tolx = 1e-16;
  if (tolf <= 0.0) tolf = 1e-14;
  if (tolg <= 0.0) tolg = 1e-8;
  const double told = 1e-4, tols = 1e-4, tolp = 0.8;
  const int mit = INT_MAX;
  const int mfv = stop->maxeval > 0 ? stop->maxeval : INT_MAX;
  const int mfg = mfv;
  int kd = 1, ld = -1;
  int kit = -(fut.ires1 * nf + fut.ires2);
  double fo = minf_est;
  double f = 0.0, p = 0.0, po = 0.0, pp = 0.0, fp = 0.0, r = 0.0, rp = 0.0, ro = 0.0, rmin = 0.0;
  double umax = 0.0, gmax = 0.0, gnorm = 0.0, snorm = 0.0, par1 = 0.0, par2 = 0.0;
  int n = nf, inew = 0, iold = 0, nred = 0, maxst = 0;
  bool xstop = false;
  Ps1l01State lss{};

  // initial operations with simple bounds
  if (kbf > 0) {
    for (int i = 0; i < nf; ++i) {
      if ((ix[i] == 3 || ix[i] == 4) && xu[i] <= xl[i]) {
        xu[i] = xl[i];
        ix[i] = 5;
      } else if (ix[i] == 5 || ix[i] == 6) {
        xl[i] = x[i];
        xu[i] = x[i];
        ix[i] = 5;
      }
    }
    pcbs04(nf, x, ix, xl.data(), xu.data(), eps9, kbf);
    pyadc0(nf, &n, x, ix, xl.data(), xu.data(), &inew);
  }
  f = objgrad(nf, x, gf.data(), data);
  ++stop->nevals;
  ++nfg;
  static const bool trace = std::getenv("ORC_TRACE") != nullptr;
  if (trace) std::fprintf(stderr, "[plis] start f %.9g n %d\n", f, nf);

  for (;;) {
    // L11120: termination tests
    pytrcg(nf, ix, gf.data(), &umax, &gmax, kbf, &iold);
    pyfut1(nf, f, &fo, umax, gmax, dmax, tolx, tolf, tolb, tolg, kd, &nit, kit, mit, stop->nevals, mfv, nfg, mfg, fut, &irest, iters, &iterm);
    if (iterm != 0) break;
    if (kbf > 0 && rmax > 0.0) pyrmc0(nf, n, ix, gf.data(), eps8, umax, gmax, rmax, &iold, &irest);
    for (;;) {
      // L11130: direction determination
      gnorm = std::sqrt(mxudot(nf, gf.data(), gf.data(), ix, kbf));
      bool steepest = irest != 0;
      int nn = 0;
      if (!steepest) {
        nn = std::min(mf, nit - kit);
        if (nn == 0) steepest = true;
      }
      if (!steepest) {
        // BFGS direction by the Strang formula
        const double b = mxudot(nf, xo.data(), go.data(), ix, kbf);
        if (b <= 0.0) {
          irest = std::max(irest, 1);
          steepest = true;
        } else {
          uo[0] = 1.0 / b;
          mxuneg(nf, gf.data(), s.data(), ix, kbf);
          mxdrcb(nf, nn, xo.data(), go.data(), uo.data(), vo.data(), s.data(), ix, kbf);
          const double a = mxudot(nf, go.data(), go.data(), ix, kbf);
          if (a > 0.0) {
            const double sc = b / a;
            for (int i = 0; i < nf; ++i) s[i] = sc * s[i];
          }
          mxdrcf(nf, nn, xo.data(), go.data(), uo.data(), vo.data(), s.data(), ix, kbf);
          snorm = std::sqrt(mxudot(nf, s.data(), s.data(), ix, kbf));
          const int k = std::min(nn, mf - 1);
          mxdrsu(nf, k, xo.data(), go.data(), uo.data());
          iterd = 1;
        }
      }
      if (steepest) {
        // L12620: restart; a second restart inside one iteration is a failure
        if (kit < nit) {
          ++nres;
          kit = nit;
        } else {
          iterm = -10;
          if (iters < 0) iterm = iters - 5;
        }
        mxuneg(nf, gf.data(), s.data(), ix, kbf);
        snorm = gnorm;
        iterd = 1;
      }
      // test on descent direction and preparation of the line search
      if (kd > 0) p = mxudot(nf, gf.data(), s.data(), ix, kbf);
      if (iterd < 0) {
        iterm = iterd;
      } else {
        if (snorm <= 0.0) irest = std::max(irest, 1);
        else if (p + told * gnorm * snorm <= 0.0) irest = 0;
        else irest = std::max(irest, 1);  // uniform descent criterion
        if (irest == 0) {
          nred = 0;
          rmin = alf1 * gnorm / snorm;
          rmax = min2(alf2 * gnorm / snorm, xmax / snorm);
        }
      }
      if (iterm != 0) break;
      if (irest != 0) continue;
      pytrcs(nf, x, ix, xo.data(), xl.data(), xu.data(), gf.data(), go.data(), s.data(), &ro, &fp, &fo, f, &po, p, &rmax, kbf);
      bool skip_to_bounds = (rmax == 0.0);
      if (!skip_to_bounds) {
        // L11170: line search
        for (;;) {
          ps1l01(&r, &rp, f, fo, &fp, p, po, &pp, minf_est, maxf, rmin, rmax, tols, tolp, &par1, &par2, &kd, &ld, nit, kit, &nred, mred,
                 &maxst, iest, inits, &iters, kters, mes, &isys, lss);
          if (isys == 0) break;
          mxudir(nf, r, s.data(), xo.data(), x, ix, kbf);
          pcbs04(nf, x, ix, xl.data(), xu.data(), eps9, kbf);
          f = objgrad(nf, x, gf.data(), data);
          ++stop->nevals;
          ++nfg;
          p = mxudot(nf, gf.data(), s.data(), ix, kbf);
          if (trace) std::fprintf(stderr, "[plis] nit %d trial r %.6g f %.9g (fo %.9g) p %.4g po %.4g mode %d nred %d\n", nit, r, f, fo, p, po, lss.mode, nred);
        }
        if (iters <= 0) {
          // L11174: line search failed -- back to the start point, restart with steepest descent
          r = 0.0;
          f = fo;
          p = po;
          for (int i = 0; i < nf; ++i) {
            x[i] = xo[i];
            gf[i] = go[i];
          }
          irest = std::max(irest, 1);
          ld = kd;
          continue;
        }
        pytrcd(nf, x, ix, xo.data(), gf.data(), go.data(), r, &f, fo, &p, &po, &dmax, kbf, kd, &ld, iters);
        xstop = nlopt_stop_dx(nf, x, xo.data(), stop->xtol_rel, stop->xtol_abs);
        if (trace) std::fprintf(stderr, "[plis] nit %d accepted r %.6g f %.9g iters %d dmax %.4g xstop %d evals %d\n", nit, r, f, iters, dmax, (int)xstop, stop->nevals);
      }
      // L11175
      if (kbf > 0) {
        for (int i = 0; i < nf; ++i) ix[i] = std::abs(ix[i]);  // MXVINE
        pyadc0(nf, &n, x, ix, xl.data(), xu.data(), &inew);
      }
      if (xstop) iterm = 1;
      break;
    }
    if (iterm != 0) break;
  }
  (void)nres;
  (void)ro;
  (void)rp;
  (void)pp;
  (void)par1;
  (void)par2;
  (void)maxst;
  *minf = f;
  if (trace) std::fprintf(stderr, "[plis] exit iterm %d evals %d f %.9g\n", iterm, stop->nevals, f);
  switch (iterm) {
    case 1: return 4;   // NLOPT_XTOL_REACHED
    case 2: return 3;   // NLOPT_FTOL_REACHED
    case 3: return 2;   // NLOPT_STOPVAL_REACHED (MINF_MAX_REACHED)
    case 4: return 1;   // NLOPT_SUCCESS (gradient tolerance)
    case 6: return 1;
    case 12:
    case 13: return 5;  // NLOPT_MAXEVAL_REACHED
    case 100: return 6;
    case -999: return -5;
    default: return -1;  // NLOPT_FAILURE
  }
}

}  // namespace orc
