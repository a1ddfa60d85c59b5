// tg_plis.cuh -- segment-time allocation driver: the optimiser behind PolynomialOptimizationNonLinear<10>::
// optimizeTimeMellingerOuterLoop (reference: nl_impl.h:159-234, objective 616-649, gradient 256-333).
//
// The reference hands the objective to nlopt::opt(LD_LBFGS) (nl_impl.h:68-74, 178-191).  NLopt's LD_LBFGS is Ladislav
// Luksan's PLIS: limited-memory BFGS through the Strang recurrence, simple bounds by an active set, the PS1L01 line search
// (unit first step, sufficient decrease 1e-4, curvature 0.8, cubic extra-/interpolation, at most 10 net reductions), FTOL /
// XTOL tests that must hold twice in a row, maxeval tested between iterations only, plus NLopt's own relative-step test
// after every accepted step.  NLopt is a third-party library that the reference does not vendor and that is absent here:
// this is a restatement of the published algorithm (luksan/plis.c, pssubs.c, mssubs.c of NLopt >= 2.4.2), arranged as a
// per-problem STATE MACHINE that is advanced once per objective evaluation.  The expensive part of an evaluation (the
// S+1 linear solves of nl_impl.h:282-323) runs as batched solve launches; this file only consumes the S+1 costs.
// One thread per problem.  PARITY UNPINNED against a real NLopt build; pinned bit for bit by the tests against the CPU
// checker's statement of the same algorithm (straight-line code with a callback, written independently of this file).
#ifndef TG_PLIS_CUH_
#define TG_PLIS_CUH_

#include "tg_common.cuh"

namespace tg {

constexpr int kPlisMfMax = 32;            // history pairs kept on the device; NLopt keeps min(maxeval, 1310720 / n)
constexpr double kTimeLowerBound = 0.01;  // nl.h:32
constexpr double kPlisHuge = 1e120;       // eta9

struct PlisScalars {
  // plis_() locals that live across evaluations
  double f, fo, fp, p, po, pp, r, rp, rmin, rmax, dmax, umax, gmax, gnorm, snorm;
  double f_last;  // cost at the last evaluated point = OptimizationInfo::cost_trajectory (nl_impl.h:646)
  // ps1l01 state
  double fl, fu, pl, pu, rl, ru;
  double uo[kPlisMfMax], vo[kPlisMfMax];
  int mode, mtyp, isys;
  int nit, kit, nred, iters, irest, ntesx, ntesf, nfree, kd, ld, iterm, xstop;
  int mf, n_evals, stage, code, done;
};

struct PlisVectors {  // all of length S for this problem; xo / go hold mf slots (newest first), stride `hstride`
  double *x, *gf, *s, *xeval, *xo, *go;
  int* ix;
  size_t hstride;
};

// f2c's MAX2 / MIN2 (they differ from std::max / std::min when an argument is NaN)
TG_HD double pmax2(double a, double b) { return a >= b ? a : b; }
TG_HD double pmin2(double a, double b) { return a <= b ? a : b; }
TG_HD int iabs(int a) { return a < 0 ? -a : a; }

// gradient of the Mellinger objective from the S+1 costs of one evaluation (nl_impl.h:319-322)
TG_HD double mellinger_grad(int S, const double* __restrict__ costs, int n) {
  if (S == 1) return 0.0;  // nl_impl.h:264-271
  return (costs[1 + n] - costs[0]) / 0.1;
}

// NLopt util/stop.c
TG_HD bool relstop(double vold, double vnew, double reltol, double abstol) {
  if (tgdm::disinf(vold)) return false;
  return (dabs(vnew - vold) < abstol || dabs(vnew - vold) < reltol * (dabs(vnew) + dabs(vold)) * 0.5 || (reltol > 0 && vnew == vold));
}

// ---- mssubs.c (kbf = 2: only indices with ix >= 0 take part) ----------------------------------------------------------
TG_HD double plis_dot(int n, const double* x, const double* y, const int* ix) {
  double t = 0.0;
  for (int i = 0; i < n; ++i)
    if (ix[i] >= 0) t = t + x[i] * y[i];
  return t;
}
TG_HD void plis_dir(int n, double a, const double* x, const double* y, double* z, const int* ix) {
  for (int i = 0; i < n; ++i)
    if (ix[i] >= 0) z[i] = y[i] + a * x[i];
}

// PCBS04
TG_HD void plis_snap(int n, double* x, const int* ix) {
  const double xl = kTimeLowerBound, xu = TG_DBL_MAX, eps9 = 1e-8;
  for (int i = 0; i < n; ++i) {
    const int ixi = iabs(ix[i]);
    if ((ixi == 1 || ixi == 3 || ixi == 4) && x[i] <= xl + eps9 * pmax2(dabs(xl), 1.0)) x[i] = xl;
    if ((ixi == 2 || ixi == 3 || ixi == 4) && x[i] >= xu - eps9 * pmax2(dabs(xu), 1.0)) x[i] = xu;
  }
}
// PYADC0
TG_HD void plis_add_active(int n, int* nfree, double* x, int* ix) {
  const double xl = kTimeLowerBound, xu = TG_DBL_MAX;
  *nfree = n;
  for (int i = 0; i < n; ++i) {
    const int ixi = iabs(ix[i]);
    if (ixi >= 5) {
      ix[i] = -ixi;
    } else if ((ixi == 1 || ixi == 3 || ixi == 4) && x[i] <= xl) {
      x[i] = xl;
      ix[i] = (ixi == 4) ? -3 : -ixi;
      --*nfree;
    } else if ((ixi == 2 || ixi == 3 || ixi == 4) && x[i] >= xu) {
      x[i] = xu;
      ix[i] = (ixi == 3) ? -4 : -ixi;
      --*nfree;
    }
  }
}

// PNINT1 (MES = 4 reaches types 4, 3, 2, 1 in turn)
TG_HD void plis_pnint1(double rl, double ru, double fl, double fu, double pl, double pu, double* r, int mode, int mtyp, int* merr) {
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
    if (ntyp == 1) {
      *r = (mode == 1) ? 4.0 * ru : 0.5 * (rl + ru);
      return;
    } else if (ntyp == mtyp) {
      a = (fu - fl) / (pl * (ru - rl));
      b = pu / pl;
    }
    if (ntyp == 2) {
      den = 2.0 * (1.0 - a);
    } else if (ntyp == 3) {
      den = 1.0 - b;
    } else if (ntyp == 4) {
      const double c = b - 2.0 * a + 1.0;
      const double d = b - 3.0 * a + 2.0;
      const double dis = d * d - 3.0 * c;
      if (dis < 0.0) continue;
      den = d + dsqrt(dis);
    } else if (ntyp == 5) {
      const double dis = a * a - b;
      if (dis < 0.0) continue;
      den = a + dsqrt(dis);
      if (den <= 0.0) continue;
      const double q = 1.0 / den;
      den = 1.0 - b * (q * (q * q));
    }
    if (mode == 1 && den > 0.0 && den < 1.0) {
      *r = rl + (ru - rl) / den;
      *r = pmax2(*r, 1.1 * ru);
      *r = pmin2(*r, 1e3 * ru);
      return;
    } else if (mode == 2 && den > 1.0) {
      *r = rl + (ru - rl) / den;
      if (rl == 0.0) *r = pmax2(*r, rl + (ru - rl) * 0.01);
      else *r = pmax2(*r, rl + (ru - rl) * 0.1);
      *r = pmin2(*r, rl + (ru - rl) * 0.9);
      return;
    }
  }
}

// PS1L01 with reverse communication.  isys == 0 on entry: start a line search; isys == 1 on return: evaluate at r.
TG_HD void plis_ps1l01(PlisScalars& st) {
  const double minf = -1.0 / 0.0, maxf = 1e20, tols = 1e-4, tolp = 0.8;
  const int mred = 10, mes = 4, kters = 3, mes1 = 2, mes2 = 2, mes3 = 2;
  int merr;
  if (st.isys != 1) {
    st.iters = 0;
    if (st.po >= 0.0) {
      st.r = 0.0;
      st.iters = -2;
      st.isys = 0;
      return;
    }
    if (st.rmax <= 0.0) {
      st.iters = 0;
      st.isys = 0;
      return;
    }
    // INITS = 2, IEST = 0: rtemp = minf - f = -inf, r = min(1, 4 rtemp / po) = 1
    const double rtemp = minf - st.f;
    st.rp = 0.0;
    st.fp = st.fo;
    st.pp = st.po;
    st.r = pmin2(1.0, 4.0 * rtemp / st.po);
    st.r = pmax2(st.r, st.rmin);
    st.r = pmin2(st.r, st.rmax);
    st.mode = 0;
    st.ru = 0.0;
    st.fu = st.fo;
    st.pu = st.po;
  } else {
    if (st.iters != 0) {
      st.isys = 0;
      return;
    }
    if (st.f <= minf) {
      st.iters = 7;
      st.isys = 0;
      return;
    }
    const bool l1 = st.r <= st.rmin && st.nit != st.kit;
    const bool l2 = st.r >= st.rmax;
    bool l3 = st.f - st.fo <= tols * st.r * st.po;
    const bool l5 = st.p >= tolp * st.po || (mes2 == 2 && st.mode == 2);
    const bool l7 = mes2 <= 2 || st.mode != 0;
    const bool m3 = l3;
    if (mes3 >= 1) {
      const bool m1 = dabs(st.p) <= 0.01 * dabs(st.po) && st.fo - st.f >= 9.9999999999999994e-12 * dabs(st.fo);
      l3 = l3 || m1;
    }
    if (mes3 >= 2) {
      const bool m2 = dabs(st.p) <= 0.5 * dabs(st.po) && dabs(st.fo - st.f) <= 2.0000000000000001e-13 * dabs(st.fo);
      l3 = l3 || m2;
    }
    if (l1 && !l3) {
      st.iters = 0;
      st.isys = 0;
      return;
    } else if (l2 && l3 && !l5) {
      st.iters = 7;
      st.isys = 0;
      return;
    } else if (m3 && mes1 == 3) {
      st.iters = 5;
      st.isys = 0;
      return;
    } else if (l3 && l5 && l7) {
      st.iters = 4;
      st.isys = 0;
      return;
    } else if (kters < 0 || (kters == 6 && l7)) {
      st.iters = 6;
      st.isys = 0;
      return;
    } else if (iabs(st.nred) >= mred) {
      st.iters = -1;
      st.isys = 0;
      return;
    } else {
      st.rp = st.r;
      st.fp = st.f;
      st.pp = st.p;
      st.mode = imax(st.mode, 1);
      st.mtyp = iabs(mes);
      if (st.f >= maxf) st.mtyp = 1;
    }
    if (st.mode == 1) {
      st.rl = st.ru;
      st.fl = st.fu;
      st.pl = st.pu;
      st.ru = st.r;
      st.fu = st.f;
      st.pu = st.p;
      if (!l3) {
        st.nred = 0;
        st.mode = 2;
      } else if (mes1 == 1) {
        st.mtyp = 1;
      }
    } else {
      if (!l3) {
        st.ru = st.r;
        st.fu = st.f;
        st.pu = st.p;
      } else {
        st.rl = st.r;
        st.fl = st.f;
        st.pl = st.p;
      }
    }
  }
  plis_pnint1(st.rl, st.ru, st.fl, st.fu, st.pl, st.pu, &st.r, st.mode, st.mtyp, &merr);
  if (merr > 0) {
    st.iters = -merr;
    st.isys = 0;
    return;
  } else if (st.mode == 1) {
    --st.nred;
    st.r = pmin2(st.r, st.rmax);
  } else if (st.mode == 2) {
    ++st.nred;
  }
  st.kd = 1;
  st.ld = -1;
  st.isys = 1;
}

TG_HD int plis_result_code(int iterm) {
  switch (iterm) {
    case 1: return 4;   // NLOPT_XTOL_REACHED
    case 2: return 3;   // NLOPT_FTOL_REACHED
    case 3: return 2;   // NLOPT_STOPVAL_REACHED
    case 4: return 1;   // NLOPT_SUCCESS
    case 6: return 1;
    case 12:
    case 13: return 5;  // NLOPT_MAXEVAL_REACHED
    case 100: return 6;
    default: return -1;  // NLOPT_FAILURE (nlopt::opt throws, the reference catches and goes on: nl_impl.h:190-208)
  }
}

// luksan_plis() + the initiation part of plis_(): leaves the first evaluation point in xeval.
// Returns false when the start point violates the bounds (nlopt_optimize_ -> NLOPT_INVALID_ARGS before any evaluation).
TG_HD bool plis_begin(int S, PlisScalars& st, const PlisVectors& v, const double* __restrict__ times0, int max_evals) {
  st.mf = (max_evals > 0) ? imin(imax(max_evals, 1), kPlisMfMax) : kPlisMfMax;
  st.f = st.fp = st.p = st.po = st.pp = st.r = st.rp = st.rmin = st.umax = st.gmax = st.gnorm = st.snorm = 0.0;
  st.fl = st.fu = st.pl = st.pu = st.rl = st.ru = 0.0;
  st.f_last = 0.0;
  st.fo = -1.0 / 0.0;  // minf_est
  st.rmax = kPlisHuge;
  st.dmax = kPlisHuge;
  st.mode = st.mtyp = st.isys = 0;
  st.nit = 0;
  st.kit = -(999 * S + 0);
  st.nred = 0;
  st.iters = 2;
  st.irest = 0;
  st.ntesx = st.ntesf = 0;
  st.kd = 1;
  st.ld = -1;
  st.iterm = 0;
  st.xstop = 0;
  st.n_evals = 0;
  st.stage = 0;
  st.code = -1;
  st.done = 0;
  bool inside = true;
  for (int i = 0; i < S; ++i) {
    v.x[i] = times0[i];
    if (times0[i] < kTimeLowerBound || times0[i] > TG_DBL_MAX) inside = false;
    v.ix[i] = 3;  // both bounds finite (0.01 and DBL_MAX < 0.99 HUGE_VAL)
    v.xeval[i] = times0[i];
  }
  if (!inside) {
    st.done = 1;
    st.code = -1;
    st.f = TG_DBL_MAX;
    st.f_last = 0.0;  // OptimizationInfo::cost_trajectory keeps its initial value
    return false;
  }
  plis_snap(S, v.x, v.ix);
  plis_add_active(S, &st.nfree, v.x, v.ix);
  for (int i = 0; i < S; ++i) v.xeval[i] = v.x[i];
  return true;
}

// Advance after one evaluation at xeval: costs[0] = J(xeval), costs[1..S] = J at the S perturbed points.
TG_HD_NOINLINE void plis_advance(int S, PlisScalars& st, const PlisVectors& v, const double* __restrict__ costs, int max_evals, double f_rel,
                                 double x_rel, double x_abs) {
  if (st.done) return;
  const int nf = S;
  const double told = 1e-4, eps8 = 1.0, alf1 = 1e-10, alf2 = 1e10, xmax = 1e16, tolg = 1e-8;
  const double tolx = (x_rel <= 0.0) ? 1e-16 : x_rel, tolf = (f_rel <= 0.0) ? 1e-14 : f_rel;
  const double tolb = -1.0 / 0.0;  // stopval
  const int mfv = max_evals > 0 ? max_evals : 0x7fffffff;
  int* ix = v.ix;
  st.n_evals += 1;
  st.f = costs[0];
  st.f_last = costs[0];
  for (int i = 0; i < nf; ++i) v.gf[i] = mellinger_grad(S, costs, i);
  // where to resume: 0 = top of the iteration (L11120), 1 = direction (L11130), 2 = line search (L11170), 3 = L11175
  int at;
  if (st.stage == 0) {
    at = 0;
  } else {
    st.p = plis_dot(nf, v.gf, v.s, ix);
    at = 2;
  }
  for (;;) {
    if (at == 0) {
      // PYTRCG
      st.gmax = 0.0;
      st.umax = 0.0;
      for (int i = 0; i < nf; ++i) {
        const double t = v.gf[i];
        if (ix[i] >= 0) {
          st.gmax = pmax2(st.gmax, dabs(t));
        } else if (ix[i] <= -5) {
        } else if ((ix[i] == -1 || ix[i] == -3) && st.umax + t >= 0.0) {
        } else if ((ix[i] == -2 || ix[i] == -4) && st.umax - t >= 0.0) {
        } else {
          st.umax = dabs(t);
        }
      }
      // PYFUT1
      if (st.iterm >= 0) {
        bool decided = false;
        if (st.iters != 0) {
          if (st.nit <= 0) st.fo = st.f + pmin2(dsqrt(dabs(st.f)), dabs(st.f) / 10.0);
          if (st.f <= tolb) {
            st.iterm = 3;
            decided = true;
          }
          if (!decided && st.kd > 0 && st.gmax <= tolg && st.umax <= tolg) {
            st.iterm = 4;
            decided = true;
          }
          if (!decided) {
            if (st.nit <= 0) {
              st.ntesx = 0;
              st.ntesf = 0;
            }
            if (st.dmax <= tolx) {
              st.iterm = 1;
              ++st.ntesx;
              if (st.ntesx >= 2) decided = true;
            } else {
              st.ntesx = 0;
            }
          }
          if (!decided) {
            const double temp = dabs(st.fo - st.f) / pmax2(dabs(st.f), 1.0);
            if (temp <= tolf) {
              st.iterm = 2;
              ++st.ntesf;
              if (st.ntesf >= 2) decided = true;
            } else {
              st.ntesf = 0;
            }
          }
        }
        if (!decided) {
          if (st.n_evals >= mfv) {
            st.iterm = 12;
          } else {
            st.iterm = 0;
            if (nf > 0 && st.nit - st.kit >= 999 * nf) st.irest = imax(st.irest, 1);
            ++st.nit;
          }
        }
      }
      if (st.iterm != 0) break;
      // PYRMC0
      if (st.rmax > 0.0 && st.umax > eps8 * st.gmax) {
        int iold = 0;
        for (int i = 0; i < nf; ++i) {
          const int ixi = ix[i];
          if (ixi >= 0) {
          } else if (ixi <= -5) {
          } else if ((ixi == -1 || ixi == -3) && -v.gf[i] <= 0.0) {
          } else if ((ixi == -2 || ixi == -4) && v.gf[i] <= 0.0) {
          } else {
            ++iold;
            ix[i] = imin(iabs(ix[i]), 3);
          }
        }
        if (iold > 1) st.irest = imax(st.irest, 1);
      }
      at = 1;
    }
    if (at == 1) {
      // direction determination
      st.gnorm = dsqrt(plis_dot(nf, v.gf, v.gf, ix));
      bool steepest = st.irest != 0;
      int nn = 0;
      if (!steepest) {
        nn = imin(st.mf, st.nit - st.kit);
        if (nn == 0) steepest = true;
      }
      if (!steepest) {
        const double b = plis_dot(nf, v.xo, v.go, ix);
        if (b <= 0.0) {
          st.irest = imax(st.irest, 1);
          steepest = true;
        } else {
          st.uo[0] = 1.0 / b;
          for (int i = 0; i < nf; ++i) v.s[i] = (ix[i] >= 0) ? -v.gf[i] : 0.0;
          for (int k = 0; k < nn; ++k) {  // MXDRCB
            st.vo[k] = st.uo[k] * plis_dot(nf, v.s, v.xo + (size_t)k * v.hstride, ix);
            plis_dir(nf, -st.vo[k], v.go + (size_t)k * v.hstride, v.s, v.s, ix);
          }
          const double a = plis_dot(nf, v.go, v.go, ix);
          if (a > 0.0) {
            const double sc = b / a;
            for (int i = 0; i < nf; ++i) v.s[i] = sc * v.s[i];
          }
          for (int k = nn - 1; k >= 0; --k) {  // MXDRCF
            const double t = st.uo[k] * plis_dot(nf, v.s, v.go + (size_t)k * v.hstride, ix);
            plis_dir(nf, st.vo[k] - t, v.xo + (size_t)k * v.hstride, v.s, v.s, ix);
          }
          st.snorm = dsqrt(plis_dot(nf, v.s, v.s, ix));
          const int kk = imin(nn, st.mf - 1);
          for (int l = kk - 1; l >= 0; --l) {  // MXDRSU
            for (int i = 0; i < nf; ++i) {
              v.xo[(size_t)(l + 1) * v.hstride + i] = v.xo[(size_t)l * v.hstride + i];
              v.go[(size_t)(l + 1) * v.hstride + i] = v.go[(size_t)l * v.hstride + i];
            }
            st.uo[l + 1] = st.uo[l];
          }
        }
      }
      if (steepest) {
        if (st.kit < st.nit) {
          st.kit = st.nit;
        } else {
          st.iterm = -10;
          if (st.iters < 0) st.iterm = st.iters - 5;
        }
        for (int i = 0; i < nf; ++i) v.s[i] = (ix[i] >= 0) ? -v.gf[i] : 0.0;
        st.snorm = st.gnorm;
      }
      if (st.kd > 0) st.p = plis_dot(nf, v.gf, v.s, ix);
      if (st.snorm <= 0.0) st.irest = imax(st.irest, 1);
      else if (st.p + told * st.gnorm * st.snorm <= 0.0) st.irest = 0;
      else st.irest = imax(st.irest, 1);
      if (st.irest == 0) {
        st.nred = 0;
        st.rmin = alf1 * st.gnorm / st.snorm;
        st.rmax = pmin2(alf2 * st.gnorm / st.snorm, xmax / st.snorm);
      }
      if (st.iterm != 0) break;
      if (st.irest != 0) continue;  // at == 1 again
      // PYTRCS
      st.fp = st.fo;
      st.fo = st.f;
      st.po = st.p;
      for (int i = 0; i < nf; ++i) {
        v.xo[i] = v.x[i];
        v.go[i] = v.gf[i];
      }
      for (int i = 0; i < nf; ++i) {
        if (v.s[i] < 0.0) {
          if (ix[i] == 1 || ix[i] >= 3) st.rmax = pmin2(st.rmax, (kTimeLowerBound - v.x[i]) / v.s[i]);
        } else if (v.s[i] > 0.0) {
          if (ix[i] == 2 || ix[i] >= 3) st.rmax = pmin2(st.rmax, (TG_DBL_MAX - v.x[i]) / v.s[i]);
        }
      }
      if (st.rmax == 0.0) {
        at = 3;
      } else {
        st.isys = 0;
        at = 2;
      }
    }
    if (at == 2) {
      plis_ps1l01(st);
      if (st.isys == 1) {
        // next trial point: x = xo + r s on the free variables, snapped to the bounds; ask for its evaluation
        plis_dir(nf, st.r, v.s, v.xo, v.x, ix);
        plis_snap(nf, v.x, ix);
        for (int i = 0; i < nf; ++i) v.xeval[i] = v.x[i];
        st.stage = 1;
        return;
      }
      if (st.iters <= 0) {
        // line search failed: back to its start point, restart with steepest descent
        st.r = 0.0;
        st.f = st.fo;
        st.p = st.po;
        for (int i = 0; i < nf; ++i) {
          v.x[i] = v.xo[i];
          v.gf[i] = v.go[i];
        }
        st.irest = imax(st.irest, 1);
        st.ld = st.kd;
        at = 1;
        continue;
      }
      // PYTRCD (iters > 0)
      for (int i = 0; i < nf; ++i) v.xo[i] = v.x[i] - v.xo[i];
      for (int i = 0; i < nf; ++i) v.go[i] = v.gf[i] - v.go[i];
      st.po = st.r * st.po;
      st.p = st.r * st.p;
      st.dmax = 0.0;
      for (int i = 0; i < nf; ++i) {
        if (ix[i] < 0) {
          v.xo[i] = 0.0;
          v.go[i] = 0.0;
          continue;
        }
        st.dmax = pmax2(st.dmax, dabs(v.xo[i]) / pmax2(dabs(v.x[i]), 1.0));
      }
      // nlopt_stop_dx
      st.xstop = 1;
      for (int i = 0; i < nf; ++i)
        if (!relstop(v.x[i] - v.xo[i], v.x[i], x_rel, x_abs)) {
          st.xstop = 0;
          break;
        }
      at = 3;
    }
    if (at == 3) {
      for (int i = 0; i < nf; ++i) ix[i] = iabs(ix[i]);  // MXVINE
      plis_add_active(nf, &st.nfree, v.x, ix);
      if (st.xstop) {
        st.iterm = 1;
        break;
      }
      at = 0;
    }
  }
  st.code = plis_result_code(st.iterm);
  st.done = 1;
}

}  // namespace tg

#endif  // TG_PLIS_CUH_
