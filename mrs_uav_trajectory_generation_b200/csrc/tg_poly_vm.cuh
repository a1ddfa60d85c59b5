// tg_poly_vm.cuh -- Jenkins-Traub (rpoly_ak1.cpp:148-932) as a warp-scheduled MICRO-OP machine with lane refill.
// Third formulation of the extremum kernels (TG_JT_IMPL=2); same arithmetic, bit for bit, as tg_poly_naive.cuh (0)
// and the stage-level state machine of tg_poly.cuh (1), and therefore as the oracle.
//
// What the round-1 measurements said (sessions 1-2 of round 1, summarised in DESIGN.md 4.2):
//   * one thread per polynomial, direct transcription: 3.8 of 32 lanes active per instruction;
//   * stage-level state machine: 6.6 lanes, fewer instructions, but 9 k SASS instructions of code (calc_sc inlined five
//     times, ~40 IEEE divisions at ~25 instructions each) and `no_instruction` the top stall -- the L1.5 I-cache holds 32 KB;
//   * a scheduling simulation on recorded state traces: even an ideal lock-step schedule reaches 26 % lane use because a
//     warp waits for its slowest polynomial (145 passes on average, 384 worst); refilling finished lanes lifts every
//     policy by ~1.35x.
// Hence:
//   * every primitive of the iteration (synthetic division of p, calc_sc, next_k, newest, ...) exists ONCE, as a
//     micro-op; the stages are short programs over the micro-ops (kProg below) and a lane carries a program counter.
//     Lanes of different stages that need the same primitive run it together.  IEEE division and square root are
//     out-of-line (tg_poly.cuh TG_DIV / TG_SQRT).  The machine is ~1/4 of the code of the earlier variants.
//   * the polynomial of every work item is prepared by a separate, fully converged kernel (ExtremaPrepFn): the
//     convolution, trimming, end-point candidates.  The machine kernel only iterates.
//   * a lane that finishes takes the next work item of its warp's chunk (chunks come from a global counter), so the
//     warp stays full until the batch is exhausted.
// Scheduling cannot change results: lanes never exchange data.
#ifndef TG_POLY_VM_CUH_
#define TG_POLY_VM_CUH_

#include "tg_poly.cuh"

namespace tg {

constexpr int kVmPolyStride = 16;  // doubles per prepared polynomial (degree <= 15): decreasing powers, p[0] != 0

// ---- preparation: one thread per (segment, quantity) ------------------------------------------------------------------
// Writes the derivative polynomial whose zeros are the extremum candidates (eth/segment.cpp:122-145 for the horizontal
// pair, eth/polynomial.cpp:69-85 for one dimension) after findRootsJenkinsTraub's trimming (rpoly_ak1.cpp:76-120) and
// the stripping of zeros at the origin (rpoly_ak1.cpp:174-180), and the maximum over the two end-point candidates.
// Returns the degree handed to the iteration (0: nothing to iterate).
template <int DERIV, int D0, int ND>
TG_HD double vm_candidate(const double* __restrict__ coef, double t) {
  double mag = 0.0;
#pragma unroll
  for (int dim = D0; dim < D0 + ND; ++dim) {
    const double v = poly_eval_s<DERIV>(coef + dim * TG_N, t);
    mag = mag + v * v;
  }
  return dsqrt(mag);
}

template <int M>
TG_HD int vm_trim_store(const double (&ci)[M + 1], double* __restrict__ poly) {
  int last = -1;
#pragma unroll
  for (int i = 0; i <= M; i++)
    if (dabs(ci[i]) >= TG_DBL_MIN) last = i;
  if (last < 1) return 0;  // all zero, or a constant: no roots
  int low = last;
#pragma unroll
  for (int i = M; i >= 0; i--)
    if (i <= last && ci[i] != 0.0) low = i;  // lowest non-zero coefficient: `low` zeros at the origin
  const int degree = last - low;
#pragma unroll
  for (int i = 0; i <= M; i++) {
    const int dst = last - i;
    if (dst >= 0 && dst <= degree) poly[dst] = ci[i];
  }
  return degree;
}

template <int Q>
struct VmQuantity {
  static constexpr int kGroup = Q / 3;
  static constexpr int kDeriv = Q % 3 + 1;
  static constexpr int kD0 = (kGroup == 0) ? 0 : (kGroup == 1 ? 2 : 3);
  static constexpr int kND = (kGroup == 0) ? 2 : 1;
  static constexpr int kMaxDegree = (kGroup == 0) ? 2 * (TG_N - kDeriv) - 3 : TG_N - kDeriv - 2;
};

template <int Q>
TG_HD int extrema_prepare(const double* __restrict__ coef, double T, double* __restrict__ poly, double* best_out) {
  using VQ = VmQuantity<Q>;
  constexpr int DERIV = VQ::kDeriv;
  double best = TG_DBL_LOWEST;
  if (0.0 > T) {  // computeMinMaxCandidates rejects an inverted interval (eth/polynomial.cpp:39-42)
    *best_out = best;
    return 0;
  }
  {
    const double m0 = vm_candidate<DERIV, VQ::kD0, VQ::kND>(coef, 0.0);
    if (best < m0) best = m0;
    const double m1 = vm_candidate<DERIV, VQ::kD0, VQ::kND>(coef, T);
    if (best < m1) best = m1;
  }
  *best_out = best;
  if constexpr (VQ::kGroup == 0) {
    constexpr int n_d = TG_N - DERIV, n_dd = n_d - 1, len = n_d + n_dd - 1, M = len - 1;
    double acc[M + 1];
#pragma unroll
    for (int i = 0; i < len; ++i) acc[i] = 0.0;
#pragma unroll
    for (int dim = 0; dim < 2; ++dim) {
      const double* c = coef + dim * TG_N;
      double dc[n_d], ddc[n_dd];
#pragma unroll
      for (int jx = 0; jx < n_d; ++jx) dc[jx] = c[jx + DERIV] * bcoef(DERIV, jx + DERIV);
#pragma unroll
      for (int jx = 0; jx < n_dd; ++jx) ddc[jx] = c[jx + DERIV + 1] * bcoef(DERIV + 1, jx + DERIV + 1);
#pragma unroll
      for (int i = 0; i < len; ++i) {
        double cv = 0.0;
        const int data_idx = i - n_dd + 1;
        const int lower = (0 > -data_idx) ? 0 : -data_idx, upper = (n_dd < n_d - data_idx) ? n_dd : n_d - data_idx;
#pragma unroll
        for (int kidx = lower; kidx < upper; ++kidx) cv = cv + ddc[n_dd - 1 - kidx] * dc[data_idx + kidx];
        acc[i] = acc[i] + cv;
      }
    }
    return vm_trim_store<M>(acc, poly);
  } else {
    constexpr int M = TG_N - DERIV - 2;
    const double* c = coef + VQ::kD0 * TG_N;
    double ddc[M + 1];
#pragma unroll
    for (int jx = 0; jx <= M; ++jx) ddc[jx] = c[jx + DERIV + 1] * bcoef(DERIV + 1, jx + DERIV + 1);
    return vm_trim_store<M>(ddc, poly);
  }
}

// ---- the machine -----------------------------------------------------------------------------------------------------------
struct JtVm {
  enum Op {
    kRootBegin, kChop, kNewton, kKInit, kShiftBegin,  // rpoly_ak1 main loop
    kQsd, kCalcSc, kNextK, kNewest,                    // primitives (advance the program counter)
    kFsPrepEnd, kFixedB,                               // Fxshfr_ak1 glue
    kQuadHead, kQuadMid, kQuadTail,                    // QuadIT_ak1
    kRealStep,                                         // RealIT_ak1
    kEmit, kDone, kNumOps
  };
  // programs over the primitives; a logic op starts one with start(), a primitive op steps to the next entry
  enum Prog { kProgFsPrep = 0, kProgFixed = 3, kProgQuad1 = 7, kProgQuadCluster = 9, kProgQuad2 = 20, kProgLen = 25 };
  TG_HD static constexpr unsigned long long pack(int op, int pos) { return (unsigned long long)op << (4 * pos); }
  TG_HD static int prog(int pc) {
    // FsPrep: Qsd CalcSc FsPrepEnd | Fixed: NextK CalcSc Newest FixedB | Quad1: Qsd QuadMid |
    // QuadCluster: Qsd (CalcSc NextK)x5, falls into Quad2 | Quad2: CalcSc NextK CalcSc Newest QuadTail
    // Four bits per entry, packed into two immediates (no table in memory).
    constexpr unsigned long long lo = pack(kQsd, 0) | pack(kCalcSc, 1) | pack(kFsPrepEnd, 2) | pack(kNextK, 3) | pack(kCalcSc, 4) |
                                      pack(kNewest, 5) | pack(kFixedB, 6) | pack(kQsd, 7) | pack(kQuadMid, 8) | pack(kQsd, 9) |
                                      pack(kCalcSc, 10) | pack(kNextK, 11) | pack(kCalcSc, 12) | pack(kNextK, 13) | pack(kCalcSc, 14) |
                                      pack(kNextK, 15);
    constexpr unsigned long long hi = pack(kCalcSc, 0) | pack(kNextK, 1) | pack(kCalcSc, 2) | pack(kNextK, 3) | pack(kCalcSc, 4) |
                                      pack(kNextK, 5) | pack(kCalcSc, 6) | pack(kNewest, 7) | pack(kQuadTail, 8);
    const unsigned long long w = (pc < 16) ? lo : hi;
    return (int)((w >> (4 * (pc & 15))) & 15ull);
  }

  WArr p, qp, K, qk;  // work arrays (degree + 1 entries each), strided
  double* svk;        // K at the start of a third stage; K at the start of the shifts (rarely read back)
  double* tmp;
  int N, NN, state, pc;
  double a, b, c, d, e, f, g, h, a1, a3, a7;
  double szr, szi, lzr, lzi;
  double xx, yy, bnd, x, xm, dx, ff;
  int jj;
  double u, v, ui, vi, nu, nv, cu, cv, betas, betav, oss, ots, otv, ovv, s, ss, ts, tss, tv, tvv, vv;
  int j, L2, tFlag, spass, vpass, stry, vtry, first, prep_tail;
  double qu, qv, qomp, qrelstp, qmp;
  int qj, qtried;
  double rs, rt, romp;
  int rj;
  int em_n, em_after;

  TG_HD void start(int program) {
    pc = program;
    state = prog(program);
  }
  TG_HD void advance() {
    pc = pc + 1;
    state = prog(pc);
  }

  TG_HD static void quad(double qa, double b1, double qc, double* sr, double* si, double* lr, double* li) {  // rpoly_ak1.cpp:881-932
    *sr = *si = *lr = *li = 0.0;
    if (qa == 0) {
      *sr = ((b1 != 0) ? -TG_DIV(qc, b1) : *sr);
      return;
    }
    if (qc == 0) {
      *lr = -TG_DIV(b1, qa);
      return;
    }
    const double bb = b1 / 2.0;
    double dd, ee;
    if (dabs(bb) < dabs(qc)) {
      ee = ((qc >= 0) ? qa : -qa);
      ee = -ee + bb * TG_DIV(bb, dabs(qc));
      dd = TG_SQRT(dabs(ee)) * TG_SQRT(dabs(qc));
    } else {
      ee = -(TG_DIV(qa, bb) * TG_DIV(qc, bb)) + 1.0;
      dd = TG_SQRT(dabs(ee)) * (dabs(bb));
    }
    if (ee >= 0) {
      dd = ((bb >= 0) ? -dd : dd);
      *lr = TG_DIV(-bb + dd, qa);
      *sr = ((*lr != 0) ? TG_DIV(TG_DIV(qc, *lr), qa) : *sr);
    } else {
      *lr = *sr = -TG_DIV(bb, qa);
      *si = dabs(TG_DIV(dd, qa));
      *li = -(*si);
    }
  }

  // ---- control-flow glue of Fxshfr_ak1's third-stage do-while (rpoly_ak1.cpp:459-523) --------------------------
  TG_HD void restore_k() {
#pragma unroll 1
    for (int i = 0; i < N; i++) K[i] = svk[i];
  }
  TG_HD void begin_quad() {
    qj = 0;
    qtried = 0;
    qu = ui;
    qv = vi;
    qomp = 0.0;
    qrelstp = 0.0;
    state = kQuadHead;
  }
  TG_HD void begin_real() {
    rj = 0;
    rs = s;
    rt = 0.0;
    romp = 0.0;
    state = kRealStep;
  }
  TG_HD void stage3_top() {  // start of one pass of the do { } while (vpass && !vtry)
    const bool shortcut = first && ((spass) && (!vpass || (tss < tvv)));
    first = 0;
    if (!shortcut) begin_quad();
    else begin_real();
  }
  TG_HD void begin_fsprep(int tail) {  // quad_sd on p + calc_sc at the fixed shift (rpoly_ak1.cpp:411-413, 527-528)
    prep_tail = tail;
    cu = u;
    cv = v;
    start(kProgFsPrep);
  }
  TG_HD void stage3_cond() {
    if (vpass && !vtry) stage3_top();
    else begin_fsprep(1);  // re-compute qp and the scalars, then finish this fixed-shift step
  }
  // which glue runs after a failed iteration is decided here; the restore of K (one loop in the whole machine) and the
  // decision itself happen in kFixedB's `resume` path
  TG_HD void fixed_next() {  // head of the fixed-shift loop (rpoly_ak1.cpp:415)
    if (j >= L2) {
#pragma unroll 1
      for (int i = 0; i < N; i++) K[i] = tmp[i];  // unsuccessful shift: restore K, next jj
      jj++;
      state = kShiftBegin;
    } else {
      cu = u;
      cv = v;
      start(kProgFixed);
    }
  }
  TG_HD void quad_failed() {
    vtry = 1;
    betav = betav * 0.25;
    restore_k();
    if (stry || (!spass)) stage3_cond();
    else begin_real();
  }
  TG_HD void real_failed(int iFlag) {
    stry = 1;
    betas = betas * 0.25;
    if (iFlag != 0) {
      ui = -(s + s);
      vi = s * s;
      stage3_cond();  // `continue`: straight to the loop condition, K is NOT restored
    } else {
      restore_k();
      stage3_cond();
    }
  }
  TG_HD void root_found(int nz) {  // rpoly_ak1.cpp:338-356; the zeros are reported by kEmit
    em_n = nz;
    em_after = kRootBegin;
    NN = NN - nz;
    N = NN - 1;
#pragma unroll 1
    for (int i = 0; i < NN; i++) p[i] = qp[i];
    state = kEmit;
  }

  // p[0..degree] holds the coefficients (decreasing powers, p[0] != 0, zeros at the origin stripped)
  TG_HD void begin(int degree) {
    N = degree;
    NN = N + 1;
    xx = 0x1.6a09e667f3bcdp-1;  // sqrt(0.5)
    yy = -xx;
    state = (degree >= 1) ? kRootBegin : kDone;
    pc = 0;
    em_n = 0;
    em_after = kDone;
  }

  // One micro-op.  `Emit` reports a zero: emit(re, im).
  template <class Emit>
  TG_HD void step(int cur, Emit& emit) {
    const double lb2 = 0x1.62e42fefa39efp-1;   // log(2.0)
    const double lo = TG_FLT_MIN / TG_DBL_EPSILON;
    const double cosr = -0x1.1db8f6d6a512ap-4;  // cos(94 deg) as glibc returns it for 94.0 * (3.14159265358979323846 / 180)
    const double sinr = 0x1.fec0b7170fff6p-1;   // sin(94 deg)
    if (cur == kQsd) {  // QuadSD_ak1 on p (rpoly_ak1.cpp:543-559)
      double bb, aa;
      qp[0] = bb = p[0];
      qp[1] = aa = -(bb * cu) + p[1];
#pragma unroll 1
      for (int i = 2; i < NN; i++) {
        const double t = -(aa * cu + bb * cv) + p[i];
        qp[i] = t;
        bb = aa;
        aa = t;
      }
      a = aa;
      b = bb;
      advance();
    } else if (cur == kCalcSc) {  // calcSC_ak1 (rpoly_ak1.cpp:561-602)
      {
        double bb, aa;
        qk[0] = bb = K[0];
        qk[1] = aa = -(bb * cu) + K[1];
#pragma unroll 1
        for (int i = 2; i < N; i++) {
          const double t = -(aa * cu + bb * cv) + K[i];
          qk[i] = t;
          bb = aa;
          aa = t;
        }
        c = aa;
        d = bb;
      }
      int tf;
      if ((dabs(c) <= (10.0 * TG_DBL_EPSILON * dabs(K[N - 1]))) && (dabs(d) <= (10.0 * TG_DBL_EPSILON * dabs(K[N - 2])))) {
        tf = 3;
      } else {
        h = cv * b;
        if (dabs(d) >= dabs(c)) {
          e = TG_DIV(a, d);
          f = TG_DIV(c, d);
          g = cu * b;
          a3 = e * (g + a) + h * TG_DIV(b, d);
          a1 = -a + f * b;
          a7 = h + (f + cu) * a;
          tf = 2;
        } else {
          e = TG_DIV(a, c);
          f = TG_DIV(d, c);
          g = e * cu;
          a3 = e * a + (g + TG_DIV(h, c)) * b;
          a1 = -(a * TG_DIV(d, c)) + b;
          a7 = g * d + h * f + a;
          tf = 1;
        }
      }
      tFlag = tf;
      advance();
    } else if (cur == kNextK) {  // nextK_ak1 (rpoly_ak1.cpp:604-645)
      if (tFlag == 3) {
        K[1] = K[0] = 0.0;
#pragma unroll 1
        for (int i = 2; i < N; i++) K[i] = qk[i - 2];
      } else {
        const double temp = ((tFlag == 1) ? b : a);
        if (dabs(a1) > (10.0 * TG_DBL_EPSILON * dabs(temp))) {
          a7 = TG_DIV(a7, a1);
          a3 = TG_DIV(a3, a1);
          K[0] = qp[0];
          K[1] = -(a7 * qp[0]) + qp[1];
#pragma unroll 1
          for (int i = 2; i < N; i++) K[i] = -(a7 * qp[i - 1]) + a3 * qk[i - 2] + qp[i];
        } else {
          K[0] = 0.0;
          K[1] = -a7 * qp[0];
#pragma unroll 1
          for (int i = 2; i < N; i++) K[i] = -(a7 * qp[i - 1]) + a3 * qk[i - 2];
        }
      }
      advance();
    } else if (cur == kNewest) {  // newest_ak1 (rpoly_ak1.cpp:647-683)
      nv = nu = 0.0;
      if (tFlag != 3) {
        double a4, a5;
        if (tFlag != 2) {
          a4 = a + cu * b + h * f;
          a5 = c + (cu + cv * f) * d;
        } else {
          a4 = (a + g) * f + h;
          a5 = (f + cu) * c + cv * d;
        }
        const double pN = p[N], pN1 = p[N - 1], kN1 = K[N - 1], kN2 = K[N - 2];
        const double b1 = TG_DIV(-kN1, pN);
        const double b2 = TG_DIV(-(kN2 + b1 * pN1), pN);
        const double c1 = cv * b2 * a1;
        const double c2 = b1 * a7;
        const double c3 = b1 * b1 * a3;
        const double c4 = -(c2 + c3) + c1;
        const double temp = -c4 + a5 + b1 * a4;
        if (temp != 0.0) {
          nu = -TG_DIV(cu * (c3 + c2) + cv * (b1 * a1 + b2 * a7), temp) + cu;
          nv = cv * (1.0 + TG_DIV(c4, temp));
        }
      }
      advance();
    } else if (cur == kFixedB) {  // rest of one pass of the fixed-shift loop (rpoly_ak1.cpp:421-538)
      ui = nu;
      vi = nv;
      vv = vi;
      const double kN1 = K[N - 1];
      ss = ((kN1 != 0.0) ? -TG_DIV(p[N], kN1) : 0.0);
      ts = tv = 1.0;
      bool stage3 = false;
      if ((j != 0) && (tFlag != 3)) {
        tv = ((vv != 0.0) ? dabs(TG_DIV(vv - ovv, vv)) : tv);
        ts = ((ss != 0.0) ? dabs(TG_DIV(ss - oss, ss)) : ts);
        tvv = ((tv < otv) ? tv * otv : 1.0);
        tss = ((ts < ots) ? ts * ots : 1.0);
        vpass = ((tvv < betav) ? 1 : 0);
        spass = ((tss < betas) ? 1 : 0);
        if ((spass) || (vpass)) {
#pragma unroll 1
          for (int i = 0; i < N; i++) svk[i] = K[i];
          s = ss;
          stry = vtry = 0;
          first = 1;
          stage3 = true;
          stage3_top();
        }
      }
      if (!stage3) {
        ovv = vv;
        oss = ss;
        otv = tv;
        ots = ts;
        j++;
        fixed_next();
      }
    } else if (cur == kFsPrepEnd) {
      if (prep_tail) {
        ovv = vv;
        oss = ss;
        otv = tv;
        ots = ts;
        j++;
      }
      fixed_next();
    } else if (cur == kQuadHead) {  // QuadIT_ak1, top of the do-while (rpoly_ak1.cpp:698-707)
      quad(1.0, qu, qv, &szr, &szi, &lzr, &lzi);
      if (dabs(dabs(szr) - dabs(lzr)) > 0.01 * dabs(lzr)) {
        quad_failed();
      } else {
        cu = qu;
        cv = qv;
        start(kProgQuad1);
      }
    } else if (cur == kQuadMid) {  // rpoly_ak1.cpp:709-757
      const double mp = dabs(-(szr * b) + a) + dabs(szi * b);
      const double zm = TG_SQRT(dabs(qv));
      double ee = 2.0 * dabs(qp[0]);
      const double t = -(szr * b);
#pragma unroll 1
      for (int i = 1; i < N; i++) ee = ee * zm + dabs(qp[i]);
      ee = ee * zm + dabs(a + t);
      ee = (9.0 * ee + 2.0 * dabs(t) - 7.0 * (dabs(a + t) + zm * dabs(b))) * TG_DBL_EPSILON;
      if (mp <= 20.0 * ee) {
        root_found(2);
      } else {
        qj++;
        if (qj > 20) {
          quad_failed();
        } else {
          bool cluster = false;
          if (qj >= 2) {
            if ((qrelstp <= 0.01) && (mp >= qomp) && (!qtried)) {
              // a cluster stalls the convergence: five fixed-shift steps close to it
              qrelstp = ((qrelstp < TG_DBL_EPSILON) ? TG_SQRT(TG_DBL_EPSILON) : TG_SQRT(qrelstp));
              qu = qu - qu * qrelstp;
              qv = qv + qv * qrelstp;
              qtried = 1;
              qj = 0;
              cluster = true;
            }
          }
          qomp = mp;
          cu = qu;
          cv = qv;
          start(cluster ? kProgQuadCluster : kProgQuad2);
        }
      }
    } else if (cur == kQuadTail) {  // rpoly_ak1.cpp:768-779
      if (nv != 0) {
        qrelstp = dabs(TG_DIV(-qv + nv, nv));
        qu = nu;
        qv = nv;
        state = kQuadHead;
      } else {
        quad_failed();
      }
    } else if (cur == kRealStep) {  // one pass of RealIT_ak1's loop (rpoly_ak1.cpp:798-875)
      const int nm1 = N - 1;
      double pv;
      qp[0] = pv = p[0];
#pragma unroll 1
      for (int i = 1; i < NN; i++) qp[i] = pv = pv * rs + p[i];
      const double mp = dabs(pv);
      const double ms = dabs(rs);
      double ee = 0.5 * dabs(qp[0]);
#pragma unroll 1
      for (int i = 1; i < NN; i++) ee = ee * ms + dabs(qp[i]);
      if (mp <= 20.0 * TG_DBL_EPSILON * (2.0 * ee - mp)) {
        szr = rs;
        szi = 0.0;
        root_found(1);
      } else {
        rj++;
        if (rj > 10) {
          real_failed(0);
        } else if ((rj >= 2) && ((dabs(rt) <= 0.001 * dabs(-rt + rs)) && (mp > romp))) {
          s = rs;  // a cluster near the real axis: hand the iterate to the quadratic iteration
          real_failed(1);
        } else {
          romp = mp;
          double kv;
          qk[0] = kv = K[0];
#pragma unroll 1
          for (int i = 1; i < N; i++) qk[i] = kv = kv * rs + K[i];
          if (dabs(kv) > dabs(K[nm1]) * 10.0 * TG_DBL_EPSILON) {
            rt = -TG_DIV(pv, kv);
            K[0] = qp[0];
#pragma unroll 1
            for (int i = 1; i < N; i++) K[i] = rt * qk[i - 1] + qp[i];
          } else {
            K[0] = 0.0;
#pragma unroll 1
            for (int i = 1; i < N; i++) K[i] = qk[i - 1];
          }
          kv = K[0];
#pragma unroll 1
          for (int i = 1; i < N; i++) kv = kv * rs + K[i];
          rt = ((dabs(kv) > (dabs(K[nm1]) * 10.0 * TG_DBL_EPSILON)) ? -TG_DIV(pv, kv) : 0.0);
          rs = rs + rt;
        }
      }
    } else if (cur == kRootBegin) {  // rpoly_ak1.cpp:188-283
      if (N < 1) {
        state = kDone;
      } else if (N <= 2) {
        if (N < 2) {
          szr = -TG_DIV(p[1], p[0]);
          szi = 0.0;
          em_n = 1;
        } else {
          quad(p[0], p[1], p[2], &szr, &szi, &lzr, &lzi);
          em_n = 2;
        }
        em_after = kDone;
        state = kEmit;
      } else {
        double moduli_max = 0.0, moduli_min = TG_FLT_MAX;
#pragma unroll 1
        for (int i = 0; i < NN; i++) {
          const double xa = dabs(p[i]);
          if (xa > moduli_max) moduli_max = xa;
          if ((xa != 0) && (xa < moduli_min)) moduli_min = xa;
        }
        double sc = TG_DIV(lo, moduli_min);
        if (((sc <= 1.0) && (moduli_max >= 10)) || ((sc > 1.0) && (TG_DIV(TG_FLT_MAX, sc) >= moduli_max))) {
          sc = ((sc == 0) ? TG_FLT_MIN : sc);
          const int l = (int)(TG_DIV(tgdm::dlog_k(sc), lb2) + 0.5);
          const double factor = tgdm::scalb(1.0, l);
          if (factor != 1.0) {
#pragma unroll 1
            for (int i = 0; i < NN; i++) p[i] = p[i] * factor;
          }
        }
        // upper estimate of the lower bound on the zero moduli; pt[i] = |p[i]|, pt[N] = -|p[N]|
        const double ptN = -dabs(p[N]), pt0 = dabs(p[0]), ptNM1 = dabs(p[N - 1]);
        x = tgdm::dexp_k(TG_DIV(tgdm::dlog_k(-ptN) - tgdm::dlog_k(pt0), (double)N));
        if (ptNM1 != 0) {
          const double xm_ = TG_DIV(-ptN, ptNM1);
          x = ((xm_ < x) ? xm_ : x);
        }
        xm = x;
        state = kChop;
      }
    } else if (cur == kChop) {  // one pass of: do { x = xm; xm = 0.1 x; ff = pt(xm) } while (ff > 0)
      x = xm;
      xm = 0.1 * x;
      ff = dabs(p[0]);
#pragma unroll 1
      for (int i = 1; i < N; i++) ff = ff * xm + dabs(p[i]);
      ff = ff * xm + (-dabs(p[N]));
      if (!(ff > 0)) {
        dx = x;
        state = kNewton;
      }
    } else if (cur == kNewton) {  // one pass of: while (|dx/x| > 0.005) { Newton step }
      if (dabs(TG_DIV(dx, x)) > 0.005) {
        double df;
        df = ff = dabs(p[0]);
#pragma unroll 1
        for (int i = 1; i < N; i++) {
          ff = x * ff + dabs(p[i]);
          df = x * df + ff;
        }
        ff = x * ff + (-dabs(p[N]));
        dx = TG_DIV(ff, df);
        x = x - dx;
      } else {
        bnd = x;
        state = kKInit;
      }
    } else if (cur == kKInit) {  // K = p'/N and five no-shift steps (rpoly_ak1.cpp:285-320)
      const int NM1 = N - 1;
#pragma unroll 1
      for (int i = 1; i < N; i++) K[i] = TG_DIV((double)(N - i) * p[i], (double)N);
      K[0] = p[0];
      const double aa = p[N], bb = p[NM1];
      int zerok = ((K[NM1] == 0) ? 1 : 0);
#pragma unroll 1
      for (int q = 0; q < 5; q++) {
        const double cc = K[NM1];
        if (zerok) {
#pragma unroll 1
          for (int i = 0; i < NM1; i++) {
            const int jx = NM1 - i;
            K[jx] = K[jx - 1];
          }
          K[0] = 0;
          zerok = ((K[NM1] == 0) ? 1 : 0);
        } else {
          const double t = TG_DIV(-aa, cc);
#pragma unroll 1
          for (int i = 0; i < NM1; i++) {
            const int jx = NM1 - i;
            K[jx] = t * K[jx - 1] + p[jx];
          }
          K[0] = p[0];
          zerok = ((dabs(K[NM1]) <= dabs(bb) * TG_DBL_EPSILON * 10.0) ? 1 : 0);
        }
      }
#pragma unroll 1
      for (int i = 0; i < N; i++) tmp[i] = K[i];
      jj = 1;
      state = kShiftBegin;
    } else if (cur == kShiftBegin) {  // next shift of the jj loop (rpoly_ak1.cpp:324-336) + Fxshfr prologue (404-411)
      if (jj > 20) {
        state = kDone;  // no convergence after 20 shifts: the zeros found so far stand
      } else {
        const double xxx = -(sinr * yy) + cosr * xx;
        yy = sinr * xx + cosr * yy;
        xx = xxx;
        const double sr = bnd * xx;
        betav = betas = 0.25;
        u = -(2.0 * sr);
        oss = sr;
        ovv = v = bnd;
        ots = otv = 0.0;
        j = 0;
        L2 = 20 * jj;
        begin_fsprep(0);
      }
    } else if (cur == kEmit) {
#pragma unroll 1
      for (int q = 0; q < em_n; ++q) emit(q == 0 ? szr : lzr, q == 0 ? szi : lzi);
      state = em_after;
    }
  }
};

// candidate filter + magnitude (eth/polynomial.cpp:50-57, eth/segment.cpp:172-178, 203-209) for quantity Q
template <int Q>
struct VmEmit {
  const double* coef;
  double T;
  double best;
  TG_HD void operator()(double re, double im) {
    using VQ = VmQuantity<Q>;
    if (dabs(im) > TG_DBL_EPSILON) return;
    if (re < 0.0 || re > T) return;
    const double mag = vm_candidate<VQ::kDeriv, VQ::kD0, VQ::kND>(coef, re);
    if (best < mag) best = mag;
  }
};

// shared-memory doubles per thread for the four work arrays of quantity Q
template <int Q>
struct VmScratch {
  static constexpr int kShared = 4 * (VmQuantity<Q>::kMaxDegree + 1);
};

// Runs the machine for ONE prepared polynomial to completion (host emulation, and the single-item entry point).
template <int Q>
TG_HD double vm_run_single(const double* __restrict__ coef, double T, const double* __restrict__ poly, int degree, double best0, double* scratch,
                           int stride) {
  constexpr int M = VmQuantity<Q>::kMaxDegree;
  double svk[M + 1], tmp[M + 1];
  JtVm m;
  m.p = WArr{scratch, stride};
  m.qp = WArr{scratch + (size_t)(M + 1) * stride, stride};
  m.K = WArr{scratch + (size_t)2 * (M + 1) * stride, stride};
  m.qk = WArr{scratch + (size_t)3 * (M + 1) * stride, stride};
  m.svk = svk;
  m.tmp = tmp;
  VmEmit<Q> emit{coef, T, best0};
  for (int i = 0; i <= degree; ++i) m.p[i] = poly[i];
  m.begin(degree);
  while (m.state != JtVm::kDone) m.step(m.state, emit);
  return emit.best;
}

}  // namespace tg

#endif  // TG_POLY_VM_CUH_
