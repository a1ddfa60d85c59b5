// tg_segment.cuh -- per-segment matrices for one segment time T: Q, A^-1 (Schur form) and H = (A^-T Q) A^-1.
// Replaces PolynomialOptimization<10>::updateSegmentTimes + the per-segment part of constructR
// (reference: lin_impl.h:288-304, 605-618, 112-121, 147-177, 317-320).  One thread computes one record;
// all arrays are statically indexed so they live in registers.  Structural zeros of A^-1 and Q are skipped:
// adding an exact zero does not change an IEEE sum, so the result equals the dense evaluation.
#ifndef TG_SEGMENT_CUH_
#define TG_SEGMENT_CUH_

#include "tg_common.cuh"

namespace tg {

// 5x5 inverse: LU with first-maximum row pivoting, multipliers by division, then identity columns solved by
// forward (ascending j) and back substitution (ascending j).  DESIGN.md "numeric contract" item C.
TG_HD void inverse5(const double (&Din)[5][5], double (&out)[5][5]) {
  double lu[5][5];
  int perm[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    perm[i] = i;
#pragma unroll
    for (int j = 0; j < 5; ++j) lu[i][j] = Din[i][j];
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    int piv = k;
    double best = dabs(lu[k][k]);
#pragma unroll
    for (int i = k + 1; i < 5; ++i) {
      const double a = dabs(lu[i][k]);
      if (a > best) { best = a; piv = i; }
    }
    // swap rows k and piv with static indices (select instead of dynamic addressing)
#pragma unroll
    for (int i = k + 1; i < 5; ++i) {
      const bool sw = (piv == i);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const double a = lu[k][j], b = lu[i][j];
        lu[k][j] = sw ? b : a;
        lu[i][j] = sw ? a : b;
      }
      const int pa = perm[k], pb = perm[i];
      perm[k] = sw ? pb : pa;
      perm[i] = sw ? pa : pb;
    }
#pragma unroll
    for (int i = k + 1; i < 5; ++i) lu[i][k] = lu[i][k] / lu[k][k];
#pragma unroll
    for (int i = k + 1; i < 5; ++i)
#pragma unroll
      for (int j = k + 1; j < 5; ++j) lu[i][j] = lu[i][j] - lu[i][k] * lu[k][j];
  }
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    double y[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double s = (perm[i] == c) ? 1.0 : 0.0;
#pragma unroll
      for (int j = 0; j < i; ++j) s = s - lu[i][j] * y[j];
      y[i] = s;
    }
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      double s = y[i];
#pragma unroll
      for (int j = i + 1; j < 5; ++j) s = s - lu[i][j] * y[j];
      y[i] = s / lu[i][i];
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) out[i][c] = y[i];
  }
}

// Computes the record for one (segment, time).  r = derivative whose squared integral is minimised (2..4).
template <int R>
TG_HD void setup_segment_record_r(double T, double* __restrict__ rec) {
  constexpr int NQ = TG_N - R;           // non-zero rows/cols of Q start at index R
  constexpr int EMAX = (TG_N - 1 - R) * 2 + 1;
  // --- Q (lin_impl.h:605-618): Q[i][j] = B[r][i]*B[r][j]*pow(T, e)*2/e, e = i+j-2r+1, left to right
  double pw[EMAX];
  tgdm::powers(T, EMAX, pw);
  double Q[NQ][NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i)
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const int e = i + j + 1;  // (i+R) + (j+R) - 2R + 1
      Q[i][j] = bcoef(R, i + R) * bcoef(R, j + R) * pw[e - 1] * 2.0 / (double)e;
    }
  // --- A rows at t = T (eth/polynomial.h:208-226): entry[j] = B[k][j]*tp, tp = tp*T (plain chain)
  double tp[TG_N];  // tp[m] = T^m by repeated multiplication, m >= 1
  tp[0] = 1.0;
  tp[1] = T;
#pragma unroll
  for (int m = 2; m < TG_N; ++m) tp[m] = tp[m - 1] * T;
  const bool tzero = dabs(T) < TG_DBL_EPSILON;
  double Dm[5][5], Cm[5][5];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      Dm[k][j] = tzero ? 0.0 : bcoef(k, 5 + j) * tp[5 + j - k];
      Cm[k][j] = (j < k) ? 0.0 : ((j == k) ? bcoef(k, k) : (tzero ? 0.0 : bcoef(k, j) * tp[j - k]));
    }
  // --- A^-1 (lin_impl.h:147-177)
  double Dinv[5][5];
  inverse5(Dm, Dinv);
  const double a_inv[5] = {1.0 / 1.0, 1.0 / 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0};
  double X[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      double m = (-Dinv[i][0]) * Cm[0][j];
#pragma unroll
      for (int k = 1; k <= j; ++k) m = m + (-Dinv[i][k]) * Cm[k][j];  // C[k][j] == 0 for k > j
      X[i][j] = m * a_inv[j];
    }
  // the record is written with 16-byte stores: one thread owns 1488 contiguous bytes, and the kernel is bound by the
  // number of store requests, not by bytes (round-1 measurement: 2.9 TB/s with 8-byte stores)
  static_assert(TG_REC_DINV == 0 && TG_REC_X == 25 && TG_REC_Q == 50 && (TG_REC_H % 2) == 0, "record layout");
#pragma unroll
  for (int e = 0; e < 50; e += 2) {
    const double v0 = (e < 25) ? Dinv[e / 5][e % 5] : X[(e - 25) / 5][(e - 25) % 5];
    const double v1 = (e + 1 < 25) ? Dinv[(e + 1) / 5][(e + 1) % 5] : X[(e + 1 - 25) / 5][(e + 1 - 25) % 5];
    store2(rec + e, v0, v1);
  }
  {
    double qt[36];  // packed upper triangle (tg_qtri); entries beyond NQ stay zero
#pragma unroll
    for (int e = 0; e < 36; ++e) qt[e] = 0.0;
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
      for (int j = i; j < NQ; ++j) qt[tg_qtri(i, j)] = Q[i][j];
#pragma unroll
    for (int e = 0; e < 36; e += 2) store2(rec + TG_REC_Q + e, qt[e], qt[e + 1]);
  }
  // --- H = (A^-T Q) A^-1 (lin_impl.h:320), inner sums over ascending k, zero terms skipped
  // A^-1 = [ diag(a_inv) 0 ; X Dinv ].  Row a of W = A^-T Q :  W[a][b] = sum_k Ainv[k][a] Q[k][b]
#pragma unroll
  for (int a = 0; a < TG_N; ++a) {
    double W[NQ];  // W[a][R + b]
#pragma unroll
    for (int b = 0; b < NQ; ++b) {
      double s;
      if (a < 5) {
        // k = a contributes only when a >= R (Q rows below R are zero); then k = 5..9 through X
        bool have = false;
        s = 0.0;
        if (a >= R) { s = a_inv[a] * Q[(a >= R) ? a - R : 0][b]; have = true; }
#pragma unroll
        for (int k = 5; k < TG_N; ++k) {
          const double t = X[k - 5][a] * Q[k - R][b];
          s = have ? s + t : t;
          have = true;
        }
      } else {
        s = Dinv[0][a - 5] * Q[5 - R][b];
#pragma unroll
        for (int k = 6; k < TG_N; ++k) s = s + Dinv[k - 5][a - 5] * Q[k - R][b];
      }
      W[b] = s;
    }
    // H[a][b] = sum_k W[a][k] Ainv[k][b], k ascending from R
    double hrow[TG_N];
#pragma unroll
    for (int b = 0; b < TG_N; ++b) {
      double s;
      if (b < 5) {
        bool have = false;
        s = 0.0;
        if (b >= R) { s = W[(b >= R) ? b - R : 0] * a_inv[b]; have = true; }
#pragma unroll
        for (int k = 5; k < TG_N; ++k) {
          const double t = W[k - R] * X[k - 5][b];
          s = have ? s + t : t;
          have = true;
        }
      } else {
        s = W[5 - R] * Dinv[0][b - 5];
#pragma unroll
        for (int k = 6; k < TG_N; ++k) s = s + W[k - R] * Dinv[k - 5][b - 5];
      }
      hrow[b] = s;
    }
#pragma unroll
    for (int b = 0; b < TG_N; b += 2) store2(rec + TG_REC_H + a * TG_N + b, hrow[b], hrow[b + 1]);
  }
}

// ---- the same record in two stages (Mellinger evaluations: 3 records per segment per evaluation, the bulk of the setup work) ----
// One thread per record needs the whole of Q, Dinv and X in registers next to the 10x10 product: 254 registers, two blocks per SM,
// and every thread walks 4.7 kflop of dependent arithmetic (ncu, round 2: issue slots 10 % busy, 29 % of the HBM roofline).
// Stage 1 (setup_record_head, one thread per record) computes Q, Dinv and X -- the serial part, 28 % of the flops -- and writes the
// first 86 doubles of the record.  Stage 2 (setup_record_hrow, one thread per (record, row a)) reads them back (the ten threads of a
// record sit in one warp and read the same addresses: broadcast loads from L1) and computes row a of H = (A^-T Q) A^-1: 400 flop
// and ~60 live values per thread, so eight blocks per SM, and the ten rows of a record leave the warp as 800 contiguous bytes.
// Every element is computed by exactly the expression the one-thread version uses (same operands, same order), so the records are
// bit-identical; Q is read back through its packed triangle, which is what the coefficient stage does too.
template <int R>
TG_HD void setup_record_head_r(double T, double* __restrict__ rec) {
  constexpr int NQ = TG_N - R;
  constexpr int EMAX = (TG_N - 1 - R) * 2 + 1;
  double pw[EMAX];
  tgdm::powers(T, EMAX, pw);
  {
    double qt[36];  // packed upper triangle (tg_qtri); entries beyond NQ stay zero
#pragma unroll
    for (int e = 0; e < 36; ++e) qt[e] = 0.0;
#pragma unroll
    for (int i = 0; i < NQ; ++i)
#pragma unroll
      for (int j = i; j < NQ; ++j) {
        const int e = i + j + 1;
        qt[tg_qtri(i, j)] = bcoef(R, i + R) * bcoef(R, j + R) * pw[e - 1] * 2.0 / (double)e;
      }
#pragma unroll
    for (int e = 0; e < 36; e += 2) store2(rec + TG_REC_Q + e, qt[e], qt[e + 1]);
  }
  double tp[TG_N];
  tp[0] = 1.0;
  tp[1] = T;
#pragma unroll
  for (int m = 2; m < TG_N; ++m) tp[m] = tp[m - 1] * T;
  const bool tzero = dabs(T) < TG_DBL_EPSILON;
  double Dm[5][5], Cm[5][5];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      Dm[k][j] = tzero ? 0.0 : bcoef(k, 5 + j) * tp[5 + j - k];
      Cm[k][j] = (j < k) ? 0.0 : ((j == k) ? bcoef(k, k) : (tzero ? 0.0 : bcoef(k, j) * tp[j - k]));
    }
  double Dinv[5][5];
  inverse5(Dm, Dinv);
  const double a_inv[5] = {1.0 / 1.0, 1.0 / 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0};
  double X[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      double m = (-Dinv[i][0]) * Cm[0][j];
#pragma unroll
      for (int k = 1; k <= j; ++k) m = m + (-Dinv[i][k]) * Cm[k][j];
      X[i][j] = m * a_inv[j];
    }
#pragma unroll
  for (int e = 0; e < 50; e += 2) {
    const double v0 = (e < 25) ? Dinv[e / 5][e % 5] : X[(e - 25) / 5][(e - 25) % 5];
    const double v1 = (e + 1 < 25) ? Dinv[(e + 1) / 5][(e + 1) % 5] : X[(e + 1 - 25) / 5][(e + 1 - 25) % 5];
    store2(rec + e, v0, v1);
  }
}
// row a of H from the head of the record; AROW = a when known at compile time is not needed: a is uniform per thread and the
// two cases (a < 5, a >= 5) are the two code paths of the one-thread version
template <int R>
TG_HD void setup_record_hrow_r(double* __restrict__ rec, int a) {
  constexpr int NQ = TG_N - R;
  const double a_inv[5] = {1.0 / 1.0, 1.0 / 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0};
  const double* __restrict__ Dinv = rec + TG_REC_DINV;  // [5][5]
  const double* __restrict__ X = rec + TG_REC_X;        // [5][5]
  const double* __restrict__ Qt = rec + TG_REC_Q;       // packed upper triangle, bitwise symmetric
  // column a of A^-1 below the diagonal block: X[k-5][a] (a < 5) or Dinv[k-5][a-5] (a >= 5), k = 5..9
  double col[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) col[k] = (a < 5) ? X[k * 5 + a] : Dinv[k * 5 + (a - 5)];
  const double ainv_a = (a < 5) ? a_inv[a < 5 ? a : 0] : 0.0;
  double W[NQ];
#pragma unroll
  for (int b = 0; b < NQ; ++b) {
    // rows a < 5 start with the k = a term when a >= R (Q rows below R are zero); every row then adds k = 5..9 in ascending order.
    // Written once for both halves: the first term present becomes the initial value, as in the one-thread version.
    bool have = false;
    double s = 0.0;
    if (a < 5 && a >= R) {
      s = ainv_a * Qt[tg_qsym(a - R, b)];
      have = true;
    }
#pragma unroll
    for (int k = 5; k < TG_N; ++k) {
      const double t = col[k - 5] * Qt[tg_qsym(k - R, b)];
      s = have ? s + t : t;
      have = true;
    }
    W[b] = s;
  }
  double hrow[TG_N];
#pragma unroll
  for (int b = 0; b < TG_N; ++b) {
    double s;
    if (b < 5) {
      bool have = false;
      s = 0.0;
      if (b >= R) { s = W[(b >= R) ? b - R : 0] * a_inv[b]; have = true; }
#pragma unroll
      for (int k = 5; k < TG_N; ++k) {
        const double t = W[k - R] * X[(k - 5) * 5 + b];
        s = have ? s + t : t;
        have = true;
      }
    } else {
      s = W[5 - R] * Dinv[0 * 5 + (b - 5)];
#pragma unroll
      for (int k = 6; k < TG_N; ++k) s = s + W[k - R] * Dinv[(k - 5) * 5 + (b - 5)];
    }
    hrow[b] = s;
  }
#pragma unroll
  for (int b = 0; b < TG_N; b += 2) store2(rec + TG_REC_H + a * TG_N + b, hrow[b], hrow[b + 1]);
}
TG_HD void setup_record_head(double T, int r, double* __restrict__ rec) {
  if (r == 2) setup_record_head_r<2>(T, rec);
  else if (r == 3) setup_record_head_r<3>(T, rec);
  else setup_record_head_r<4>(T, rec);
}
TG_HD void setup_record_hrow(int r, double* __restrict__ rec, int a) {
  if (r == 2) setup_record_hrow_r<2>(rec, a);
  else if (r == 3) setup_record_hrow_r<3>(rec, a);
  else setup_record_hrow_r<4>(rec, a);
}

TG_HD_NOINLINE void setup_segment_record(double T, int r, double* __restrict__ rec) {
  if (r == 2) setup_segment_record_r<2>(T, rec);
  else if (r == 3) setup_segment_record_r<3>(T, rec);
  else setup_segment_record_r<4>(T, rec);
}

}  // namespace tg

#endif  // TG_SEGMENT_CUH_
