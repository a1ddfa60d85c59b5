// tg_solve_thread.cuh -- the reduced min-derivative solve with ONE THREAD PER INSTANCE.
// Same job and the same arithmetic, element for element, as solve_warp() (tg_solve.cuh) and solve_octets()
// (tg_solve_octet.cuh) (reference: lin_impl.h:310-334 constructR, 340-373 solveLinear); a different mapping onto the SM.
//
// Why (round-2 profile, profiles/r02_solve_thread.md): the eight-lanes-per-problem kernel issues ~4500 warp instructions per
// solve -- shuffles of the pivot row, a division replicated in every lane, predication of the owner / other-lanes roles --
// for ~12 k useful flops: 7 % of the FP64 pipe.  A Mellinger evaluation has (S+1) x B independent instances (7 x 10^5 at
// bench size): there is no need to split one of them over lanes.  Here a thread owns an instance end to end; nothing is
// exchanged between lanes, every register index is a compile-time constant, and the FP64 instructions are the arithmetic.
//
// The reduced matrix is block tridiagonal by vertex: the free derivatives of vertex v couple to those of v-1, v, v+1 only.
// Every vertex block is padded to FOUR unknowns (a padded unknown is the equation 1 * x = 0: its multipliers are exact
// zeros, so every real element receives the same updates a - l * u, in the same pivot order, as in the scalar band LU of
// the numeric contract; structural zeros are skipped, which is exact as well).  The window in registers is
//   D  (4x4) diagonal block of vertex v        U  (4x4) coupling (v, v+1)        b  (4x4) right-hand sides of vertex v
//   L  (4x4) coupling (v+1, v)                 Dn (4x4) diagonal block of v+1    bn (4x4) right-hand sides of vertex v+1
// A finished row (1/pivot, 3 + 4 upper entries, 4 right-hand sides = 12 doubles) goes to a slab in global memory,
// interleaved by thread so that a warp's accesses coalesce, and comes back for the back substitution.
//
// Eligibility: every vertex has at most four free derivatives (always true for the node's recipe: position is fixed at
// every waypoint, node.cpp:931-977) -- SolveInst::fmax <= 4; other problems take the older kernels.
#ifndef TG_SOLVE_THREAD_CUH_
#define TG_SOLVE_THREAD_CUH_

#include "tg_solve.cuh"

namespace tg {

constexpr int kThrB = 4;      // padded unknowns per vertex
constexpr int kThrRow = 12;   // doubles of a finished row: 1/pivot, D[q][q+1..3] (3 slots), U[q][0..3], b[q][0..3]

TG_HD bool thread_eligible(const SolveInst& I) { return I.fmax <= kThrB && I.np > 0 && I.dp_out == nullptr && I.x_out != nullptr; }
// rows of slab one instance needs
TG_HD int thread_slab_rows(int S) { return kThrB * (S + 1); }

// software prefetch (device only): the H block of a record (800 bytes) / one slab element, into L1
TG_HD void thr_prefetch_H(const double* __restrict__ rec) {
#if defined(__CUDA_ARCH__)
  const char* p = reinterpret_cast<const char*>(rec + TG_REC_H);
#pragma unroll
  for (int o = 0; o < 800 + 127; o += 128) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p + o));
#else
  (void)rec;
#endif
}
// a slab element read for the last time (back substitution): evict-first, so that it does not push out rows still waiting
TG_HD double thr_last_use(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldcs(p);
#else
  return *p;
#endif
}
TG_HD void thr_prefetch(const double* __restrict__ p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p));
#else
  (void)p;
#endif
}

struct ThrVertex {
  int f;        // real free unknowns (<= 4)
  int a[kThrB]; // derivative index of free unknown q (q < f)
  uint32_t m;   // fixed mask
};
TG_HD ThrVertex thr_vertex(uint32_t m) {
  ThrVertex t;
  t.m = m & 31u;
  t.f = 0;
#pragma unroll
  for (int q = 0; q < kThrB; ++q) t.a[q] = 0;
#pragma unroll
  for (int a = 0; a < TG_HALF; ++a) {
    if (!((m >> a) & 1u)) {
#pragma unroll
      for (int q = 0; q < kThrB; ++q)
        if (q == t.f) t.a[q] = a;
      t.f += 1;
    }
  }
  return t;
}

// R entry between slot (v, a) and slot (w, b), |v - w| <= 1; the contribution of segment v-1 is added first (lin_impl.h:317-333)
TG_HD double thr_rentry_diag(const double* __restrict__ Hp, const double* __restrict__ Hc, int a, int b) {
  if (Hp && Hc) return Hp[(TG_HALF + a) * TG_N + (TG_HALF + b)] + Hc[a * TG_N + b];
  if (Hp) return Hp[(TG_HALF + a) * TG_N + (TG_HALF + b)];
  return Hc[a * TG_N + b];
}

// right-hand sides of the rows of vertex v: rhs = (-Rpf) d_f accumulated over ascending fixed column (vertex v-1, v, v+1)
TG_HD void thr_rhs(const SolveInst& I, int v, const ThrVertex& tv, const double* __restrict__ Hp, const double* __restrict__ Hc, double (&out)[kThrB][TG_D]) {
#pragma unroll
  for (int q = 0; q < kThrB; ++q)
#pragma unroll
    for (int d = 0; d < TG_D; ++d) out[q][d] = 0.0;
  for (int g = 0; g < 3; ++g) {
    const int w = v - 1 + g;
    if (w < 0 || w > I.S) continue;
    const uint32_t mw = I.vmask[w];
    for (int bb = 0; bb < TG_HALF; ++bb) {
      if (!((mw >> bb) & 1u)) continue;
      const double* f = I.vval + ((size_t)w * TG_HALF + bb) * TG_D;
      const double f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3];
#pragma unroll
      for (int q = 0; q < kThrB; ++q) {
        if (q < tv.f) {
          const int a = tv.a[q];
          double rv;
          if (g == 0) rv = Hp[(TG_HALF + a) * TG_N + bb];
          else if (g == 2) rv = Hc[a * TG_N + (TG_HALF + bb)];
          else rv = thr_rentry_diag(Hp, Hc, a, bb);
          const double nr = -rv;
          out[q][0] = out[q][0] + nr * f0;
          out[q][1] = out[q][1] + nr * f1;
          out[q][2] = out[q][2] + nr * f2;
          out[q][3] = out[q][3] + nr * f3;
        }
      }
    }
  }
}

// diagonal block of vertex v (identity on the padding)
TG_HD void thr_diag(const ThrVertex& tv, const double* __restrict__ Hp, const double* __restrict__ Hc, double (&D)[kThrB][kThrB]) {
#pragma unroll
  for (int i = 0; i < kThrB; ++i)
#pragma unroll
    for (int j = 0; j < kThrB; ++j) {
      if (i < tv.f && j < tv.f) D[i][j] = thr_rentry_diag(Hp, Hc, tv.a[i], tv.a[j]);
      else D[i][j] = (i == j) ? 1.0 : 0.0;
    }
}

// slab: element e of row r of this thread at slab[(r * kThrRow + e) * estride]
TG_HD void solve_thread(const SolveInst& I, double* __restrict__ slab, size_t estride) {
  const int S = I.S, V = S + 1;
  double D[kThrB][kThrB], U[kThrB][kThrB], L[kThrB][kThrB], Dn[kThrB][kThrB], b[kThrB][TG_D], bn[kThrB][TG_D];
  thr_prefetch_H(solve_rec(I, 0));
  ThrVertex tv = thr_vertex(I.vmask[0]);
  if (S > 1) thr_prefetch_H(solve_rec(I, 1));
  {
    const double* Hc = solve_rec(I, 0) + TG_REC_H;
    thr_diag(tv, nullptr, Hc, D);
    thr_rhs(I, 0, tv, nullptr, Hc, b);
  }
  // ---- forward elimination, vertex by vertex ----------------------------------------------------------------------------
  for (int v = 0; v < V; ++v) {
    const bool has_next = v < S;
    ThrVertex tn = tv;
    if (v + 2 < S) thr_prefetch_H(solve_rec(I, v + 2));  // needed one step from now (Hn of step v+1)
    if (has_next && v + 2 <= S && tv.m == 1u && (I.vmask[v + 1] & 31u) == 1u && (I.vmask[v + 2] & 31u) == 1u) {
      // the common interior step: vertices v, v+1, v+2 fix the position only (free derivatives 1..4 at static indices)
      tn = tv;
      const double* Hv = solve_rec(I, v) + TG_REC_H;
      const double* Hn = solve_rec(I, v + 1) + TG_REC_H;
#pragma unroll
      for (int i = 0; i < kThrB; ++i)
#pragma unroll
        for (int j = 0; j < kThrB; ++j) {
          U[i][j] = Hv[(i + 1) * TG_N + (TG_HALF + j + 1)];
          L[i][j] = Hv[(TG_HALF + i + 1) * TG_N + (j + 1)];
          Dn[i][j] = Hv[(TG_HALF + i + 1) * TG_N + (TG_HALF + j + 1)] + Hn[(i + 1) * TG_N + (j + 1)];
        }
      const double* f0 = I.vval + (size_t)v * TG_HALF * TG_D;         // fixed position of vertex v, v+1, v+2
      const double* f1 = f0 + TG_HALF * TG_D;
      const double* f2 = f1 + TG_HALF * TG_D;
#pragma unroll
      for (int i = 0; i < kThrB; ++i) {
        const double r0 = -Hv[(TG_HALF + i + 1) * TG_N];
        const double r1 = -(Hv[(TG_HALF + i + 1) * TG_N + TG_HALF] + Hn[(i + 1) * TG_N]);
        const double r2 = -Hn[(i + 1) * TG_N + TG_HALF];
#pragma unroll
        for (int d = 0; d < TG_D; ++d) {
          double acc = 0.0;
          acc = acc + r0 * f0[d];
          acc = acc + r1 * f1[d];
          acc = acc + r2 * f2[d];
          bn[i][d] = acc;
        }
      }
    } else if (has_next) {
      tn = thr_vertex(I.vmask[v + 1]);
      const double* Hv = solve_rec(I, v) + TG_REC_H;                                   // segment v: couples v and v+1
      const double* Hn = (v + 1 < S) ? solve_rec(I, v + 1) + TG_REC_H : nullptr;        // segment v+1
#pragma unroll
      for (int i = 0; i < kThrB; ++i)
#pragma unroll
        for (int j = 0; j < kThrB; ++j) {
          U[i][j] = (i < tv.f && j < tn.f) ? Hv[tv.a[i] * TG_N + (TG_HALF + tn.a[j])] : 0.0;
          L[i][j] = (i < tn.f && j < tv.f) ? Hv[(TG_HALF + tn.a[i]) * TG_N + tv.a[j]] : 0.0;
        }
      thr_diag(tn, Hv, Hn, Dn);
      thr_rhs(I, v + 1, tn, Hv, Hn, bn);
    }
#pragma unroll
    for (int q = 0; q < kThrB; ++q) {
      const double rinv = 1.0 / D[q][q];
#pragma unroll
      for (int i = q + 1; i < kThrB; ++i) {
        const double l = D[i][q] * rinv;
#pragma unroll
        for (int j = q + 1; j < kThrB; ++j) D[i][j] = D[i][j] - l * D[q][j];
        if (has_next) {
#pragma unroll
          for (int j = 0; j < kThrB; ++j) U[i][j] = U[i][j] - l * U[q][j];
        }
#pragma unroll
        for (int d = 0; d < TG_D; ++d) b[i][d] = b[i][d] - l * b[q][d];
      }
      if (has_next) {
#pragma unroll
        for (int i = 0; i < kThrB; ++i) {
          const double l = L[i][q] * rinv;
#pragma unroll
          for (int j = q + 1; j < kThrB; ++j) L[i][j] = L[i][j] - l * D[q][j];
#pragma unroll
          for (int j = 0; j < kThrB; ++j) Dn[i][j] = Dn[i][j] - l * U[q][j];
#pragma unroll
          for (int d = 0; d < TG_D; ++d) bn[i][d] = bn[i][d] - l * b[q][d];
        }
      }
      // row (v, q) is final
      double* row = slab + (size_t)((v * kThrB + q) * kThrRow) * estride;
      row[0] = rinv;
#pragma unroll
      for (int j = 1; j < kThrB; ++j)
        if (j > q) row[(size_t)j * estride] = D[q][j];  // the back substitution never reads the entries left of the diagonal ...
      if (has_next)                                     // ... nor the coupling block of the last vertex
#pragma unroll
        for (int j = 0; j < kThrB; ++j) row[(size_t)(4 + j) * estride] = U[q][j];
#pragma unroll
      for (int d = 0; d < TG_D; ++d) row[(size_t)(8 + d) * estride] = b[q][d];
    }
    if (has_next) {
#pragma unroll
      for (int i = 0; i < kThrB; ++i) {
#pragma unroll
        for (int j = 0; j < kThrB; ++j) D[i][j] = Dn[i][j];
#pragma unroll
        for (int d = 0; d < TG_D; ++d) b[i][d] = bn[i][d];
      }
      tv = tn;
    }
  }
  // ---- back substitution: far columns first (vertex v+1's unknowns 3..0, then this vertex's 3..q+1), x = s * (1/pivot) -----
  double xn[kThrB][TG_D], x[kThrB][TG_D];
#pragma unroll
  for (int q = 0; q < kThrB; ++q)
#pragma unroll
    for (int d = 0; d < TG_D; ++d) {
      xn[q][d] = 0.0;
      x[q][d] = 0.0;
    }
  for (int v = V - 1; v >= 0; --v) {
    const bool has_next = v < S;
    if (v > 1) {  // the rows of vertex v-2 start travelling towards L1 (two vertices ahead of their use)
      const double* nx = slab + (size_t)((v - 2) * kThrB * kThrRow) * estride;
#pragma unroll
      for (int e = 0; e < kThrB * kThrRow; ++e) thr_prefetch(nx + (size_t)e * estride);
    }
#pragma unroll
    for (int qq = 0; qq < kThrB; ++qq) {
      const int q = kThrB - 1 - qq;
      const double* row = slab + (size_t)((v * kThrB + q) * kThrRow) * estride;
      const double rinv = thr_last_use(row);
      double s[TG_D];
#pragma unroll
      for (int d = 0; d < TG_D; ++d) s[d] = thr_last_use(row + (size_t)(8 + d) * estride);
      if (has_next) {
#pragma unroll
        for (int jj = 0; jj < kThrB; ++jj) {
          const int j = kThrB - 1 - jj;
          const double u = thr_last_use(row + (size_t)(4 + j) * estride);
#pragma unroll
          for (int d = 0; d < TG_D; ++d) s[d] = s[d] - u * xn[j][d];
        }
      }
#pragma unroll
      for (int jj = 0; jj < kThrB; ++jj) {
        const int j = kThrB - 1 - jj;
        if (j > q) {
          const double u = thr_last_use(row + (size_t)j * estride);
#pragma unroll
          for (int d = 0; d < TG_D; ++d) s[d] = s[d] - u * x[j][d];
        }
      }
#pragma unroll
      for (int d = 0; d < TG_D; ++d) x[q][d] = s[d] * rinv;
    }
    // the real unknowns of vertex v leave through global memory (CoefCostFn reads them)
    const uint32_t m = I.vmask[v];
    const int f = TG_HALF - (int)((m & 1u) + ((m >> 1) & 1u) + ((m >> 2) & 1u) + ((m >> 3) & 1u) + ((m >> 4) & 1u));
    const int j0 = I.vfree[v];
#pragma unroll
    for (int q = 0; q < kThrB; ++q) {
      if (q < f) {
#pragma unroll
        for (int d = 0; d < TG_D; d += 2) store2(I.x_out + (size_t)(j0 + q) * 4 + d, x[q][d], x[q][d + 1]);
      }
#pragma unroll
      for (int d = 0; d < TG_D; ++d) xn[q][d] = x[q][d];
    }
  }
}

}  // namespace tg

#endif  // TG_SOLVE_THREAD_CUH_
