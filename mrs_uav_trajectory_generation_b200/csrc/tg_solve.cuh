// tg_solve.cuh -- one warp solves one unconstrained QP instance (one problem at one segment-time vector).
// Replaces PolynomialOptimization<10>::constructR + solveLinear + updateSegmentsFromCompactConstraints +
// computeCost (reference: lin_impl.h:310-334, 340-373, 263-282, 127-141).
//
// The reduced matrix Rpp is block tridiagonal (a free derivative of vertex v couples only to v-1, v, v+1), so it
// is assembled directly into banded storage and factorised by LU without pivoting on the FULL non-symmetric
// band (why not Cholesky).  The work is written as a sequence of PHASES; inside a phase every lane
// owns disjoint outputs and reads only data written in earlier phases, so a __syncwarp() between phases is the
// only synchronisation.  tests/host_emu runs the same phases lane by lane on the CPU.
#ifndef TG_SOLVE_CUH_
#define TG_SOLVE_CUH_

#include "tg_common.cuh"

namespace tg {

struct SolveInst {
  int S;                  // segments
  int np;                 // free unknowns per dimension
  int hbw;                // half bandwidth of Rpp
  int variant;            // 0: base times; n >= 1: Mellinger perturbation of segment n-1 (nl_impl.h:282-311)
  int rec_stride;         // records per segment (1: base only, 3: base / +0.1 / -corr)
  int r;                  // derivative whose squared integral is minimised
  const uint8_t* vmask;   // [V] fixed-derivative bit masks
  const int* vfree;       // [V+1] index of the first free unknown of each vertex
  const double* vval;     // [V][5][4] fixed values
  const double* recs;     // [S*rec_stride][TG_REC_SIZE]
  double* coef_out;       // [S][4][10] or null
  double* cost_out;       // or null
  double* dp_out;         // [4][np] or null
  // workspace (shared or global memory)
  double* band;           // max(np*W, 40*S)
  double* rhs;            // 4*np
  double* xs;             // 4*np
  double* part;           // 4*S
};

TG_HD int solve_ws_doubles(int S, int np, int hbw) {
  const int W = 2 * hbw + 1;
  return imax(np * W, 40 * S) + 8 * np + 4 * S;
}
TG_HD void solve_ws_bind(SolveInst& I, double* ws) {
  const int W = 2 * I.hbw + 1;
  I.band = ws;
  I.rhs = ws + imax(I.np * W, 40 * I.S);
  I.xs = I.rhs + 4 * I.np;
  I.part = I.xs + 4 * I.np;
}

TG_HD const double* solve_rec(const SolveInst& I, int s) {
  const int which = (I.variant == 0) ? 0 : ((s == I.variant - 1) ? 1 : 2);
  return I.recs + (size_t)(s * I.rec_stride + (I.rec_stride == 1 ? 0 : which)) * TG_REC_SIZE;
}

// rank of slot a among the free slots of a vertex with mask m
TG_HD int free_rank(uint32_t m, int a) {
  const uint32_t below = (~m) & ((1u << a) - 1u) & 31u;
  return (int)((below & 1u) + ((below >> 1) & 1u) + ((below >> 2) & 1u) + ((below >> 3) & 1u) + ((below >> 4) & 1u));
}

// R entry between slot (v,a) and slot (w,b), |v-w| <= 1.  Segment v-1 contributes first (sum of two terms).
TG_HD double solve_R(const SolveInst& I, int v, int a, int w, int b) {
  if (w == v) {
    double s = 0.0;
    bool have = false;
    if (v > 0) { s = solve_rec(I, v - 1)[TG_REC_H + (TG_HALF + a) * TG_N + (TG_HALF + b)]; have = true; }
    if (v < I.S) {
      const double h = solve_rec(I, v)[TG_REC_H + a * TG_N + b];
      s = have ? s + h : h;
    }
    return s;
  }
  if (w == v + 1) return solve_rec(I, v)[TG_REC_H + a * TG_N + (TG_HALF + b)];
  return solve_rec(I, w)[TG_REC_H + (TG_HALF + a) * TG_N + b];
}

// phases: 0 zero band | 1 assemble | 2..2+np-1 LU steps | then np back-substitution steps | coefficients | cost partials | total
TG_HD int solve_num_phases(const SolveInst& I) { return (I.np > 0 ? 2 + 2 * I.np : 0) + 3; }

TG_HD void solve_phase(const SolveInst& I, int ph, int lane) {
  const int S = I.S, V = S + 1, np = I.np, hbw = I.hbw, W = 2 * hbw + 1;
  if (np > 0) {
    if (ph == 0) {
      for (int e = lane; e < np * W; e += 32) I.band[e] = 0.0;
      return;
    }
    if (ph == 1) {
      // one item per (vertex, slot); fixed slots have no row
      for (int it = lane; it < V * TG_HALF; it += 32) {
        const int v = it / TG_HALF, a = it - v * TG_HALF;
        const uint32_t mv = I.vmask[v];
        if ((mv >> a) & 1u) continue;
        const int i = I.vfree[v] + free_rank(mv, a);
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        const int w0 = imax(0, v - 1), w1 = imin(S, v + 1);
        for (int w = w0; w <= w1; ++w) {
          const uint32_t mw = I.vmask[w];
          for (int b = 0; b < TG_HALF; ++b) {
            const double rv = solve_R(I, v, a, w, b);
            if ((mw >> b) & 1u) {
              // rhs = (-Rpf) d_f over ascending fixed column (lin_impl.h:367)
              const double* f = I.vval + ((size_t)w * TG_HALF + b) * TG_D;
              const double nr = -rv;
              acc0 = acc0 + nr * f[0];
              acc1 = acc1 + nr * f[1];
              acc2 = acc2 + nr * f[2];
              acc3 = acc3 + nr * f[3];
            } else {
              const int j = I.vfree[w] + free_rank(mw, b);
              I.band[i * W + (j - i + hbw)] = rv;
            }
          }
        }
        I.rhs[0 * np + i] = acc0;
        I.rhs[1 * np + i] = acc1;
        I.rhs[2 * np + i] = acc2;
        I.rhs[3 * np + i] = acc3;
      }
      return;
    }
    if (ph < 2 + np) {
      // LU step k, right-looking, no pivoting.  lane -> (row group, column group)
      const int k = ph - 2;
      const int iend = imin(np - 1, k + hbw);
      const int ncol = (iend - k) + TG_D;  // band columns k+1..iend, then the 4 right-hand sides
      const double piv = I.band[k * W + hbw];
      for (int i = k + 1 + (lane >> 2); i <= iend; i += 8) {
        const double l = I.band[i * W + (k - i + hbw)] / piv;
        for (int c = (lane & 3); c < ncol; c += 4) {
          if (c < iend - k) {
            const int j = k + 1 + c;
            I.band[i * W + (j - i + hbw)] = I.band[i * W + (j - i + hbw)] - l * I.band[k * W + (j - k + hbw)];
          } else {
            const int d = c - (iend - k);
            I.rhs[d * np + i] = I.rhs[d * np + i] - l * I.rhs[d * np + k];
          }
        }
      }
      return;
    }
    if (ph < 2 + 2 * np) {
      // back substitution, column oriented: x_j = rhs_j / a_jj, then rhs_i -= a_ij x_j for the rows above.
      // Per row this subtracts the far columns first (descending j), the order of the contract.
      const int j = np - 1 - (ph - 2 - np);
      const int d = lane & 3;
      const double xj = I.rhs[d * np + j] / I.band[j * W + hbw];
      if ((lane >> 2) == 0) I.xs[d * np + j] = xj;
      const int i0 = imax(0, j - hbw);
      for (int i = j - 1 - (lane >> 2); i >= i0; i -= 8) I.rhs[d * np + i] = I.rhs[d * np + i] - I.band[i * W + (j - i + hbw)] * xj;
      return;
    }
    ph -= 2 + 2 * np;
  }
  if (ph == 0) {
    // coefficients c = A^-1 * [slots of vertex s ; slots of vertex s+1]  (lin_impl.h:271-280)
    for (int it = lane; it < S * TG_D * TG_N; it += 32) {
      const int s = it / (TG_D * TG_N), rem = it - s * (TG_D * TG_N), d = rem / TG_N, a = rem - d * TG_N;
      const double* rec = solve_rec(I, s);
      double nd[TG_N];
#pragma unroll
      for (int k = 0; k < TG_N; ++k) {
        const int v = s + (k >= TG_HALF ? 1 : 0), sl = k - (k >= TG_HALF ? TG_HALF : 0);
        const uint32_t m = I.vmask[v];
        nd[k] = ((m >> sl) & 1u) ? I.vval[((size_t)v * TG_HALF + sl) * TG_D + d] : I.xs[d * np + I.vfree[v] + free_rank(m, sl)];
      }
      double c;
      if (a < TG_HALF) {
        const double a_inv[5] = {1.0 / 1.0, 1.0 / 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0};
        double ai = a_inv[0];
#pragma unroll
        for (int q = 1; q < 5; ++q) ai = (a == q) ? a_inv[q] : ai;
        double ndv = nd[0];
#pragma unroll
        for (int q = 1; q < 5; ++q) ndv = (a == q) ? nd[q] : ndv;
        c = ai * ndv;
      } else {
        const double* xr = rec + TG_REC_X + (a - TG_HALF) * 5;
        const double* dr = rec + TG_REC_DINV + (a - TG_HALF) * 5;
        c = xr[0] * nd[0];
#pragma unroll
        for (int k = 1; k < 5; ++k) c = c + xr[k] * nd[k];
#pragma unroll
        for (int k = 0; k < 5; ++k) c = c + dr[k] * nd[5 + k];
      }
      I.band[it] = c;  // scratch for the cost (the band is dead now)
      if (I.coef_out) I.coef_out[it] = c;
    }
    if (I.dp_out)
      for (int e = lane; e < TG_D * np; e += 32) I.dp_out[e] = I.xs[e];
    return;
  }
  if (ph == 1) {
    // partial cost of (segment, dimension): (c^T Q) c over the non-zero block (lin_impl.h:135-137)
    if (!I.cost_out) return;
    const int r = I.r;
    for (int it = lane; it < S * TG_D; it += 32) {
      const int s = it / TG_D;
      const double* Q = solve_rec(I, s) + TG_REC_Q;
      const double* c = I.band + it * TG_N;
      const int nq = TG_N - r;
      double partial = 0.0;
      for (int b = 0; b < nq; ++b) {
        double sum = c[r] * Q[0 * 8 + b];
        for (int k = 1; k < nq; ++k) sum = sum + c[r + k] * Q[k * 8 + b];
        partial = (b == 0) ? sum * c[r + b] : partial + sum * c[r + b];
      }
      I.part[it] = partial;
    }
    return;
  }
  if (ph == 2) {
    if (lane == 0 && I.cost_out) {
      double total = 0.0;
      for (int it = 0; it < S * TG_D; ++it) total += I.part[it];
      *I.cost_out = 0.5 * total;
    }
    return;
  }
}

}  // namespace tg

#endif  // TG_SOLVE_CUH_
