// tg_solve_octet.cuh -- the reduced min-derivative solve with EIGHT LANES PER PROBLEM (four problems per warp).
// Same job and the same arithmetic, element for element, as solve_warp() in tg_solve.cuh (reference:
// lin_impl.h:310-334, 340-373, 263-282, 127-141); a different mapping onto the SM.
//
// Why: the round-1 profile of the warp-per-problem kernel (profiles/r01_solve_v2.md) showed 13.5 k warp instructions
// per solve, issue-bound, with 8 of 32 lanes useful in the factorisation (half bandwidth 7 => 7 rows change per
// elimination step).  Here the band's natural width IS the lane group:
//   * row i of the banded system lives in the REGISTERS of lane (i mod 8) of the octet while it is inside the
//     elimination window (steps i-7 .. i-1).  Column j is kept at register index (j - k0) mod 16, k0 = first step of
//     the current block of eight steps, so inside the eight-times unrolled block every register index is a
//     compile-time constant; between blocks the two register halves swap;
//   * the pivot row reaches the other seven lanes by warp SHUFFLES -- no shared-memory round trip and no
//     __syncwarp() inside the factorisation or the back substitution;
//   * shared memory holds the assembled rows until they enter the window, then (in place) the U rows for the back
//     substitution, which runs the same window upwards;
//   * coefficients and cost: the X / Dinv / Q blocks of a few segments at a time are staged into the (by then dead)
//     band storage with coalesced 16-byte loads and consumed from shared memory per (segment, dimension).
//
// Eligibility: half bandwidth exactly 7 (every interior vertex has position fixed and v, a, j, s free -- the node's
// recipe, node.cpp:931-977), at least 8 unknowns, and a workspace that fits shared memory; anything else takes
// solve_warp().  Written in the TG_PHASE style of tg_solve.cuh so that tests/host_emu can run it lane by lane: a
// shuffle becomes a read of the source lane's state in a phase of its own.
#ifndef TG_SOLVE_OCTET_CUH_
#define TG_SOLVE_OCTET_CUH_

#include "tg_solve.cuh"

namespace tg {

constexpr int kOctRow = 20;    // doubles per banded row: 16 column slots (column j at j & 15) + 4 right-hand sides
constexpr int kOctHbw = 7;
constexpr int kOctMinNp = 8;
constexpr int kOctStage = TG_REC_H;  // doubles staged per segment in phase 4: Dinv, X, Q (everything before H)
constexpr int kOctStageStride = 128;  // staging stride per segment (multiple of 16: static shared-memory offsets)

// shared-memory doubles for one octet solve: rows | partial costs | slot table | row -> (vertex, slot) table.
// Sized = 2 (mod 16) so that the four octets of a warp start in different banks.
TG_HD int octet_ws_doubles(int S, int np) {
  int n = np * kOctRow + 4 * S + (5 * (S + 1) + 3) / 4 + (np + 3) / 4;
  n = (n + 1) & ~1;
  while ((n & 15) != 2) n += 2;
  return n;
}
TG_HD void octet_ws_bind(SolveInst& I, double* ws) {
  I.W = kOctRow;
  I.rows = ws;
  I.xs = nullptr;
  I.part = ws + I.np * kOctRow;
  I.slot = (int16_t*)(I.part + 4 * I.S);
  I.rowva = I.slot + ((5 * (I.S + 1) + 3) / 4) * 4;
}
TG_HD bool octet_eligible(const SolveInst& I) { return I.hbw == kOctHbw && I.np >= kOctMinNp && I.dp_out == nullptr; }

struct alignas(16) Dbl2 {
  double x, y;
};

struct OctLane {
  double reg[16];  // band entries of the row this lane holds, column j at index (j - block start) & 15
  double rhs[4];
  double x[4];     // solution of the row, once final (back substitution)
  double rinv;
  int myrow;
  // values fetched from the lane that owns the pivot row (shuffle targets)
  double f_diag, f_u[7], f_r[4];
};

#if defined(__CUDA_ARCH__)
#define TG_OCT_LANES 1
#define TG_OCT_INST(insts, lane) (insts)[0]
#define TG_OCT_STATE(all, lane) (all)[0]
#define TG_OCT_FETCH(all, src, field) __shfl_sync(0xffffffffu, (all)[0].field, (src))
// straight-line section: no synchronisation (lanes exchange data by shuffles only)
#define TG_PHASE_NS(lane) for (int tg_once_ = 0; tg_once_ < 1; ++tg_once_)
#else
#define TG_OCT_LANES 32
#define TG_OCT_INST(insts, lane) (insts)[(lane) >> 3]
#define TG_OCT_STATE(all, lane) (all)[lane]
#define TG_OCT_FETCH(all, src, field) (all)[src].field
#define TG_PHASE_NS(lane) for (int lane = 0; lane < 32; ++lane)
#endif

// loads the 16 column slots and 4 right-hand sides of a stored row into the lane's registers; `flip` = 8 when the
// current block of steps starts at an odd multiple of eight (register index q holds stored slot q ^ flip)
TG_HD void octet_load_row(OctLane& st, const double* __restrict__ src, int flip) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const Dbl2 v = *reinterpret_cast<const Dbl2*>(src + ((2 * p) ^ flip));
    st.reg[2 * p] = v.x;
    st.reg[2 * p + 1] = v.y;
  }
  const Dbl2 r0 = *reinterpret_cast<const Dbl2*>(src + 16), r1 = *reinterpret_cast<const Dbl2*>(src + 18);
  st.rhs[0] = r0.x;
  st.rhs[1] = r0.y;
  st.rhs[2] = r1.x;
  st.rhs[3] = r1.y;
}
TG_HD void octet_swap_halves(OctLane& st) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double t = st.reg[q];
    st.reg[q] = st.reg[q + 8];
    st.reg[q + 8] = t;
  }
}

// index of staged double e inside the band storage: only the 16 column slots of every row are free, the
// right-hand-side slots hold the solution
TG_HD int octet_stage_index(int e) { return (e >> 4) * kOctRow + (e & 15); }

// (c^T Q) c over the non-zero block; `q` points at the staged record of the segment (offsets are compile-time)
template <int R>
TG_HD double octet_cost_partial(const double (&c)[TG_N], const double* __restrict__ q) {
  constexpr int nq = TG_N - R;
  double partial = 0.0;
#pragma unroll
  for (int b = 0; b < nq; ++b) {
    double sum = c[R] * q[octet_stage_index(TG_REC_Q + 0 * 8 + b)];
#pragma unroll
    for (int k = 1; k < nq; ++k) sum = sum + c[R + k] * q[octet_stage_index(TG_REC_Q + k * 8 + b)];
    partial = (b == 0) ? sum * c[R + b] : partial + sum * c[R + b];
  }
  return partial;
}

// Device: `insts` points at the calling lane's own instance (lanes of one octet hold identical copies); an octet
// without work has np == 0 and S == 0.  Host emulation: insts[4], one per octet.  nmax = max np over the warp.
TG_HD void solve_octets(const SolveInst* insts, int lane, int nmax) {
  OctLane st_all[TG_OCT_LANES];
  (void)lane;
  // ---- phase 0: slot tables, zeroed rows ---------------------------------------------------------------------
  TG_PHASE(lane) {
    const SolveInst& I = TG_OCT_INST(insts, lane);
    const int sub = lane & 7, V = I.S + 1;
    if (I.np > 0) {
      for (int it = sub; it < V * TG_HALF; it += 8) {
        const int v = it / TG_HALF, a = it - v * TG_HALF;
        const uint32_t m = I.vmask[v];
        const bool fixed = ((m >> a) & 1u) != 0;
        const int j = I.vfree[v] + free_rank(m, a);
        I.slot[it] = fixed ? (int16_t)-1 : (int16_t)j;
        if (!fixed) I.rowva[j] = (int16_t)it;
      }
      Dbl2 z;
      z.x = 0.0;
      z.y = 0.0;
      for (int e = sub; e < I.np * (kOctRow / 2); e += 8) *reinterpret_cast<Dbl2*>(I.rows + 2 * e) = z;
    }
  }
  // ---- phase 1: assemble Rpp and rhs = (-Rpf) d_f, one lane per row (identical sums to solve_warp phase 1) -------
  TG_PHASE(lane) {
    const SolveInst& I = TG_OCT_INST(insts, lane);
    const int sub = lane & 7, S = I.S;
    for (int i = sub; i < I.np; i += 8) {
      const int it = I.rowva[i];
      const int v = it / TG_HALF, a = it - v * TG_HALF;
      double hp[TG_N], hc[TG_N];
      const bool has_p = v > 0, has_c = v < S;
      if (has_p) {
        const double* src = solve_rec(I, v - 1) + TG_REC_H + (TG_HALF + a) * TG_N;  // 16-byte aligned: TG_REC_H and TG_N are even
#pragma unroll
        for (int q = 0; q < TG_N; q += 2) {
          const Dbl2 t = *reinterpret_cast<const Dbl2*>(src + q);
          hp[q] = t.x;
          hp[q + 1] = t.y;
        }
      }
      if (has_c) {
        const double* src = solve_rec(I, v) + TG_REC_H + a * TG_N;
#pragma unroll
        for (int q = 0; q < TG_N; q += 2) {
          const Dbl2 t = *reinterpret_cast<const Dbl2*>(src + q);
          hc[q] = t.x;
          hc[q + 1] = t.y;
        }
      }
      double* row = I.rows + i * kOctRow;
      double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int w = v - 1 + g;
        if (w < 0 || w > S) continue;
#pragma unroll
        for (int b = 0; b < TG_HALF; ++b) {
          double rv;
          if (g == 0) rv = hp[b];
          else if (g == 2) rv = hc[TG_HALF + b];
          else rv = has_p ? (has_c ? hp[TG_HALF + b] + hc[b] : hp[TG_HALF + b]) : hc[b];
          const int j = I.slot[w * TG_HALF + b];
          if (j >= 0) {
            row[j & 15] = rv;
          } else {
            const double* f = I.vval + ((size_t)w * TG_HALF + b) * TG_D;
            const double nr = -rv;
            acc0 = acc0 + nr * f[0];
            acc1 = acc1 + nr * f[1];
            acc2 = acc2 + nr * f[2];
            acc3 = acc3 + nr * f[3];
          }
        }
      }
      row[16] = acc0;
      row[17] = acc1;
      row[18] = acc2;
      row[19] = acc3;
    }
  }
  // ---- phase 2: LU without pivoting.  Lane sub holds row `myrow` (== sub mod 8) while it is in the window. -------------
  // (each lane reads back only rows it assembled itself, so no synchronisation is needed from here to phase 4)
  TG_PHASE_NS(lane) {
    const SolveInst& I = TG_OCT_INST(insts, lane);
    OctLane& st = TG_OCT_STATE(st_all, lane);
    st.myrow = lane & 7;
    st.rinv = 0.0;
#pragma unroll
    for (int d = 0; d < 4; ++d) st.x[d] = 0.0;
    if (st.myrow < I.np) {
      octet_load_row(st, I.rows + st.myrow * kOctRow, 0);
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) st.reg[q] = 0.0;
#pragma unroll
      for (int d = 0; d < 4; ++d) st.rhs[d] = 0.0;
    }
  }
  for (int k0 = 0; k0 < nmax; k0 += 8) {
    const int flip = k0 & 8;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = k0 + m;
      if (k < nmax) {
        TG_PHASE_NS(lane) {  // fetch the pivot row k from its owner
          OctLane& st = TG_OCT_STATE(st_all, lane);
          const int src = (lane & 24) | m;
          (void)src;
          const double diag = TG_OCT_FETCH(st_all, src, reg[m]);
          double u[7], ur[4];
#pragma unroll
          for (int c = 0; c < 7; ++c) u[c] = TG_OCT_FETCH(st_all, src, reg[(m + 1 + c) & 15]);
#pragma unroll
          for (int d = 0; d < 4; ++d) ur[d] = TG_OCT_FETCH(st_all, src, rhs[d]);
          st.f_diag = diag;
#pragma unroll
          for (int c = 0; c < 7; ++c) st.f_u[c] = u[c];
#pragma unroll
          for (int d = 0; d < 4; ++d) st.f_r[d] = ur[d];
        }
        TG_PHASE_NS(lane) {
          const SolveInst& I = TG_OCT_INST(insts, lane);
          OctLane& st = TG_OCT_STATE(st_all, lane);
          if (k < I.np) {
            const double rinv = 1.0 / st.f_diag;
            if (st.myrow == k) {
              // row k is final: keep its U part, right-hand sides and reciprocal pivot (in the one slot outside its
              // band) for the back substitution, then take row k+8 into the window
              double* rk = I.rows + k * kOctRow;
              rk[((m + 8) & 15) ^ flip] = rinv;
#pragma unroll
              for (int c = 0; c < 7; ++c) rk[((m + 1 + c) & 15) ^ flip] = st.reg[(m + 1 + c) & 15];
#pragma unroll
              for (int d = 0; d < 4; ++d) rk[16 + d] = st.rhs[d];
              st.myrow = k + 8;
              if (st.myrow < I.np) octet_load_row(st, I.rows + st.myrow * kOctRow, flip);
            } else if (st.myrow < I.np) {  // rows k+1 .. min(np-1, k+7)
              const double l = st.reg[m] * rinv;
#pragma unroll
              for (int c = 0; c < 7; ++c) st.reg[(m + 1 + c) & 15] = st.reg[(m + 1 + c) & 15] - l * st.f_u[c];
#pragma unroll
              for (int d = 0; d < 4; ++d) st.rhs[d] = st.rhs[d] - l * st.f_r[d];
            }
          }
        }
      }
    }
    TG_PHASE_NS(lane) { octet_swap_halves(TG_OCT_STATE(st_all, lane)); }
  }
  // ---- phase 3: back substitution, column oriented, the window moving upwards -----------------------------------
  const int jtop = (nmax > 0) ? ((nmax - 1) & ~7) : 0;
  TG_PHASE_NS(lane) {
    const SolveInst& I = TG_OCT_INST(insts, lane);
    OctLane& st = TG_OCT_STATE(st_all, lane);
    const int sub = lane & 7, np = I.np;
    st.myrow = (np > 0) ? (np - 1) - (((np - 1) - sub) & 7) : -1;
    if (st.myrow >= 0) {
      const double* src = I.rows + st.myrow * kOctRow;
      octet_load_row(st, src, jtop & 8);
      st.rinv = src[(st.myrow + 8) & 15];
      if (st.myrow == np - 1) {
#pragma unroll
        for (int d = 0; d < 4; ++d) st.x[d] = st.rhs[d] * st.rinv;
      }
    }
  }
  for (int j0 = jtop; j0 >= 0; j0 -= 8) {
    const int flip = j0 & 8;
#pragma unroll
    for (int mm = 0; mm < 8; ++mm) {
      const int m = 7 - mm;
      const int j = j0 + m;
      if (j < nmax) {
        TG_PHASE_NS(lane) {  // fetch x_j from the owner of row j
          OctLane& st = TG_OCT_STATE(st_all, lane);
          const int src = (lane & 24) | m;
          (void)src;
#pragma unroll
          for (int d = 0; d < 4; ++d) st.f_r[d] = TG_OCT_FETCH(st_all, src, x[d]);
        }
        TG_PHASE_NS(lane) {
          const SolveInst& I = TG_OCT_INST(insts, lane);
          OctLane& st = TG_OCT_STATE(st_all, lane);
          if (j < I.np) {
            if (st.myrow == j) {
              double* rj = I.rows + j * kOctRow;
#pragma unroll
              for (int d = 0; d < 4; ++d) rj[16 + d] = st.x[d];
              st.myrow = j - 8;
              if (st.myrow >= 0) {
                const double* src = I.rows + st.myrow * kOctRow;
                octet_load_row(st, src, flip);
                st.rinv = src[m ^ flip];  // row j-8 keeps its reciprocal pivot at stored slot (j-8+8) & 15 = j & 15
              }
            } else if (st.myrow >= 0 && st.myrow < j) {  // rows j-7 .. j-1
              const double a = st.reg[m];
#pragma unroll
              for (int d = 0; d < 4; ++d) st.rhs[d] = st.rhs[d] - a * st.f_r[d];
              if (st.myrow == j - 1) {
#pragma unroll
                for (int d = 0; d < 4; ++d) st.x[d] = st.rhs[d] * st.rinv;
              }
            }
          }
        }
      }
    }
    TG_PHASE_NS(lane) { octet_swap_halves(TG_OCT_STATE(st_all, lane)); }
  }
  // ---- phase 4: coefficients (lin_impl.h:271-280) and the partial costs (lin_impl.h:135-137) per (segment, dimension).
  // Rounds of up to four segments: stage Dinv | X | Q of the round's segments into the dead band storage, then
  // lane (dimension d = sub & 3, segment parity sub >> 2) computes its (segment, dimension) tasks from shared memory.
  TG_PHASE(lane) {}  // the solution (right-hand-side slots of every row) becomes visible to all lanes
  int nrounds = 0;
  {
    // rounds needed by the slowest octet of the warp (warp-uniform): every octet stages min(4, np*16/kOctStageStride) segments per round
    int worst = 1;
#if defined(__CUDA_ARCH__)
    const SolveInst& I0 = insts[0];
    const int per = (I0.np > 0) ? imin(4, (I0.np * 16) / kOctStageStride) : 4;
    const int need = (I0.S + per - 1) / per;
    worst = __reduce_max_sync(0xffffffffu, need);
#else
    for (int o = 0; o < 4; ++o) {
      const SolveInst& I0 = insts[o];
      const int per = (I0.np > 0) ? imin(4, (I0.np * 16) / kOctStageStride) : 4;
      worst = imax(worst, (I0.S + per - 1) / per);
    }
#endif
    nrounds = worst;
  }
  for (int round = 0; round < nrounds; ++round) {
    TG_PHASE(lane) {  // stage
      const SolveInst& I = TG_OCT_INST(insts, lane);
      const int sub = lane & 7;
      if (I.np > 0) {
        const int per = imin(4, (I.np * 16) / kOctStageStride);
        const int s_lo = round * per, s_hi = imin(I.S, s_lo + per);
        for (int s = s_lo; s < s_hi; ++s) {
          const double* rec = solve_rec(I, s);
          const int base = (s - s_lo) * kOctStageStride;
          for (int e = 2 * sub; e < kOctStage; e += 16) {
            const Dbl2 t = *reinterpret_cast<const Dbl2*>(rec + e);
            *reinterpret_cast<Dbl2*>(I.rows + octet_stage_index(base + e)) = t;
          }
        }
      }
    }
    TG_PHASE(lane) {  // compute
      const SolveInst& I = TG_OCT_INST(insts, lane);
      const int sub = lane & 7;
      if (I.np > 0) {
        const int per = imin(4, (I.np * 16) / kOctStageStride);
        const int s_lo = round * per, s_hi = imin(I.S, s_lo + per);
        for (int it = s_lo * TG_D + sub; it < s_hi * TG_D; it += 8) {
          const int s = it >> 2, d = it & 3;
          const int base = (s - s_lo) * kOctStageStride;
          double nd[TG_N], c[TG_N];
#pragma unroll
          for (int k = 0; k < TG_N; ++k) {
            const int j = I.slot[s * TG_HALF + k];  // slots of vertex s then vertex s+1 are contiguous in the table
            nd[k] = (j >= 0) ? I.rows[j * kOctRow + 16 + d] : I.vval[((size_t)s * TG_HALF + k) * TG_D + d];
          }
          c[0] = 1.0 * nd[0];
          c[1] = 1.0 * nd[1];
          c[2] = (1.0 / 2.0) * nd[2];
          c[3] = (1.0 / 6.0) * nd[3];
          c[4] = (1.0 / 24.0) * nd[4];
#pragma unroll
          for (int a = 0; a < TG_HALF; ++a) {
            double xr[5], dr[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              xr[k] = I.rows[octet_stage_index(base) + octet_stage_index(TG_REC_X + a * 5 + k)];
              dr[k] = I.rows[octet_stage_index(base) + octet_stage_index(TG_REC_DINV + a * 5 + k)];
            }
            double acc = xr[0] * nd[0];
#pragma unroll
            for (int k = 1; k < 5; ++k) acc = acc + xr[k] * nd[k];
#pragma unroll
            for (int k = 0; k < 5; ++k) acc = acc + dr[k] * nd[5 + k];
            c[TG_HALF + a] = acc;
          }
          if (I.coef_out) {
#pragma unroll
            for (int a = 0; a < TG_N; a += 2) {
              Dbl2 t;
              t.x = c[a];
              t.y = c[a + 1];
              *reinterpret_cast<Dbl2*>(I.coef_out + it * TG_N + a) = t;
            }
          }
          if (I.cost_out) {
            const double* q = I.rows + octet_stage_index(base);
            I.part[it] = (I.r == 2) ? octet_cost_partial<2>(c, q) : ((I.r == 3) ? octet_cost_partial<3>(c, q) : octet_cost_partial<4>(c, q));
          }
        }
      }
    }
  }
  // ---- phase 5: total in (segment, dimension) order (lin_impl.h:131-140) ----------------------------------------------
  TG_PHASE(lane) {
    const SolveInst& I = TG_OCT_INST(insts, lane);
    if ((lane & 7) == 0 && I.cost_out && I.S > 0) {
      double total = 0.0;
      for (int it = 0; it < I.S * TG_D; ++it) total += I.part[it];
      *I.cost_out = 0.5 * total;
    }
  }
}

}  // namespace tg

#endif  // TG_SOLVE_OCTET_CUH_
