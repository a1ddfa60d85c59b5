// tg_solve_octet.cuh -- the reduced min-derivative solve with EIGHT LANES PER PROBLEM (four problems per warp).
// Same job and the same arithmetic, element for element, as solve_warp() in tg_solve.cuh (reference:
// lin_impl.h:310-334, 340-373, 263-282, 127-141); a different mapping onto the SM.
//
// Why: the round-1 profile of the warp-per-problem kernel (ncu in session 1 of round 1, figures in DESIGN.md 4.1) showed
// 13.5 k warp instructions per solve, issue-bound, with 8 of 32 lanes useful in the factorisation (half bandwidth 7 => 7 rows change per
// elimination step).  Here the band's natural width IS the lane group:
//   * row i of the banded system lives in the REGISTERS of lane (i mod 8) of the octet while it is inside the
//     elimination window (steps i-7 .. i-1).  Column j is kept at register index (j - k0) mod 16, k0 = first step of
//     the current block of eight steps, so inside the eight-times unrolled block every register index is a
//     compile-time constant; between blocks the two register halves swap;
//   * the pivot row reaches the other seven lanes by warp SHUFFLES -- no shared-memory round trip and no
//     __syncwarp() inside the factorisation or the back substitution;
//   * the rows STREAM through shared memory: a ring of two blocks of eight rows holds the block that is about to enter
//     the window (assembled) and the H rows of the block after it (cp.async in flight); finished rows (1/pivot, U part,
//     right-hand sides: 96 bytes) wait for the back substitution in a per-warp slab of global memory that stays in L2 and
//     return through the same ring.  Shared memory per problem is 2.9 KB whatever its size (see solve_octets);
//   * the routine ends with the solution of the reduced system written to global memory; coefficients and cost are a
//     separate flat kernel (CoefCostFn, one thread per (segment, dimension)) -- fused into this kernel they cost a third
//     of its time at 8 warps per SM (measured in session 2 of round 1).
//
// Eligibility: half bandwidth exactly 7 (every interior vertex has position fixed and v, a, j, s free -- the node's
// recipe, node.cpp:931-977) and at least 8 unknowns; anything else takes solve_warp() in a second launch (k_solve).  Written in the TG_PHASE style of tg_solve.cuh so that tests/host_emu can run it lane by lane: a
// shuffle becomes a read of the source lane's state in a phase of its own.
#ifndef TG_SOLVE_OCTET_CUH_
#define TG_SOLVE_OCTET_CUH_

#include "tg_solve.cuh"

namespace tg {

constexpr int kOctRow = 20;    // doubles per banded row: 16 column slots (column j at j & 15) + 4 right-hand sides
constexpr int kOctHbw = 7;
constexpr int kOctMinNp = 8;
constexpr int kOctRing = 16;   // rows of the shared-memory ring: two blocks of eight (row i at ((i >> 3) & 1) * 8 + (i & 7))
constexpr int kOctStride = 22;  // doubles between ring rows: 176 B, so that the eight lanes of an octet hit eight different 16-byte bank groups
constexpr int kOctURow = 12;   // doubles of a finished row kept for the back substitution: 1/pivot, U[i][i+1..i+7], 4 right-hand sides

// Shared-memory doubles for one octet solve: row ring | slot table | row -> (vertex, slot) table.  Independent of the
// problem size but for the two small tables: the rows themselves stream through the ring (assembled one block ahead of
// the elimination) and the U rows wait for the back substitution in a per-warp slab of global memory that stays in L2.
// Sized = 2 (mod 16) so that the four octets of a warp start in different banks.
TG_HD int octet_ws_doubles(int S, int np) {
  int n = kOctRing * kOctStride + (5 * (S + 1) + 3) / 4 + (np + 3) / 4;
  n = (n + 1) & ~1;
  while ((n & 15) != 2) n += 2;
  return n;
}
// urows: np * kOctURow doubles of global memory owned by this octet
TG_HD void octet_ws_bind(SolveInst& I, double* ws, double* urows) {
  I.W = kOctRow;
  I.rows = ws;
  I.xs = urows;
  I.part = nullptr;
  I.slot = (int16_t*)(ws + kOctRing * kOctStride);
  I.rowva = I.slot + ((5 * (I.S + 1) + 3) / 4) * 4;
}
TG_HD int octet_ring_pos(int i) { return ((i >> 3) & 1) * 8 + (i & 7); }
// the routine produces the solution of the reduced system only (x_out); coefficients and cost are CoefCostFn's job
TG_HD bool octet_eligible(const SolveInst& I) { return I.hbw == kOctHbw && I.np >= kOctMinNp && I.dp_out == nullptr && I.x_out != nullptr; }


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

// asynchronous copy of one 10-entry row of H (global memory) into shared memory: five 16-byte cp.async on the device,
// a plain copy in the host emulation
TG_HD void octet_stage_row(double* dst, const double* __restrict__ src) {
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
  for (int q = 0; q < TG_N / 2; ++q) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + 16u * q), "l"(src + 2 * q) : "memory");
#else
  for (int q = 0; q < TG_N; ++q) dst[q] = src[q];
#endif
}
// the four fixed values of one (vertex, derivative) are pulled into L1 ahead of their use (staging them in shared
// memory costs one resident warp per SM at S = 10, measured slower: profiles/r01_solve_octet_s3.md)
TG_HD void octet_prefetch_fixed(const double* __restrict__ src) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];\n" ::"l"(src));
#else
  (void)src;
#endif
}
TG_HD void octet_stage_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// One lane's part of a block of eight rows.  stage: the two rows of H that row i of R is built from -- row (5+a) of
// H_{v-1} into slots 0..9, row a of H_v into slots 10..19 of the row's ring position -- as asynchronous 16-byte copies.
TG_HD void octet_stage_block(const SolveInst& I, int sub, int blk) {
  const int i = blk * 8 + sub;
  if (i >= I.np) return;
  const int it = I.rowva[i];
  const int v = it / TG_HALF, a = it - v * TG_HALF;
  double* row = I.rows + octet_ring_pos(i) * kOctStride;
  if (v > 0) octet_stage_row(row, solve_rec(I, v - 1) + TG_REC_H + (TG_HALF + a) * TG_N);  // 16-byte aligned: TG_REC_H and TG_N are even
  if (v < I.S) octet_stage_row(row + TG_N, solve_rec(I, v) + TG_REC_H + a * TG_N);
}
// assemble: row i of Rpp and of rhs = (-Rpf) d_f in place, from the staged H rows (identical sums to solve_warp phase 1)
TG_HD void octet_assemble_block(const SolveInst& I, int sub, int blk) {
  const int i = blk * 8 + sub, S = I.S;
  if (i >= I.np) return;
  const int it = I.rowva[i];
  const int v = it / TG_HALF;
  double hp[TG_N], hc[TG_N];
  const bool has_p = v > 0, has_c = v < S;
  double* row = I.rows + octet_ring_pos(i) * kOctStride;
#pragma unroll
  for (int q = 0; q < TG_N; q += 2) {
    const Dbl2 t = *reinterpret_cast<const Dbl2*>(row + q), u = *reinterpret_cast<const Dbl2*>(row + TG_N + q);
    hp[q] = t.x;
    hp[q + 1] = t.y;
    hc[q] = u.x;
    hc[q + 1] = u.y;
  }
  Dbl2 z;
  z.x = 0.0;
  z.y = 0.0;
#pragma unroll
  for (int q = 0; q < kOctRow; q += 2) *reinterpret_cast<Dbl2*>(row + q) = z;
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
// back substitution: the stored U row i (global memory, L2) travels back into its ring position (first kOctURow doubles)
TG_HD void octet_stage_urow(const SolveInst& I, int i) {
  if (i < 0 || i >= I.np) return;
  double* dst = I.rows + octet_ring_pos(i) * kOctStride;
  const double* src = I.xs + (size_t)i * kOctURow;
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
  for (int q = 0; q < kOctURow / 2; ++q) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + 16u * q), "l"(src + 2 * q) : "memory");
#else
  for (int q = 0; q < kOctURow; ++q) dst[q] = src[q];
#endif
}
// A U row in the ring -> the lane's registers.  rel = (row index) - (first column of the current block of steps): column
// i+1+c sits at register index (rel + 1 + c) & 15.  Called from fully unrolled code with a literal rel, so that every
// register index is static after inlining.
TG_HD void octet_load_urow_static(OctLane& st, const double* __restrict__ src, const int REL) {
  const Dbl2 a = *reinterpret_cast<const Dbl2*>(src), b = *reinterpret_cast<const Dbl2*>(src + 2), c = *reinterpret_cast<const Dbl2*>(src + 4),
             d = *reinterpret_cast<const Dbl2*>(src + 6), e = *reinterpret_cast<const Dbl2*>(src + 8), f = *reinterpret_cast<const Dbl2*>(src + 10);
  st.rinv = a.x;
  st.reg[(REL + 1) & 15] = a.y;
  st.reg[(REL + 2) & 15] = b.x;
  st.reg[(REL + 3) & 15] = b.y;
  st.reg[(REL + 4) & 15] = c.x;
  st.reg[(REL + 5) & 15] = c.y;
  st.reg[(REL + 6) & 15] = d.x;
  st.reg[(REL + 7) & 15] = d.y;
  st.rhs[0] = e.x;
  st.rhs[1] = e.y;
  st.rhs[2] = f.x;
  st.rhs[3] = f.y;
}
// the same with rel known only at run time (start of the back substitution): register index q static, source index computed
TG_HD void octet_load_urow_dyn(OctLane& st, const double* __restrict__ src, int rel) {
  st.rinv = src[0];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int c = (q - rel - 1) & 15;  // column offset held by register q
    st.reg[q] = (c < 7) ? src[1 + c] : 0.0;
  }
#pragma unroll
  for (int d = 0; d < 4; ++d) st.rhs[d] = src[8 + d];
}

// Device: `insts` points at the calling lane's own instance (lanes of one octet hold identical copies); an octet
// without work has np == 0 and S == 0.  Host emulation: insts[4], one per octet.  nmax = max np over the warp.
//
// Memory plan (profiles/r01_solve_octet_s3.md): the shared-memory ring holds two blocks of eight rows.  While the
// elimination works on block b (its rows are in registers), block b+1 sits assembled in the ring -- each lane takes its
// next row from there when the one it holds becomes final -- and the H rows of block b+2 are in flight (cp.async) into
// the positions block b has just left.  A final row goes to the octet's slab in global memory (U part, right-hand sides,
// reciprocal pivot: it stays in L2) and comes back the same way, one block ahead, for the back substitution.  Every ring
// position is written and read by one lane only (row i <-> lane i mod 8), so the whole routine needs no __syncwarp()
// after the slot tables; shared memory per problem no longer depends on its size.
TG_HD void solve_octets(const SolveInst* insts, int lane, int nmax) {
  OctLane st_all[TG_OCT_LANES];
  (void)lane;
  // ---- phase 0: slot tables ----------------------------------------------------------------------------------------
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
        else octet_prefetch_fixed(I.vval + (size_t)it * TG_D);  // read during assembly by the rows that have this column
      }
    }
  }
  // ---- phase 1: the H rows of block 0 start flowing into the ring -------------------------------------------------------
  TG_PHASE_NS(lane) { octet_stage_block(TG_OCT_INST(insts, lane), lane & 7, 0); }
  // ---- phase 2: LU without pivoting.  Lane sub holds row `myrow` (== sub mod 8) while it is in the window. -------------
  // The loop starts one block early (k0 = -8): that pass only assembles block 0 and takes it into registers, so that the
  // assembly code exists once (instruction cache: profiles/r01_solve_octet_s3.md).
  for (int k0 = -8; k0 < nmax; k0 += 8) {
    const int flip = k0 & 8, blk = k0 >> 3;  // blk = -1 in the first pass
    TG_PHASE_NS(lane) {
      // block blk+1 has landed (its copies were issued one block ago): assemble it.  The rows of block blk are all in
      // registers by now, so the H rows of block blk+2 may start flowing into the ring positions they came from.
      const SolveInst& I = TG_OCT_INST(insts, lane);
      const int sub = lane & 7;
      octet_stage_wait();
      octet_assemble_block(I, sub, blk + 1);
      octet_stage_block(I, sub, blk + 2);
      if (blk < 0) {
        OctLane& st = TG_OCT_STATE(st_all, lane);
        st.myrow = sub;
        st.rinv = 0.0;
#pragma unroll
        for (int d = 0; d < 4; ++d) st.x[d] = 0.0;
        if (st.myrow < I.np) {
          octet_load_row(st, I.rows + octet_ring_pos(st.myrow) * kOctStride, 0);
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) st.reg[q] = 0.0;
#pragma unroll
          for (int d = 0; d < 4; ++d) st.rhs[d] = 0.0;
        }
      }
    }
    if (blk < 0) continue;
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
              // row k is final: its U part, right-hand sides and reciprocal pivot (in the one slot outside its band) go
              // to the slab for the back substitution, then row k+8 comes out of the ring into the window
              double* rk = I.xs + (size_t)k * kOctURow;
              store2(rk, rinv, st.reg[(m + 1) & 15]);
              store2(rk + 2, st.reg[(m + 2) & 15], st.reg[(m + 3) & 15]);
              store2(rk + 4, st.reg[(m + 4) & 15], st.reg[(m + 5) & 15]);
              store2(rk + 6, st.reg[(m + 6) & 15], st.reg[(m + 7) & 15]);
              store2(rk + 8, st.rhs[0], st.rhs[1]);
              store2(rk + 10, st.rhs[2], st.rhs[3]);
              st.myrow = k + 8;
              if (st.myrow < I.np) octet_load_row(st, I.rows + octet_ring_pos(st.myrow) * kOctStride, flip);
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
  // Each lane starts with the highest row congruent to its index (rows np-8 .. np-1 span the octet's top two blocks) and
  // afterwards takes row j-8 whenever its row j is done; U rows come back from the slab through the ring one block ahead.
  const int jtop = (nmax > 0) ? ((nmax - 1) & ~7) : 0;
  TG_PHASE_NS(lane) {
    const SolveInst& I = TG_OCT_INST(insts, lane);
    OctLane& st = TG_OCT_STATE(st_all, lane);
    const int sub = lane & 7, np = I.np;
    octet_stage_wait();  // nothing of the elimination is still in flight
    st.myrow = (np > 0) ? (np - 1) - (((np - 1) - sub) & 7) : -1;
    if (st.myrow >= 0) {
      const int btop = (np - 1) >> 3;
      octet_stage_urow(I, btop * 8 + sub);
      octet_stage_urow(I, (btop - 1) * 8 + sub);
      octet_stage_wait();
      octet_load_urow_dyn(st, I.rows + octet_ring_pos(st.myrow) * kOctStride, st.myrow - jtop);
      if (st.myrow == np - 1) {
#pragma unroll
        for (int d = 0; d < 4; ++d) st.x[d] = st.rhs[d] * st.rinv;
      }
    }
  }
  for (int j0 = jtop; j0 >= 0; j0 -= 8) {
    const int blk = j0 >> 3;
    TG_PHASE_NS(lane) {
      // rows of block blk are in registers (or this octet ends below it); block blk-1 must have landed before the steps
      // below take rows from it, and block blk-2 may now use the ring positions of block blk
      const SolveInst& I = TG_OCT_INST(insts, lane);
      const int sub = lane & 7;
      octet_stage_wait();
      // (the octet's own top block brought block blk-1 along in the prologue; every lane has taken its row of block blk
      // out of the ring by now, whichever of the two top blocks it started in)
      if (I.np > 0 && blk <= ((I.np - 1) >> 3)) octet_stage_urow(I, (blk - 2) * 8 + sub);
    }
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
#pragma unroll
              for (int d = 0; d < 4; ++d) I.x_out[j * 4 + d] = st.x[d];  // the solution leaves through global memory
              st.myrow = j - 8;
              if (st.myrow >= 0) octet_load_urow_static(st, I.rows + octet_ring_pos(st.myrow) * kOctStride, m - 8);  // row j-8: rel = m - 8
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
}

}  // namespace tg

#endif  // TG_SOLVE_OCTET_CUH_
