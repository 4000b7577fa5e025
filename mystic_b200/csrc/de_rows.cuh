// K2d: several candidate rows per warp for short rows (D a multiple of 4, 16 <= D <= 128),
// built to stay under the instruction budget of an HBM-bound generation (a 64-dimensional
// Rand1Exp row moves ~1.26 KB, so a B200 SM has ~340 issue slots per row at 60 % of HBM).
//
// A group of LPR lanes (4 for D <= 64, 8 for D <= 128) owns a row; a warp works on 32/LPR rows
// at once, so every per-row step (draws, window bookkeeping, selection) is one SIMT instruction
// stream for 4-8 rows.  Per chunk of rows:
//   1. the parent rows go global -> shared with 16-byte cp.async copies (consecutive lanes of a
//      group copy consecutive 16-byte pieces, so every 32-byte sector is requested once) into the
//      warp's tile; no registers are held while they are in flight;
//   2. meanwhile the draws of a row are spread over the lanes of its group: ONE Philox4x32-10
//      evaluation per lane yields the donor blocks (lanes 0..NB-1) and the first crossover blocks
//      (other lanes) at once; the exponential crossover length is found by ballot over the group
//      and further rounds of LPR blocks are only computed while some row of the warp needs them
//      (same counters and words as philox.cuh / oracle/de_oracle.py: results are identical);
//   3. only the 4-element chunks inside the crossover window are visited: lane `sub` of the group
//      takes the window's chunks sub, sub+LPR, ...; donor segments are 16-byte vector loads along
//      D; the trial is built IN PLACE in the tile and the original chunk goes to an undo tile;
//   4. cost and penalty read the tile: lane `sub` accumulates the chunks q = sub, sub+LPR, ... in
//      order and the group combines the LPR partial sums by a xor tree (corana: lane j evaluates
//      term j and the four terms are summed in the reference's order);
//   5. selection streams the tile to the next generation; a rejected trial takes the window's
//      chunks from the undo tile instead.
// The kernel serves every mode (step, generation 0, mb2_eval_cost) for its row lengths, so all
// energies of a context come from one summation order.
//
// Reference semantics: mystic/strategy.py:34-311 (trial), mystic/symbolic.py:1139-1144 and
// mystic/tools.py:420-426 (bounds), mystic/differential_evolution.py:603-613 (selection).
#pragma once
#include "de_kernels.cuh"

namespace mb2 {

#ifndef MB2_ROWS_MINBLOCKS
#define MB2_ROWS_MINBLOCKS 2
#endif
#ifndef MB2_ROWS_CPT
#define MB2_ROWS_CPT 1
#endif
#ifndef MB2_ROWS_WARPS
#define MB2_ROWS_WARPS 8
#endif
constexpr int kRowsWarps = MB2_ROWS_WARPS;
constexpr int kRowsTrips = 4;     // 4-element chunks per lane: D <= 16 * LPR
constexpr int kRowsMinDim = 16;
constexpr int kRowsMaxDim = 128;

template <int LPR, bool PROD>
__device__ __forceinline__ double rows_reduce(double v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = PROD ? dmul(v, w) : dadd(v, w);
  }
  return v;
}

// c / s for a constant s with r = RN(1/s): Markstein refinement gives the correctly rounded
// quotient (cf. div_by_table); operands whose residuals could over- or underflow take the IEEE divide
__device__ __forceinline__ double div_by_const(double c, double s, double r) {
  const double q0 = __dmul_rn(c, r);
  const double q1 = __fma_rn(__fma_rn(-q0, s, c), r, q0);
  const double q2 = __fma_rn(__fma_rn(-q1, s, c), r, q1);
  const double a = fabs(c);
  if (!(a < 1e300) || a < 1e-280) return c / s;
  return q2;
}

// cost of the row in the trial tile `xs`; every lane of the group returns the same value
template <int COST, int LPR>
__device__ __forceinline__ double rows_cost(const double* xs, int NQ, int sub, int gbase, const CostParams& cp) {
  if (COST == COST_ROSEN) {
    // dejong.py:98-101
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < kRowsTrips; ++t) {
      const int q = t * LPR + sub;
      if (q < NQ) {
        const double2 a = *reinterpret_cast<const double2*>(xs + 4 * q);
        const double2 b = *reinterpret_cast<const double2*>(xs + 4 * q + 2);
        const bool more = q + 1 < NQ;
        const double x[5] = {a.x, a.y, b.x, b.y, more ? xs[4 * q + 4] : 0.0};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (e < 3 || more) {
            const double tt = dsub(x[e + 1], dmul(x[e], x[e]));
            const double u = dsub(1.0, x[e]);
            acc = dadd(acc, dadd(dmul(100.0, dmul(tt, tt)), dmul(u, u)));
          }
        }
      }
    }
    return rows_reduce<LPR, false>(acc);
  } else if (COST == COST_CORANA) {
    // storn.py:77-96 (D >= 4 here): lane j of the group evaluates term j
    const int j = sub & 3;
    const double dj = (j == 0) ? 1. : ((j == 1) ? 1000. : ((j == 2) ? 10. : 100.));
    const double xj = xs[j];
    const double sg = (xj > 0.0) ? 1.0 : ((xj < 0.0) ? -1.0 : 0.0);  // numpy.sign
    const double zj = dmul(dmul(floor(dadd(fabs(div_by_const(xj, 0.2, 5.0)), 0.49999)), sg), 0.2);
    double term;
    if (fabs(dsub(xj, zj)) < 0.05) {
      const double sz = (zj > 0.0) ? 1.0 : ((zj < 0.0) ? -1.0 : 0.0);
      const double w = dsub(zj, dmul(0.05, sz));
      term = dmul(dmul(0.15, dmul(w, w)), dj);
    } else {
      term = dmul(dmul(dj, xj), xj);
    }
    double r = 0.0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) r = dadd(r, __shfl_sync(0xffffffffu, term, gbase + jj));
    return r;
  } else {  // COST_GRIEWANGK, storn.py:135-139
    double s = 0.0, p = 1.0;
#pragma unroll
    for (int t = 0; t < kRowsTrips; ++t) {
      const int q = t * LPR + sub;
      if (q < NQ) {
        const double2 a = *reinterpret_cast<const double2*>(xs + 4 * q);
        const double2 b = *reinterpret_cast<const double2*>(xs + 4 * q + 2);
        const double x[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s = dadd(s, dmul(x[e], x[e]));
          const double2 tb = __ldg(cp.grw_tab + 4 * q + e);
          p = dmul(p, grw_factor(x[e], tb.x, tb.y, cp.grw_bounded != 0));
        }
      }
    }
    s = rows_reduce<LPR, false>(s);
    p = rows_reduce<LPR, true>(p);
    return dadd(dsub(s / 4000.0, p), 1.0);
  }
}

template <int LPR>
__device__ __forceinline__ double rows_penalty(const double* xs, int D, int NQ, int sub, const PenaltyParams& pp) {
  double total = 0.0;
  for (int t = 0; t < pp.n; ++t) {
    const double* g = pp.G + (int64_t)t * D;
    double acc = 0.0;
#pragma unroll
    for (int tr = 0; tr < kRowsTrips; ++tr) {
      const int q = tr * LPR + sub;
      if (q < NQ) {
        const double2 a = *reinterpret_cast<const double2*>(xs + 4 * q);
        const double2 b = *reinterpret_cast<const double2*>(xs + 4 * q + 2);
        const double2 ga = __ldg(reinterpret_cast<const double2*>(g + 4 * q));
        const double2 gb = __ldg(reinterpret_cast<const double2*>(g + 4 * q + 2));
        acc = dadd(acc, dmul(ga.x, a.x));
        acc = dadd(acc, dmul(ga.y, a.y));
        acc = dadd(acc, dmul(gb.x, b.x));
        acc = dadd(acc, dmul(gb.y, b.y));
      }
    }
    const double pf = dsub(rows_reduce<LPR, false>(acc), __ldg(pp.h + t));
    const double v = penalty_term(__ldg(pp.ptype + t), __ldg(pp.kscale + t), __ldg(pp.mult + t), pf);
    total = dadd(v, total);
  }
  return total;
}

// the draws of one candidate (strategy.py:27,42,48-56)
template <int K>
struct RowDraws {
  int r[K];      // donor rows (the kernel is only used for shards of < 2^31 rows)
  int nstart;    // first mutated dimension
  int L;         // exponential crossover length
};

// Warp-collective: every lane of the warp calls it for the row of its group.
template <int LPR, int K>
__device__ __forceinline__ void rows_draw(const StepParams& P, int xover, int64_t cc, int sub, int gbase, RowDraws<K>& d) {
  constexpr int NB = (K + 2) / 2;  // Philox blocks of the donor stream (K donors + start index)
  const int D = P.dim;
  const uint32_t cand_g = (uint32_t)(P.global_offset + cc);
  d.nstart = 0;
  d.L = 0;
  if (P.replay) {
#pragma unroll
    for (int j = 0; j < K; ++j) d.r[j] = __ldg(P.r_donors + cc * 5 + j);
    d.nstart = __ldg(P.r_nstart + cc);
    if (xover == XOVER_EXP) d.L = exp_length_thread(P, 0u, cc, D);
    return;
  }
  // lane `sub` of the row's group evaluates block `sub` of the donor stream (NB <= 3 <= LPR blocks: donors,
  // start index and -- in its low half -- the geometric draw of the crossover length); ONE Philox evaluation
  // per lane makes all the draws of the row
  u32x4 v = philox4x32_10((uint32_t)(sub < NB ? sub : 0), cand_g, P.generation, STREAM_DONORS, P.key0, P.key1);
  const uint64_t even = ((uint64_t)v.x << 32) | v.y, odd = ((uint64_t)v.z << 32) | v.w;
  uint64_t rw[2 * NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    rw[2 * b] = __shfl_sync(0xffffffffu, even, gbase + b);
    rw[2 * b + 1] = __shfl_sync(0xffffffffu, odd, gbase + b);
  }
  int64_t r64[K];
  uint32_t xword;
  resolve_donors<K>(rw, cc, P.np, D, r64, &d.nstart, &xword);
#pragma unroll
  for (int j = 0; j < K; ++j) d.r[j] = (int)r64[j];
  if (xover == XOVER_EXP) d.L = exp_length_v2(xword, P.xlen, D);
}

// the K donor segments of one 4-element chunk
template <int K>
struct DonorChunk {
  double2 v[K][2];
};

template <int K>
__device__ __forceinline__ void rows_load_donors(const double* __restrict__ pop, const RowDraws<K>& d, int D, int i0,
                                                 DonorChunk<K>& dc) {
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double* dr = pop + (int64_t)d.r[j] * D + i0;
    dc.v[j][0] = ld_stream2(dr);
    dc.v[j][1] = ld_stream2(dr + 2);
  }
}

// mutate the crossed elements of chunk q of the trial tile, apply the bounds to them; the parent's
// chunk goes to the 4 doubles at `undo_chunk` (nullptr: the caller restores a rejected trial itself)
template <int K>
__device__ __forceinline__ bool rows_mutate_chunk(const StepParams& P, int base, int xover, const RowDraws<K>& d, int64_t cc,
                                                  int q, const DonorChunk<K>& dc, const double* s_best, double* t_trial, double* undo_chunk) {
  const int D = P.dim;
  const int i0 = q << 2;
  bool cross[4];
  if (xover == XOVER_BIN) {
    // strategy.py:79-85: i == n or random() < CR, one draw per dimension
    if (P.replay) {
      const double* u = P.r_uni + cc * (int64_t)(D + 1) + i0;
#pragma unroll
      for (int e = 0; e < 4; ++e) cross[e] = (i0 + e == d.nstart) || (__ldg(u + e) < P.CR);
    } else {
      // block q/2 decides the 8 dimensions of chunks 2*(q/2) and 2*(q/2)+1: the even chunk reads words x, y
      const u32x4 wv = philox4x32_10((uint32_t)(q >> 1), (uint32_t)(P.global_offset + cc), P.generation, STREAM_XOVER, P.key0, P.key1);
      const uint32_t wa = (q & 1) ? wv.z : wv.x, wb = (q & 1) ? wv.w : wv.y;
      cross[0] = (i0 == d.nstart) || field_lo_hit(wa, P.thr16);
      cross[1] = (i0 + 1 == d.nstart) || field_hi_hit(wa, P.thr16);
      cross[2] = (i0 + 2 == d.nstart) || field_lo_hit(wb, P.thr16);
      cross[3] = (i0 + 3 == d.nstart) || field_hi_hit(wb, P.thr16);
    }
  } else {
    int rel = i0 - d.nstart;
    if (rel < 0) rel += D;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int re = rel + e;
      if (re >= D) re -= D;
      cross[e] = re < d.L;
    }
  }
  const double2 xa = *reinterpret_cast<const double2*>(t_trial + i0);
  const double2 xb = *reinterpret_cast<const double2*>(t_trial + i0 + 2);
  const double x[4] = {xa.x, xa.y, xb.x, xb.y};
  if (undo_chunk != nullptr) {  // the parent's chunk, for a rejected trial
    *reinterpret_cast<double2*>(undo_chunk) = xa;
    *reinterpret_cast<double2*>(undo_chunk + 2) = xb;
  }
  double bst[4] = {0.0, 0.0, 0.0, 0.0};
  if (K != 3 && K != 5) {  // Best1, RandToBest1, Best2 read bestSolution
    const double2 ba = *reinterpret_cast<const double2*>(s_best + i0);
    const double2 bb = *reinterpret_cast<const double2*>(s_best + i0 + 2);
    bst[0] = ba.x;
    bst[1] = ba.y;
    bst[2] = bb.x;
    bst[3] = bb.y;
  }
  double y[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    double pe[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < K; ++j) pe[j] = (e == 0) ? dc.v[j][0].x : ((e == 1) ? dc.v[j][0].y : ((e == 2) ? dc.v[j][1].x : dc.v[j][1].y));
    if (K == 2) {
      y[e] = (base == BASE_BEST1) ? mutant<BASE_BEST1>(x[e], bst[e], pe, P.F) : mutant<BASE_RANDTOBEST1>(x[e], bst[e], pe, P.F);
    } else if (K == 3) {
      y[e] = mutant<BASE_RAND1>(x[e], bst[e], pe, P.F);
    } else if (K == 4) {
      y[e] = mutant<BASE_BEST2>(x[e], bst[e], pe, P.F);
    } else {
      y[e] = mutant<BASE_RAND2>(x[e], bst[e], pe, P.F);
    }
  }
  bool oob = false;
  if (P.bounds_mode != BOUNDS_NONE) {
    double l[4], h[4];
    if (P.bounds_uniform) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        l[e] = P.lo_s;
        h[e] = P.hi_s;
      }
    } else {
      const double2 la = __ldg(reinterpret_cast<const double2*>(P.lo + i0)), lb = __ldg(reinterpret_cast<const double2*>(P.lo + i0 + 2));
      const double2 ha = __ldg(reinterpret_cast<const double2*>(P.hi + i0)), hb = __ldg(reinterpret_cast<const double2*>(P.hi + i0 + 2));
      l[0] = la.x, l[1] = la.y, l[2] = lb.x, l[3] = lb.y;
      h[0] = ha.x, h[1] = ha.y, h[2] = hb.x, h[3] = hb.y;
    }
    if (P.bounds_mode == BOUNDS_CLAMP) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        y[e] = (y[e] > l[e]) ? y[e] : l[e];  // python max(lo, x)
        y[e] = (y[e] < h[e]) ? y[e] : h[e];  // python min(hi, x)
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) oob |= cross[e] && ((y[e] < l[e]) || (y[e] > h[e]));
    }
  }
  *reinterpret_cast<double2*>(t_trial + i0) = make_double2(cross[0] ? y[0] : x[0], cross[1] ? y[1] : x[1]);
  *reinterpret_cast<double2*>(t_trial + i0 + 2) = make_double2(cross[2] ? y[2] : x[2], cross[3] ? y[3] : x[3]);
  return oob;
}

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// a parent row global -> shared without passing through registers: lane `sub` copies the
// 16-byte pieces sub, sub+LPR, ...
template <int LPR>
__device__ __forceinline__ void rows_fetch_parent(const double* __restrict__ prow, int D, int sub, double* tile) {
#pragma unroll
  for (int t = 0; t < 2 * kRowsTrips; ++t) {
    const int i = 2 * (t * LPR + sub);
    if (i < D) cp_async16(tile + i, prow + i);
  }
  cp_async_commit();
}

template <int COST, int LPR, int K>
__global__ void __launch_bounds__(kRowsWarps * 32, MB2_ROWS_MINBLOCKS)
    de_step_rows_kernel(const __grid_constant__ StepParams P, int base, int xover, int /*ndonors*/) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(16) double smem[];
  __shared__ double sh_e[kRowsWarps];
  __shared__ int64_t sh_i[kRowsWarps];
  __shared__ unsigned long long sh_ninf[kRowsWarps], sh_nacc[kRowsWarps];
  constexpr int RPW = 32 / LPR;  // rows per warp
  constexpr unsigned LMASK = (1u << LPR) - 1u;
  constexpr int CPT = (K <= 3) ? MB2_ROWS_CPT : 1;  // window chunks per lane whose donor loads are issued together
  static_assert(LPR >= 4 && LPR <= 16 && (K + 2) / 2 < LPR, "group too narrow for the draws");

  const int D = P.dim;
  const int NQ = D >> 2;
  const int TS = D + 2;  // tile row stride: 16 bytes of padding keep the 16-byte accesses of a quarter warp on distinct banks
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const int grp = lane / LPR;
  const int sub = lane % LPR;
  const int gbase = grp * LPR;
  const bool uni_b = P.bounds_uniform != 0;
  const bool step = P.mode == MODE_STEP;
  const bool needs_best = (base == BASE_BEST1 || base == BASE_RANDTOBEST1 || base == BASE_BEST2);
  // the untouched elements of a trial need no bounds handling when generation 0 clipped the rows
  const bool check_own = P.bounds_mode != BOUNDS_NONE && !(step && P.rows_in_range);
  double* s_best = smem;                                                        // [D]
  double* t_row = smem + D + (size_t)warp * (2 * RPW * TS) + (size_t)grp * TS;  // this row: parent, then trial in place
  double* t_undo = t_row + RPW * TS;                                             // parent chunks the trial replaced

  if (step && needs_best) {
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_best[i] = __ldg(P.best_x + i);
  }
  __syncthreads();

  double w_best_e = CUDART_INF;
  int64_t w_best_idx = INT64_MAX;
  unsigned long long w_ninf = 0, w_nacc = 0;

  const int64_t nchunks = (P.np + RPW - 1) / RPW;
  const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
  const int64_t nw = (int64_t)gridDim.x * wpc;
  for (int64_t ch = gw; ch < nchunks; ch += nw) {
    const int64_t c = ch * RPW + grp;
    const bool live = c < P.np;
    const int64_t cc = live ? c : (P.np - 1);  // idle groups shadow the last row and discard the result

    // ---- 1. parent row -> tile (asynchronous), 2. draws
    rows_fetch_parent<LPR>(P.pop_cur + cc * D, D, sub, t_row);
    double e_old = CUDART_INF;
    RowDraws<K> dr;
    int q0 = 0, nq = 0;
    if (step) {
      e_old = __ldg(P.e_cur + cc);
      rows_draw<LPR, K>(P, xover, cc, sub, gbase, dr);
      nq = NQ;
      if (xover == XOVER_EXP) {
        q0 = dr.nstart >> 2;
        nq = (dr.L > 0) ? (((dr.nstart + dr.L - 1) >> 2) - q0 + 1) : 0;
        if (nq > NQ) nq = NQ;  // a full-length window that starts inside a chunk: each chunk once
      }
    }

    // ---- 3. crossover window: chunks q0 .. q0+nq-1 (mod NQ); the donor loads of the lane's first
    // CPT chunks are issued before the parent row is waited for
    DonorChunk<K> dc[CPT];
    int qs[CPT];
    if (step) {
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        int q = q0 + sub + u * LPR;
        if (q >= NQ) q -= NQ;
        qs[u] = q;
        if (sub + u * LPR < nq) rows_load_donors<K>(P.pop_cur, dr, D, q << 2, dc[u]);
      }
    }
    cp_async_wait_all();
    __syncwarp();

    bool oob = false;
    if (check_own) {
      // generation 0: clip into the strict range (differential_evolution.py:516-520); evaluation:
      // range test; a step whose parents were not clipped by this context: the whole trial is
      // clamped / tested like the reference does, and the raw parent is kept in the undo tile
#pragma unroll
      for (int t = 0; t < kRowsTrips; ++t) {
        const int q = t * LPR + sub;
        if (q < NQ) {
          const double2 pa = *reinterpret_cast<const double2*>(t_row + 4 * q);
          const double2 pb = *reinterpret_cast<const double2*>(t_row + 4 * q + 2);
          double x[4] = {pa.x, pa.y, pb.x, pb.y};
          if (step) {
            *reinterpret_cast<double2*>(t_undo + 4 * q) = pa;
            *reinterpret_cast<double2*>(t_undo + 4 * q + 2) = pb;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const double l = uni_b ? P.lo_s : __ldg(P.lo + 4 * q + e);
            const double h = uni_b ? P.hi_s : __ldg(P.hi + 4 * q + e);
            if (P.mode == MODE_GEN0) {
              x[e] = fmin(fmax(x[e], l), h);
              oob |= (x[e] < l) || (x[e] > h);
            } else if (step && P.bounds_mode == BOUNDS_CLAMP) {
              x[e] = (x[e] > l) ? x[e] : l;  // python max(lo, x)
              x[e] = (x[e] < h) ? x[e] : h;  // python min(hi, x)
            } else {
              oob |= (x[e] < l) || (x[e] > h);
            }
          }
          *reinterpret_cast<double2*>(t_row + 4 * q) = make_double2(x[0], x[1]);
          *reinterpret_cast<double2*>(t_row + 4 * q + 2) = make_double2(x[2], x[3]);
        }
      }
      __syncwarp();
    }

    if (step) {
#pragma unroll
      for (int u = 0; u < CPT; ++u) {
        if (sub + u * LPR < nq) oob |= rows_mutate_chunk<K>(P, base, xover, dr, cc, qs[u], dc[u], s_best, t_row, t_undo + (qs[u] << 2));
      }
      for (int w = sub + CPT * LPR; w < nq; w += LPR) {  // long windows: the remaining chunks
        int q = q0 + w;
        if (q >= NQ) q -= NQ;
        DonorChunk<K> dx;
        rows_load_donors<K>(P.pop_cur, dr, D, q << 2, dx);
        oob |= rows_mutate_chunk<K>(P, base, xover, dr, cc, q, dx, s_best, t_row, t_undo + (q << 2));
      }
    }
    // any lane of the group out of range?
    oob = ((__ballot_sync(0xffffffffu, oob) >> gbase) & LMASK) != 0;
    __syncwarp();

    // ---- 4. energy = (inf if out of range else cost(x)) + penalty(x); idle groups evaluate their
    // shadow row because the shuffles inside need every lane
    double E = CUDART_INF;
    {
      const double cost = rows_cost<COST, LPR>(t_row, NQ, sub, gbase, P.cost);
      if (!oob) E = cost;
      E = dadd(E, rows_penalty<LPR>(t_row, D, NQ, sub, P.pen));
    }

    if (live && P.cap_trial != nullptr && P.mode != MODE_EVAL) {
      for (int i = sub; i < D; i += LPR) P.cap_trial[c * D + i] = t_row[i];
      if (sub == 0) P.cap_energy[c] = E;
    }

    // ---- 5. selection (differential_evolution.py:603-613) and the next-generation row
    if (P.mode == MODE_EVAL) {
      if (live && sub == 0) P.e_next[c] = E;
    } else if (live) {
      const bool acc = E < e_old;
      const bool keep = acc || P.mode == MODE_GEN0;
      double* __restrict__ nrow = P.pop_next + c * D;
      const bool store = keep || !P.sparse;  // sparse: a rejected row is not written
#pragma unroll
      for (int t = 0; t < 2 * kRowsTrips; ++t) {
        const int i = 2 * (t * LPR + sub);  // 16-byte piece of chunk i/4
        if (store && i < D) {
          int rq = (i >> 2) - q0;
          if (rq < 0) rq += NQ;
          // a rejected trial: chunks the window (or the clamp of the whole row) replaced come from the undo tile
          const bool undo = !keep && (check_own || rq < nq);
          const double2 a = *reinterpret_cast<const double2*>((undo ? t_undo : t_row) + i);
          st_stream2(nrow + i, a.x, a.y);
        }
      }
      if (sub == 0) {
        P.e_next[c] = acc ? E : e_old;
        if (acc) {
          ++w_nacc;
          if (E < w_best_e || (E == w_best_e && c < w_best_idx)) {
            w_best_e = E;
            w_best_idx = c;
          }
        }
        if (isinf(E)) ++w_ninf;
      }
    }
    __syncwarp();  // the tiles are reused by the next chunk
  }

  if (P.mode == MODE_EVAL) return;

  // warp argmin over the group leaders (other lanes carry +inf / 0), lowest index on ties
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oe = __shfl_xor_sync(0xffffffffu, w_best_e, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, w_best_idx, o);
    if (oe < w_best_e || (oe == w_best_e && oi < w_best_idx)) {
      w_best_e = oe;
      w_best_idx = oi;
    }
    w_ninf += __shfl_xor_sync(0xffffffffu, w_ninf, o);
    w_nacc += __shfl_xor_sync(0xffffffffu, w_nacc, o);
  }
  if (lane == 0) {
    sh_e[warp] = w_best_e;
    sh_i[warp] = w_best_idx;
    sh_ninf[warp] = w_ninf;
    sh_nacc[warp] = w_nacc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double be = sh_e[0];
    int64_t bi = sh_i[0];
    unsigned long long ni = sh_ninf[0], na = sh_nacc[0];
    for (int w = 1; w < wpc; ++w) {
      if (sh_e[w] < be || (sh_e[w] == be && sh_i[w] < bi)) {
        be = sh_e[w];
        bi = sh_i[w];
      }
      ni += sh_ninf[w];
      na += sh_nacc[w];
    }
    P.part_e[blockIdx.x] = be;
    P.part_idx[blockIdx.x] = bi;
    P.part_ninf[blockIdx.x] = ni;
    P.part_nacc[blockIdx.x] = na;
  }
}

}  // namespace mb2
