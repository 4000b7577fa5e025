// K2d: several candidate rows per warp for short rows (D a multiple of 4, 16 <= D <= 128),
// built to stay under the instruction budget of an HBM-bound generation (a 64-dimensional
// Rand1Exp row moves ~1.26 KB, so a B200 SM has ~340 issue slots per row at 60 % of HBM).
//
// A group of LPR lanes (4 for D <= 64, 8 for D <= 128) owns a row; a warp works on 32/LPR rows
// at once, so every per-row step (draws, window bookkeeping, selection) is one SIMT instruction
// stream for 4-8 rows.  Per chunk of rows:
//   1. the parent rows are read with 16-byte streaming loads (lane `sub` owns the 4-element
//      chunks q = sub, sub+LPR, ...) and written to two shared-memory tiles of the warp: the
//      parent tile (kept for rejected trials) and the trial tile;
//   2. the draws of a row are spread over the lanes of its group: ONE Philox4x32-10 evaluation
//      per lane yields the donor blocks (lanes 0..NB-1) and the first crossover blocks (other
//      lanes) at once; the exponential crossover length is found by ballot over the group and
//      further rounds of LPR blocks are only computed while some row of the warp needs them
//      (same counters and words as philox.cuh / oracle/de_oracle.py: results are identical);
//   3. only the chunks inside the crossover window are visited: lane `sub` of the group takes the
//      window's chunks sub, sub+LPR, ...; donor segments are 16-byte vector loads along D; the
//      mutated and clamped chunk goes back to the trial tile;
//   4. cost and penalty read the trial tile: lane `sub` accumulates its own elements in order and
//      the group combines the LPR partial sums by a xor tree (corana: lane j evaluates term j and
//      the four terms are summed in the reference's order);
//   5. selection picks the tile (trial or parent) that is streamed to the next generation.
// The kernel serves every mode (step, generation 0, mb2_eval_cost) for its row lengths, so all
// energies of a context come from one summation order.
//
// Reference semantics: mystic/strategy.py:34-311 (trial), mystic/symbolic.py:1139-1144 and
// mystic/tools.py:420-426 (bounds), mystic/differential_evolution.py:603-613 (selection).
#pragma once
#include "de_kernels.cuh"

namespace mb2 {

constexpr int kRowsWarps = 8;
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
          p = dmul(p, cos(div_by_table(x[e], tb.x, tb.y)));
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
    const double k = __ldg(pp.kscale + t);
    const int ty = __ldg(pp.ptype + t);
    const double m = (pf > 0.0) ? pf : 0.0;
    double v;
    if (ty == 0) v = dmul(k, dmul(pf, pf));
    else if (ty == 1) v = dmul(k, fabs(pf));
    else if (ty == 2) v = dmul(2.0 * k, dmul(m, m));
    else v = dmul(2.0 * k, fabs(m));
    total = dadd(v, total);
  }
  return total;
}

// index of the first word of a crossover block that is not a success (4: all four succeed)
__device__ __forceinline__ int first_failure(const u32x4& w, uint64_t thr) {
  int first = 4;
  if (!((uint64_t)w.w < thr)) first = 3;
  if (!((uint64_t)w.z < thr)) first = 2;
  if (!((uint64_t)w.y < thr)) first = 1;
  if (!((uint64_t)w.x < thr)) first = 0;
  return first;
}

template <int COST, int LPR, int K>
__global__ void __launch_bounds__(kRowsWarps * 32)
    de_step_rows_kernel(const __grid_constant__ StepParams P, int base, int xover, int /*ndonors*/) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(16) double smem[];
  __shared__ double sh_e[kRowsWarps];
  __shared__ int64_t sh_i[kRowsWarps];
  __shared__ unsigned long long sh_ninf[kRowsWarps], sh_nacc[kRowsWarps];
  constexpr int RPW = 32 / LPR;  // rows per warp
  constexpr unsigned LMASK = (1u << LPR) - 1u;
  constexpr int NB = (K + 2) / 2;  // Philox blocks of the donor stream (K donors + start index)
  static_assert(LPR >= 4 && LPR <= 16 && NB < LPR, "group too narrow for the draws");

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
  const bool needs_best = (base == BASE_BEST1 || base == BASE_RANDTOBEST1 || base == BASE_BEST2);
  double* s_best = smem;                                                         // [D]
  double* t_trial = smem + D + (size_t)warp * (2 * RPW * TS) + (size_t)grp * TS;  // this row's trial tile
  double* t_parent = t_trial + RPW * TS;

  if (P.mode == MODE_STEP && needs_best) {
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
    const double* __restrict__ prow = P.pop_cur + cc * D;

    // ---- 1. parent row: issue the loads first, they are consumed after the draws
    double2 p[kRowsTrips][2];
#pragma unroll
    for (int t = 0; t < kRowsTrips; ++t) {
      const int q = t * LPR + sub;
      if (q < NQ) {
        p[t][0] = ld_stream2(prow + 4 * q);
        p[t][1] = ld_stream2(prow + 4 * q + 2);
      } else {
        p[t][0] = make_double2(0.0, 0.0);
        p[t][1] = make_double2(0.0, 0.0);
      }
    }
    double e_old = CUDART_INF;
    if (P.mode == MODE_STEP) e_old = __ldg(P.e_cur + cc);

    // ---- 2. draws
    int64_t r[K];
    int nstart = 0, L = 0;
    const uint32_t cand_g = (uint32_t)(P.global_offset + cc);
    if (P.mode == MODE_STEP) {
      if (P.replay) {
#pragma unroll
        for (int j = 0; j < K; ++j) r[j] = __ldg(P.r_donors + cc * 5 + j);
        nstart = __ldg(P.r_nstart + cc);
        if (xover == XOVER_EXP) L = exp_length_thread(P, cand_g, cc, D);
      } else {
        const bool donor_lane = sub < NB;
        u32x4 v = philox4x32_10((uint32_t)(donor_lane ? sub : sub - NB), cand_g, P.generation,
                                donor_lane ? STREAM_DONORS : STREAM_XOVER, P.key0, P.key1);
        const uint64_t even = ((uint64_t)v.x << 32) | v.y, odd = ((uint64_t)v.z << 32) | v.w;
        uint64_t rw[2 * NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          rw[2 * b] = __shfl_sync(0xffffffffu, even, gbase + b);
          rw[2 * b + 1] = __shfl_sync(0xffffffffu, odd, gbase + b);
        }
        resolve_donors<K>(rw, cc, P.np, D, r, &nstart);
        if (xover == XOVER_EXP) {
          // strategy.py:48-56: L = number of leading successes, at most D
          int first = donor_lane ? 4 : first_failure(v, P.thr);
          int done = LPR - NB;  // crossover blocks examined so far
          bool resolved = false;
          {
            const unsigned seg = (__ballot_sync(0xffffffffu, first < 4) >> gbase) & LMASK;
            const int src = seg ? (__ffs(seg) - 1) : sub;
            const int f = __shfl_sync(0xffffffffu, first, gbase + src);
            if (seg) {
              L = 4 * (src - NB) + f;
              resolved = true;
            }
          }
          while (__any_sync(0xffffffffu, !resolved && 4 * done < D)) {
            v = philox4x32_10((uint32_t)(done + sub), cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
            first = first_failure(v, P.thr);
            const unsigned seg = (__ballot_sync(0xffffffffu, first < 4) >> gbase) & LMASK;
            const int src = seg ? (__ffs(seg) - 1) : sub;
            const int f = __shfl_sync(0xffffffffu, first, gbase + src);
            if (!resolved && seg) {
              L = 4 * (done + src) + f;
              resolved = true;
            }
            done += LPR;
          }
          if (!resolved || L > D) L = D;
        }
      }
    }

    // ---- own chunks -> tiles (generation 0: clipped into the strict range first)
    bool oob = false;
#pragma unroll
    for (int t = 0; t < kRowsTrips; ++t) {
      const int q = t * LPR + sub;
      if (q < NQ) {
        double x[4] = {p[t][0].x, p[t][0].y, p[t][1].x, p[t][1].y};
        if (P.mode == MODE_STEP) {
          *reinterpret_cast<double2*>(t_parent + 4 * q) = p[t][0];
          *reinterpret_cast<double2*>(t_parent + 4 * q + 2) = p[t][1];
        }
        // a parent row is inside the range once generation 0 has clipped it (rows_in_range);
        // otherwise the untouched elements of the trial are treated like the reference treats them
        if (P.bounds_mode != BOUNDS_NONE && !(P.mode == MODE_STEP && P.rows_in_range)) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const double l = uni_b ? P.lo_s : __ldg(P.lo + 4 * q + e);
            const double h = uni_b ? P.hi_s : __ldg(P.hi + 4 * q + e);
            if (P.mode == MODE_GEN0) {
              x[e] = fmin(fmax(x[e], l), h);
              oob |= (x[e] < l) || (x[e] > h);
            } else if (P.mode == MODE_STEP && P.bounds_mode == BOUNDS_CLAMP) {
              x[e] = (x[e] > l) ? x[e] : l;  // python max(lo, x)
              x[e] = (x[e] < h) ? x[e] : h;  // python min(hi, x)
            } else {
              oob |= (x[e] < l) || (x[e] > h);
            }
          }
        }
        *reinterpret_cast<double2*>(t_trial + 4 * q) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2*>(t_trial + 4 * q + 2) = make_double2(x[2], x[3]);
      }
    }
    __syncwarp();

    // ---- 3. crossover window: chunks q0 .. q0+nq-1 (mod NQ) of the trial tile
    if (P.mode == MODE_STEP) {
      int q0 = 0, nq = NQ;
      if (xover == XOVER_EXP) {
        q0 = nstart >> 2;
        nq = (L > 0) ? (((nstart + L - 1) >> 2) - q0 + 1) : 0;
        if (nq > NQ) nq = NQ;  // a full-length window that starts inside a chunk: each chunk once
      }
      const double* __restrict__ d0 = P.pop_cur + r[0] * D;
      const double* __restrict__ d1 = P.pop_cur + r[1] * D;
      const double* __restrict__ d2 = P.pop_cur + r[K > 2 ? 2 : 0] * D;
      const double* __restrict__ d3 = P.pop_cur + r[K > 3 ? 3 : 0] * D;
      const double* __restrict__ d4 = P.pop_cur + r[K > 4 ? 4 : 0] * D;
      for (int w = sub; w < nq; w += LPR) {
        int q = q0 + w;
        if (q >= NQ) q -= NQ;
        const int i0 = q << 2;
        bool cross[4];
        if (xover == XOVER_BIN) {
          // strategy.py:79-85: i == n or random() < CR, one draw per dimension
          if (P.replay) {
            const double* u = P.r_uni + cc * (int64_t)(D + 1) + i0;
#pragma unroll
            for (int e = 0; e < 4; ++e) cross[e] = (i0 + e == nstart) || (__ldg(u + e) < P.CR);
          } else {
            const u32x4 wv = philox4x32_10((uint32_t)q, cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
            cross[0] = (i0 == nstart) || ((uint64_t)wv.x < P.thr);
            cross[1] = (i0 + 1 == nstart) || ((uint64_t)wv.y < P.thr);
            cross[2] = (i0 + 2 == nstart) || ((uint64_t)wv.z < P.thr);
            cross[3] = (i0 + 3 == nstart) || ((uint64_t)wv.w < P.thr);
          }
        } else {
          int rel = i0 - nstart;
          if (rel < 0) rel += D;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int re = rel + e;
            if (re >= D) re -= D;
            cross[e] = re < L;
          }
        }
        if (cross[0] | cross[1] | cross[2] | cross[3]) {
          double dn[5][4];
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            if (j < K) {
              const double* dr = (j == 0) ? d0 : ((j == 1) ? d1 : ((j == 2) ? d2 : ((j == 3) ? d3 : d4)));
              const double2 a = ld_stream2(dr + i0), b = ld_stream2(dr + i0 + 2);
              dn[j][0] = a.x;
              dn[j][1] = a.y;
              dn[j][2] = b.x;
              dn[j][3] = b.y;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) dn[j][e] = 0.0;
            }
          }
          const double2 xa = *reinterpret_cast<const double2*>(t_trial + i0);
          const double2 xb = *reinterpret_cast<const double2*>(t_trial + i0 + 2);
          double x[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (cross[e]) {
              const double pe[5] = {dn[0][e], dn[1][e], dn[2][e], dn[3][e], dn[4][e]};
              const double b = needs_best ? s_best[i0 + e] : 0.0;
              double y;
              if (K == 2) {
                y = (base == BASE_BEST1) ? mutant<BASE_BEST1>(x[e], b, pe, P.F) : mutant<BASE_RANDTOBEST1>(x[e], b, pe, P.F);
              } else if (K == 3) {
                y = mutant<BASE_RAND1>(x[e], b, pe, P.F);
              } else if (K == 4) {
                y = mutant<BASE_BEST2>(x[e], b, pe, P.F);
              } else {
                y = mutant<BASE_RAND2>(x[e], b, pe, P.F);
              }
              if (P.bounds_mode != BOUNDS_NONE) {
                const double l = uni_b ? P.lo_s : __ldg(P.lo + i0 + e);
                const double h = uni_b ? P.hi_s : __ldg(P.hi + i0 + e);
                if (P.bounds_mode == BOUNDS_CLAMP) {
                  y = (y > l) ? y : l;
                  y = (y < h) ? y : h;
                } else {
                  oob |= (y < l) || (y > h);
                }
              }
              x[e] = y;
            }
          }
          *reinterpret_cast<double2*>(t_trial + i0) = make_double2(x[0], x[1]);
          *reinterpret_cast<double2*>(t_trial + i0 + 2) = make_double2(x[2], x[3]);
        }
      }
    }
    // any lane of the group out of range?
    oob = ((__ballot_sync(0xffffffffu, oob) >> gbase) & LMASK) != 0;
    __syncwarp();

    // ---- 4. energy = (inf if out of range else cost(x)) + penalty(x); idle groups evaluate their
    // shadow row because the shuffles inside need every lane
    double E = CUDART_INF;
    {
      const double cost = rows_cost<COST, LPR>(t_trial, NQ, sub, gbase, P.cost);
      if (!oob) E = cost;
      E = dadd(E, rows_penalty<LPR>(t_trial, D, NQ, sub, P.pen));
    }

    if (live && P.cap_trial != nullptr && P.mode != MODE_EVAL) {
      for (int i = sub; i < D; i += LPR) P.cap_trial[c * D + i] = t_trial[i];
      if (sub == 0) P.cap_energy[c] = E;
    }

    // ---- 5. selection (differential_evolution.py:603-613) and the next-generation row
    if (P.mode == MODE_EVAL) {
      if (live && sub == 0) P.e_next[c] = E;
    } else if (live) {
      const bool acc = E < e_old;
      const double* src = (acc || P.mode == MODE_GEN0) ? t_trial : t_parent;
      double* __restrict__ nrow = P.pop_next + c * D;
#pragma unroll
      for (int t = 0; t < kRowsTrips; ++t) {
        const int q = t * LPR + sub;
        if (q < NQ) {
          const double2 a = *reinterpret_cast<const double2*>(src + 4 * q);
          const double2 b = *reinterpret_cast<const double2*>(src + 4 * q + 2);
          st_stream2(nrow + 4 * q, a.x, a.y);
          st_stream2(nrow + 4 * q + 2, b.x, b.y);
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
