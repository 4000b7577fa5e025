// Several candidate rows per warp, for short rows (D <= 128).
//
// With one warp per row, a 64-dimensional row (512 B) costs ~1000 warp instructions, most of
// them per-row overhead (draws, reductions, selection) executed by lanes that have no data.
// Here a group of LPR lanes (4, 8, 16 or 32) owns a row and a warp processes 32/LPR rows at
// once, so that overhead is shared by SIMT: the same instructions serve 2-8 rows.  Loads stay
// 16-byte vectors on the contiguous row segments.
//
// The kernel serves every mode (step, generation 0, mb2_eval_cost) for its row lengths, so the
// energies of a context always come from one summation order: lane `sub` of a group accumulates
// elements sub, sub+LPR, ... in order and the group combines the lanes by a xor tree.
#pragma once
#include "de_kernels.cuh"

namespace mb2 {

constexpr int kSubMaxDim = 128;
constexpr int kSubWarps = 8;

template <int LPR, bool PROD>
__device__ __forceinline__ double seg_reduce(double v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = PROD ? dmul(v, w) : dadd(v, w);
  }
  return v;
}

template <int COST, int LPR>
__device__ __forceinline__ double seg_cost(const double* xs, int D, int sub, const CostParams& cp) {
  if (COST == COST_ROSEN) {
    double acc = 0.0;
    for (int i = sub; i < D - 1; i += LPR) {
      const double a = xs[i], b = xs[i + 1];
      const double t = dsub(b, dmul(a, a));
      const double u = dsub(1.0, a);
      acc = dadd(acc, dadd(dmul(100.0, dmul(t, t)), dmul(u, u)));
    }
    return seg_reduce<LPR, false>(acc);
  } else if (COST == COST_CORANA) {
    const double d[4] = {1., 1000., 10., 100.};
    double x[4] = {0., 0., 0., 0.};
    if (D >= 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = xs[j];
    } else {
      if (D >= 1) x[0] = xs[0];
      if (D >= 2) x[2] = xs[1];
      if (D >= 3) x[3] = xs[2];
    }
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double xj = x[j];
      const double sg = (xj > 0.0) ? 1.0 : ((xj < 0.0) ? -1.0 : 0.0);
      const double zj = dmul(dmul(floor(dadd(fabs(xj / 0.2), 0.49999)), sg), 0.2);
      double term;
      if (fabs(dsub(xj, zj)) < 0.05) {
        const double sz = (zj > 0.0) ? 1.0 : ((zj < 0.0) ? -1.0 : 0.0);
        const double w = dsub(zj, dmul(0.05, sz));
        term = dmul(dmul(0.15, dmul(w, w)), d[j]);
      } else {
        term = dmul(dmul(d[j], xj), xj);
      }
      r = dadd(r, term);
    }
    return r;
  } else if (COST == COST_GRIEWANGK) {
    double s = 0.0, p = 1.0;
    for (int i = sub; i < D; i += LPR) {
      const double c = xs[i];
      s = dadd(s, dmul(c, c));
      const double2 t = __ldg(cp.grw_tab + i);
      p = dmul(p, grw_factor(c, t.x, t.y, cp.grw_bounded != 0));
    }
    s = seg_reduce<LPR, false>(s);
    p = seg_reduce<LPR, true>(p);
    return dadd(dsub(s / 4000.0, p), 1.0);
  } else if (COST == COST_ZIMMERMANN) {
    return fixed_cost(cp.fixed_kind, xs, D, cp);
  } else if (COST == COST_SEPARABLE) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = sub; i < D; i += LPR) sep_accumulate(cp.sep_kind, xs[i], i, s1, s2);
    s1 = seg_reduce<LPR, false>(s1);
    s2 = seg_reduce<LPR, false>(s2);
    return sep_finish(cp.sep_kind, s1, s2, D);
  } else if (COST == COST_POLYFIT || COST == COST_LORENTZIAN) {
    double acc = 0.0;
    const int M = cp.cheb_M;
    for (int i = sub; i < M; i += LPR) acc = dadd(acc, fit_term<COST>(xs, D, cp, i));
    return seg_reduce<LPR, false>(acc);
  } else {
    double acc = 0.0;
    const int M = cp.cheb_M;
    for (int i = sub; i < M; i += LPR) {
      const double px = horner(xs, D, __ldg(cp.cheb_x + i));
      if (px < -1.0 || px > 1.0) {
        const double u = dsub(1.0, px);
        acc = dadd(acc, dmul(u, u));
      }
    }
    double r = seg_reduce<LPR, false>(acc);
    double px = dsub(horner(xs, D, 1.2), cp.cheb_tp);
    if (px < 0.0) r = dadd(r, dmul(px, px));
    px = dsub(horner(xs, D, -1.2), cp.cheb_tm);
    if (px < 0.0) r = dadd(r, dmul(px, px));
    return r;
  }
}

template <int LPR>
__device__ __forceinline__ double seg_penalty(const double* xs, int D, int sub, const PenaltyParams& pp) {
  double total = 0.0;
  for (int t = 0; t < pp.n; ++t) {
    const double* g = pp.G + (int64_t)t * D;
    double acc = 0.0;
    for (int i = sub; i < D; i += LPR) acc = dadd(acc, dmul(__ldg(g + i), xs[i]));
    const double pf = dsub(seg_reduce<LPR, false>(acc), __ldg(pp.h + t));
    const double v = penalty_term(__ldg(pp.ptype + t), __ldg(pp.kscale + t), __ldg(pp.mult + t), pf);
    total = dadd(v, total);
  }
  return total;
}

// The kernel body; `P` is the launch's parameter block (de_step_subwarp_kernel) or a per-island copy
// of it (de_step_subwarp_islands_kernel); `part` indexes this CTA's partials.
template <int COST, int LPR, bool VEC>
__device__ __forceinline__ void subwarp_body(const StepParams& P, int base, int xover, int ndonors, int part) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(16) double smem[];
  __shared__ double sh_e[kSubWarps];
  __shared__ int64_t sh_i[kSubWarps];
  __shared__ unsigned long long sh_ninf[kSubWarps], sh_nacc[kSubWarps];
  constexpr int RPW = 32 / LPR;  // rows per warp

  const int D = P.dim;
  const int Dp = P.dim_pad;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const int grp = lane / LPR;
  const int sub = lane % LPR;
  const bool uni_b = P.bounds_uniform != 0;
  const bool needs_best = (base == BASE_BEST1 || base == BASE_RANDTOBEST1 || base == BASE_BEST2);
  double* s_best = smem;                                                  // [Dp]
  double* s_trial = smem + Dp + ((size_t)warp * RPW + grp) * Dp;          // [Dp] per row

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
    bool oob = false;

    if (P.mode == MODE_STEP) {
      int64_t r[5] = {0, 0, 0, 0, 0};
      int nstart;
      const uint32_t cand_g = (uint32_t)(P.global_offset + cc);
      uint32_t xword = 0;  // low half of the start-index word: the geometric draw of the crossover length
      if (P.replay) {
        for (int j = 0; j < ndonors; ++j) r[j] = __ldg(P.r_donors + cc * 5 + j);
        nstart = __ldg(P.r_nstart + cc);
      } else {
        switch (ndonors) {
          case 2: draw_donors<2>(P.key0, P.key1, P.generation, cand_g, cc, P.np, D, r, &nstart, &xword); break;
          case 3: draw_donors<3>(P.key0, P.key1, P.generation, cand_g, cc, P.np, D, r, &nstart, &xword); break;
          case 4: draw_donors<4>(P.key0, P.key1, P.generation, cand_g, cc, P.np, D, r, &nstart, &xword); break;
          default: draw_donors<5>(P.key0, P.key1, P.generation, cand_g, cc, P.np, D, r, &nstart, &xword); break;
        }
      }
      // exponential crossover length (strategy.py:48-56): one geometric draw (philox.cuh, protocol 2)
      const int L = (xover == XOVER_EXP) ? exp_length_thread(P, xword, cc, D) : 0;
      const double* __restrict__ d0 = P.pop_cur + r[0] * D;
      const double* __restrict__ d1 = P.pop_cur + r[1] * D;
      const double* __restrict__ d2 = P.pop_cur + r[2] * D;
      const double* __restrict__ d3 = P.pop_cur + r[3] * D;
      const double* __restrict__ d4 = P.pop_cur + r[4] * D;

      for (int i0 = sub * 4; i0 < D; i0 += LPR * 4) {
        double x[4];
        load4<VEC>(prow, i0, D, x);
        bool cross[4];
        if (xover == XOVER_BIN) {
          if (P.replay) {
            const double* u = P.r_uni + cc * (int64_t)(D + 1) + i0;
#pragma unroll
            for (int e = 0; e < 4; ++e) cross[e] = (i0 + e < D) && ((i0 + e == nstart) || (__ldg(u + e) < P.CR));
          } else {
            const u32x4 w = philox4x32_10((uint32_t)(i0 >> 3), cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
            const uint32_t wa = (i0 & 4) ? w.z : w.x, wb = (i0 & 4) ? w.w : w.y;
            cross[0] = (i0 == nstart) || field_lo_hit(wa, P.thr16);
            cross[1] = (i0 + 1 < D) && ((i0 + 1 == nstart) || field_hi_hit(wa, P.thr16));
            cross[2] = (i0 + 2 < D) && ((i0 + 2 == nstart) || field_lo_hit(wb, P.thr16));
            cross[3] = (i0 + 3 < D) && ((i0 + 3 == nstart) || field_hi_hit(wb, P.thr16));
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int rel = i0 + e - nstart;
            if (rel < 0) rel += D;
            cross[e] = (i0 + e < D) && (rel < L);
          }
        }
        if (cross[0] | cross[1] | cross[2] | cross[3]) {
          double q[5][4];
          load4<VEC>(d0, i0, D, q[0]);
          load4<VEC>(d1, i0, D, q[1]);
          if (ndonors > 2) load4<VEC>(d2, i0, D, q[2]);
          if (ndonors > 3) load4<VEC>(d3, i0, D, q[3]);
          if (ndonors > 4) load4<VEC>(d4, i0, D, q[4]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (cross[e]) {
              double p[5];
#pragma unroll
              for (int j = 0; j < 5; ++j) p[j] = (j < ndonors) ? q[j][e] : 0.0;
              const double b = needs_best ? s_best[i0 + e] : 0.0;
              switch (base) {
                case BASE_BEST1: x[e] = mutant<BASE_BEST1>(x[e], b, p, P.F); break;
                case BASE_RAND1: x[e] = mutant<BASE_RAND1>(x[e], b, p, P.F); break;
                case BASE_RANDTOBEST1: x[e] = mutant<BASE_RANDTOBEST1>(x[e], b, p, P.F); break;
                case BASE_BEST2: x[e] = mutant<BASE_BEST2>(x[e], b, p, P.F); break;
                default: x[e] = mutant<BASE_RAND2>(x[e], b, p, P.F); break;
              }
            }
          }
        }
        if (P.bounds_mode != BOUNDS_NONE) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (i0 + e < D) {
              const double l = uni_b ? P.lo_s : __ldg(P.lo + i0 + e);
              const double h = uni_b ? P.hi_s : __ldg(P.hi + i0 + e);
              if (P.bounds_mode == BOUNDS_CLAMP) {
                x[e] = (x[e] > l) ? x[e] : l;  // python max(lo, x)
                x[e] = (x[e] < h) ? x[e] : h;  // python min(hi, x)
              } else {
                oob |= (x[e] < l) || (x[e] > h);
              }
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) s_trial[i0 + e] = x[e];
      }
    } else {
      for (int i0 = sub * 4; i0 < D; i0 += LPR * 4) {
        double x[4];
        load4<VEC>(prow, i0, D, x);
        if (P.bounds_mode != BOUNDS_NONE) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (i0 + e < D) {
              const double l = uni_b ? P.lo_s : __ldg(P.lo + i0 + e);
              const double h = uni_b ? P.hi_s : __ldg(P.hi + i0 + e);
              if (P.mode == MODE_GEN0) x[e] = fmin(fmax(x[e], l), h);
              oob |= (x[e] < l) || (x[e] > h);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) s_trial[i0 + e] = x[e];
      }
    }
    // any lane of the group out of range?  (segmented OR over the LPR lanes)
    {
      const unsigned m = __ballot_sync(0xffffffffu, oob);
      const unsigned seg = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (grp * LPR));
      oob = (m & seg) != 0;
    }
    __syncwarp();

    double E = CUDART_INF;
    {
      // every group evaluates (idle groups on a shadow row): the shuffles inside need all lanes
      const double cost = seg_cost<COST, LPR>(s_trial, D, sub, P.cost);
      if (!oob) E = cost;
      E = dadd(E, seg_penalty<LPR>(s_trial, D, sub, P.pen));
    }

    if (live && P.cap_trial != nullptr && P.mode != MODE_EVAL) {
      for (int i = sub; i < D; i += LPR) P.cap_trial[c * D + i] = s_trial[i];
      if (sub == 0) P.cap_energy[c] = E;
    }

    if (P.mode == MODE_EVAL) {
      if (live && sub == 0) P.e_next[c] = E;
    } else if (live) {
      const double e_old = (P.mode == MODE_GEN0) ? CUDART_INF : __ldg(P.e_cur + c);
      const bool acc = E < e_old;
      double* __restrict__ nrow = P.pop_next + c * D;
      if (acc || P.mode == MODE_GEN0) {
        for (int i0 = sub * 4; i0 < D; i0 += LPR * 4) {
          double x[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = s_trial[i0 + e];
          store4<VEC>(nrow, i0, D, x);
        }
      } else {
        for (int i0 = sub * 4; i0 < D; i0 += LPR * 4) {
          double x[4];
          load4<VEC>(prow, i0, D, x);
          store4<VEC>(nrow, i0, D, x);
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
    __syncwarp();  // s_trial is reused by the next chunk
  }

  if (P.mode == MODE_EVAL) return;

  // warp argmin over the group leaders (other lanes carry +inf / 0)
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
    P.part_e[part] = be;
    P.part_idx[part] = bi;
    P.part_ninf[part] = ni;
    P.part_nacc[part] = na;
  }
}

template <int COST, int LPR, bool VEC>
__global__ void __launch_bounds__(kSubWarps * 32)
    de_step_subwarp_kernel(const __grid_constant__ StepParams P, int base, int xover, int ndonors) {
  subwarp_body<COST, LPR, VEC>(P, base, xover, ndonors, (int)blockIdx.x);
}

// Batch of independent solvers: blockIdx.y is the island.  P describes ONE island
// (np rows, first island's pointers); island y works on rows [y*np, (y+1)*np) of the buffers, draws
// its donors among them, follows its own best member best_x[y*dim..] and reports into its own
// partials -- exactly what a shard of a sharded population does (global_offset keys the Philox
// counters, so every island has its own draws).
template <int COST, int LPR, bool VEC>
__global__ void __launch_bounds__(kSubWarps * 32)
    de_step_subwarp_islands_kernel(const __grid_constant__ StepParams P, int base, int xover, int ndonors) {
  StepParams Q = P;
  const int64_t isl = blockIdx.y;
  const int64_t row0 = isl * P.np;
  Q.pop_cur += row0 * P.dim;
  if (Q.pop_next != nullptr) Q.pop_next += row0 * P.dim;
  if (Q.e_cur != nullptr) Q.e_cur += row0;
  Q.e_next += row0;
  if (P.stop != nullptr) {
    // device-side termination of a batch of solvers: P.stop is the stop flag of island 0's RunState.  An island that has
    // met its condition keeps the state it stopped with (abstract_ensemble_solver.py: a nested solver returns when ITS
    // Solve() ends) while the others go on: its rows and energies pass to the next buffer unchanged.
    const int* stop = reinterpret_cast<const int*>(reinterpret_cast<const char*>(P.stop) + isl * (int64_t)sizeof(RunState));
    if (*stop) {
      if (Q.pop_next != nullptr && Q.e_cur != nullptr) {
        const int64_t n = P.np * P.dim, step = (int64_t)gridDim.x * blockDim.x;
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += step) Q.pop_next[t] = Q.pop_cur[t];
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P.np; t += step) Q.e_next[t] = Q.e_cur[t];
      }
      return;
    }
    Q.stop = nullptr;
  }
  Q.best_x += isl * P.dim;
  Q.global_offset += row0;
  if (Q.r_donors != nullptr) {
    Q.r_donors += row0 * 5;
    Q.r_nstart += row0;
    Q.r_uni += row0 * (int64_t)(P.dim + 1);
  }
  if (Q.cap_trial != nullptr) {
    Q.cap_trial += row0 * P.dim;
    Q.cap_energy += row0;
  }
  subwarp_body<COST, LPR, VEC>(Q, base, xover, ndonors, (int)(isl * gridDim.x + blockIdx.x));
}

}  // namespace mb2
