// Thread-per-candidate generation kernel for short rows (D <= 128).
//
// With a warp (or several) per row, a 64-dimensional candidate (512 B) costs ~800 warp
// instructions of mostly per-row overhead; here a warp owns a batch of 32 consecutive candidates:
//   A  the 32 parent rows (one contiguous 32*8*D-byte block) are loaded cooperatively, coalesced,
//      into a padded shared-memory tile (odd row stride: conflict-free column access);
//   B  every lane turns its own row into the trial vector in place: draws (replayed or Philox),
//      crossover window / mask, mutation from the donors' global rows (only the crossed
//      dimensions are read), bounds;
//   C  every lane evaluates cost and penalty of its row SEQUENTIALLY, in the reference's own
//      summation order (rosen: numpy's pairwise sum; griewangk, chebyshev, penalties: left to
//      right), so energies are bit-identical to the reference except for libm differences;
//   D  greedy selection; rejected rows get their crossed dimensions back from the parent;
//   E  the 32 next-generation rows leave cooperatively, coalesced.
// The kernel also serves generation 0 and mb2_eval_cost for these row lengths, so energies within
// one context always come from the same code.
#pragma once
#include "de_kernels.cuh"

namespace mb2 {

constexpr int kSmallMaxDim = 128;
constexpr int kSmallWarps = 4;

// numpy.sum of n <= 128 doubles (pairwise_sum_DOUBLE: 8 running sums, then the tail)
template <typename F>
__device__ __forceinline__ double np_sum_le128(int n, F term) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = dadd(res, term(i));
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = term(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = dadd(r[j], term(i + j));
  }
  double res = dadd(dadd(dadd(r[0], r[1]), dadd(r[2], r[3])), dadd(dadd(r[4], r[5]), dadd(r[6], r[7])));
  for (; i < n; ++i) res = dadd(res, term(i));
  return res;
}

// cost of one row held by one thread, reference summation order
template <int COST>
__device__ __forceinline__ double thread_cost(const double* x, int D, const CostParams& cp) {
  if (COST == COST_ROSEN) {
    // dejong.py:98-101, numpy.sum order
    if (D < 2) return 0.0;
    return np_sum_le128(D - 1, [&](int i) {
      const double a = x[i], b = x[i + 1];
      const double t = dsub(b, dmul(a, a));
      const double u = dsub(1.0, a);
      return dadd(dmul(100.0, dmul(t, t)), dmul(u, u));
    });
  } else if (COST == COST_CORANA) {
    const double d[4] = {1., 1000., 10., 100.};
    double v[4] = {0., 0., 0., 0.};
    if (D >= 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = x[j];
    } else {
      if (D >= 1) v[0] = x[0];
      if (D >= 2) v[2] = x[1];
      if (D >= 3) v[3] = x[2];
    }
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double xj = v[j];
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
    // storn.py:135-139, left to right
    double s = 0.0, p = 1.0;
    for (int i = 0; i < D; ++i) s = dadd(s, dmul(x[i], x[i]));
    for (int i = 0; i < D; ++i) {
      const double2 t = __ldg(cp.grw_tab + i);
      p = dmul(p, grw_factor(x[i], t.x, t.y, cp.grw_bounded != 0));
    }
    return dadd(dsub(s / 4000.0, p), 1.0);
  } else if (COST == COST_ZIMMERMANN) {
    return fixed_cost(cp.fixed_kind, x, D, cp);
  } else if (COST == COST_SEPARABLE) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < D; ++i) sep_accumulate(cp.sep_kind, x[i], i, s1, s2);
    return sep_finish(cp.sep_kind, s1, s2, D);
  } else if (COST == COST_POLYFIT || COST == COST_LORENTZIAN) {
    double result = 0.0;
    const int M = cp.cheb_M;
    for (int i = 0; i < M; ++i) result = dadd(result, fit_term<COST>(x, D, cp, i));
    return result;
  } else {
    // _model_helper.py:16-33, sequential over the abscissae
    double result = 0.0;
    const int M = cp.cheb_M;
    for (int i = 0; i < M; ++i) {
      const double px = horner(x, D, __ldg(cp.cheb_x + i));
      if (px < -1.0 || px > 1.0) {
        const double u = dsub(1.0, px);
        result = dadd(result, dmul(u, u));
      }
    }
    double px = dsub(horner(x, D, 1.2), cp.cheb_tp);
    if (px < 0.0) result = dadd(result, dmul(px, px));
    px = dsub(horner(x, D, -1.2), cp.cheb_tm);
    if (px < 0.0) result = dadd(result, dmul(px, px));
    return result;
  }
}

__device__ __forceinline__ double thread_penalty(const double* x, int D, const PenaltyParams& pp) {
  double pen = 0.0;
  for (int t = 0; t < pp.n; ++t) {
    const double* g = pp.G + (int64_t)t * D;
    double acc = 0.0;
    for (int j = 0; j < D; ++j) acc = dadd(acc, dmul(__ldg(g + j), x[j]));
    const double pf = dsub(acc, __ldg(pp.h + t));
    const double v = penalty_term(__ldg(pp.ptype + t), __ldg(pp.kscale + t), __ldg(pp.mult + t), pf);
    pen = dadd(v, pen);
  }
  return pen;
}

template <int COST>
__global__ void __launch_bounds__(kSmallWarps * 32)
    de_step_small_kernel(const __grid_constant__ StepParams P, int base, int xover, int ndonors) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(16) double smem[];
  __shared__ double sh_e[kSmallWarps];
  __shared__ int64_t sh_i[kSmallWarps];
  __shared__ unsigned long long sh_ninf[kSmallWarps], sh_nacc[kSmallWarps];

  const int D = P.dim;
  const int Dp = D | 1;  // odd row stride: lanes reading column i of 32 rows hit distinct banks
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  double* tile = smem + (size_t)warp * 32 * Dp;
  double* row = tile + lane * Dp;

  double w_best_e = CUDART_INF;
  int64_t w_best_idx = INT64_MAX;
  unsigned long long w_ninf = 0, w_nacc = 0;
  const bool uni_b = P.bounds_uniform != 0;
  const unsigned inv_d = (unsigned)((0x100000000ull + (unsigned)D - 1) / (unsigned)D);  // ceil(2^32 / D)

  const int64_t nbatch = (P.np + 31) / 32;
  const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
  const int64_t nw = (int64_t)gridDim.x * wpc;
  for (int64_t b = gw; b < nbatch; b += nw) {
    const int64_t c0 = b * 32;
    const int nrows = (int)((P.np - c0 < 32) ? (P.np - c0) : 32);
    const int64_t c = c0 + lane;
    const bool live = lane < nrows;

    // ---- A: coalesced load of the batch's parent rows (contiguous in memory), eight loads in
    //      flight per lane; element idx of the block lives at tile[idx / D][idx % D]
    {
      const double* __restrict__ src = P.pop_cur + c0 * D;
      const int total = nrows * D;
      for (int base = lane; base < total; base += 256) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int idx = base + 32 * k;
          v[k] = (idx < total) ? ld_stream1(src + idx) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int idx = base + 32 * k;
          if (idx < total) {
            const int r = (int)__umulhi((unsigned)idx, inv_d);  // idx / D (idx < 4096, D <= 128)
            tile[r * Dp + (idx - r * D)] = v[k];
          }
        }
      }
    }
    __syncwarp();

    // ---- B: trial vector in place
    bool oob = false, binomial = false;
    int nstart = 0, L = 0;
    double undo[kSmallMaxDim];  // parent values of the crossed dimensions (exponential crossover)
    if (live) {
      if (P.mode == MODE_STEP) {
        int64_t r[5] = {0, 0, 0, 0, 0};
        const uint32_t cand_g = (uint32_t)(P.global_offset + c);
        uint32_t xword = 0;  // low half of the start-index word: the geometric draw of the crossover length
        if (P.replay) {
          for (int j = 0; j < ndonors; ++j) r[j] = __ldg(P.r_donors + c * 5 + j);
          nstart = __ldg(P.r_nstart + c);
        } else {
          switch (ndonors) {
            case 2: draw_donors<2>(P.key0, P.key1, P.generation, cand_g, c, P.np, D, r, &nstart, &xword); break;
            case 3: draw_donors<3>(P.key0, P.key1, P.generation, cand_g, c, P.np, D, r, &nstart, &xword); break;
            case 4: draw_donors<4>(P.key0, P.key1, P.generation, cand_g, c, P.np, D, r, &nstart, &xword); break;
            default: draw_donors<5>(P.key0, P.key1, P.generation, cand_g, c, P.np, D, r, &nstart, &xword); break;
          }
        }
        const double* u = P.r_uni + c * (int64_t)(D + 1);
        const double* __restrict__ q0 = P.pop_cur + r[0] * D;
        const double* __restrict__ q1 = P.pop_cur + r[1] * D;
        const double* __restrict__ q2 = P.pop_cur + r[2] * D;
        const double* __restrict__ q3 = P.pop_cur + r[3] * D;
        const double* __restrict__ q4 = P.pop_cur + r[4] * D;
        auto mutate_dim = [&](int i) {
          double p[5];
          p[0] = ld_stream1(q0 + i);
          p[1] = ld_stream1(q1 + i);
          p[2] = ndonors > 2 ? ld_stream1(q2 + i) : 0.0;
          p[3] = ndonors > 3 ? ld_stream1(q3 + i) : 0.0;
          p[4] = ndonors > 4 ? ld_stream1(q4 + i) : 0.0;
          const double bst = __ldg(P.best_x + i);
          const double x = row[i];
          double v;
          switch (base) {
            case BASE_BEST1: v = mutant<BASE_BEST1>(x, bst, p, P.F); break;
            case BASE_RAND1: v = mutant<BASE_RAND1>(x, bst, p, P.F); break;
            case BASE_RANDTOBEST1: v = mutant<BASE_RANDTOBEST1>(x, bst, p, P.F); break;
            case BASE_BEST2: v = mutant<BASE_BEST2>(x, bst, p, P.F); break;
            default: v = mutant<BASE_RAND2>(x, bst, p, P.F); break;
          }
          row[i] = v;
        };
        if (xover == XOVER_BIN) {
          // strategy.py:77-85
          u32x4 w = {0, 0, 0, 0};
          for (int i = 0; i < D; ++i) {
            bool hit;
            if (P.replay) {
              hit = __ldg(u + i) < P.CR;
            } else {
              if ((i & 7) == 0) w = philox4x32_10((uint32_t)(i >> 3), cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
              hit = xover_field(w, i & 7) < P.thr16;
            }
            if (hit || i == nstart) mutate_dim(i);
          }
          binomial = true;  // every dimension may have changed: a rejected row is re-read
        } else {
          // strategy.py:48-56: L leading successes, capped at D
          L = exp_length_thread(P, xword, c, D);
          // window dims i_j = (nstart + j) mod D, four at a time: all donor loads of a chunk are
          // issued before any is used (the scattered 8-byte loads are the latency of this kernel).
          // Only crossed dimensions can leave the bounds (parents are inside them since generation
          // 0), so the clamp / range test runs on the window only.  The parent's values go to an
          // undo log that restores the row if the trial is rejected.
          for (int j0 = 0; j0 < L; j0 += 4) {
            int idx[4];
            double pv[4][5];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              int i = nstart + j0 + t;
              if (i >= D) i -= D;
              if (i >= D) i -= D;
              idx[t] = i;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (j0 + t < L) {
                pv[t][0] = ld_stream1(q0 + idx[t]);
                pv[t][1] = ld_stream1(q1 + idx[t]);
                pv[t][2] = ndonors > 2 ? ld_stream1(q2 + idx[t]) : 0.0;
                pv[t][3] = ndonors > 3 ? ld_stream1(q3 + idx[t]) : 0.0;
                pv[t][4] = ndonors > 4 ? ld_stream1(q4 + idx[t]) : 0.0;
              }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (j0 + t < L) {
                const int i = idx[t];
                const double x = row[i];
                undo[j0 + t] = x;
                const double bst = __ldg(P.best_x + i);
                double v;
                switch (base) {
                  case BASE_BEST1: v = mutant<BASE_BEST1>(x, bst, pv[t], P.F); break;
                  case BASE_RAND1: v = mutant<BASE_RAND1>(x, bst, pv[t], P.F); break;
                  case BASE_RANDTOBEST1: v = mutant<BASE_RANDTOBEST1>(x, bst, pv[t], P.F); break;
                  case BASE_BEST2: v = mutant<BASE_BEST2>(x, bst, pv[t], P.F); break;
                  default: v = mutant<BASE_RAND2>(x, bst, pv[t], P.F); break;
                }
                if (P.bounds_mode != BOUNDS_NONE) {
                  const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
                  const double h = uni_b ? P.hi_s : __ldg(P.hi + i);
                  if (P.bounds_mode == BOUNDS_CLAMP) {
                    v = (v > l) ? v : l;  // python max(lo, x); min(hi, x)
                    v = (v < h) ? v : h;
                  } else {
                    oob |= (v < l) || (v > h);
                  }
                }
                row[i] = v;
              }
            }
          }
        }
        if (xover == XOVER_BIN && P.bounds_mode != BOUNDS_NONE) {
          for (int i = 0; i < D; ++i) {
            const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
            const double h = uni_b ? P.hi_s : __ldg(P.hi + i);
            double x = row[i];
            if (P.bounds_mode == BOUNDS_CLAMP) {
              x = (x > l) ? x : l;  // python max(lo, x); min(hi, x)
              x = (x < h) ? x : h;
              row[i] = x;
            } else {
              oob |= (x < l) || (x > h);
            }
          }
        }
      } else if (P.bounds_mode != BOUNDS_NONE) {
        for (int i = 0; i < D; ++i) {
          const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
          const double h = uni_b ? P.hi_s : __ldg(P.hi + i);
          double x = row[i];
          if (P.mode == MODE_GEN0) {
            x = fmin(fmax(x, l), h);  // numpy clip (differential_evolution.py:516-520)
            row[i] = x;
          }
          oob |= (x < l) || (x > h);
        }
      }
    }

    // ---- C: energy = (inf if out of range else cost) + penalty, sequential per thread
    double E = CUDART_INF;
    bool acc = false;
    if (live) {
      if (!oob) E = thread_cost<COST>(row, D, P.cost);
      E = dadd(E, thread_penalty(row, D, P.pen));
      if (P.cap_trial != nullptr && P.mode != MODE_EVAL) {
        for (int i = 0; i < D; ++i) P.cap_trial[c * D + i] = row[i];
        P.cap_energy[c] = E;
      }
      if (P.mode == MODE_EVAL) {
        P.e_next[c] = E;
      } else {
        // ---- D: selection (differential_evolution.py:603-613)
        const double e_old = (P.mode == MODE_GEN0) ? CUDART_INF : __ldg(P.e_cur + c);
        acc = E < e_old;
        P.e_next[c] = acc ? E : e_old;
        if (acc) {
          ++w_nacc;
          if (E < w_best_e) {
            w_best_e = E;
            w_best_idx = c;
          }
        }
        if (isinf(E)) ++w_ninf;
        if (!acc && P.mode == MODE_STEP) {
          // rejected: the crossed dimensions get the parent's values back
          if (binomial) {
            const double* __restrict__ prow = P.pop_cur + c * D;
            for (int i = 0; i < D; ++i) row[i] = ld_stream1(prow + i);
          } else {
            int i = nstart;
            for (int j = 0; j < L; ++j) {
              row[i] = undo[j];
              if (++i == D) i = 0;
            }
          }
        }
      }
    }
    __syncwarp();

    // ---- E: coalesced store of the batch's next-generation rows
    if (P.mode != MODE_EVAL) {
      double* __restrict__ dst = P.pop_next + c0 * D;
      const int total = nrows * D;
      for (int idx = lane; idx < total; idx += 32) {
        const int r = (int)__umulhi((unsigned)idx, inv_d);
        dst[idx] = tile[r * Dp + (idx - r * D)];
      }
    }
    __syncwarp();  // the tile is reused by the next batch
  }

  if (P.mode == MODE_EVAL) return;

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
