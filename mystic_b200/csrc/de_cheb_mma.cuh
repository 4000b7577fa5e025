// Chebyshev polynomial-fit cost on FP64 tensor cores (SASS DMMA.8x8x4), fused into the DE
// generation step (MB2_COST_CHEBYSHEV_MMA).
//
// chebyshevNcost (mystic/models/_model_helper.py:16-33) evaluates the trial polynomial at M
// abscissae: px[c,i] = sum_j a[c,j] * x_i^(n-j), a population x coefficients by coefficients x
// points contraction.  A warp owns a batch of 64 candidates (8 row tiles of 8):
//   phase 1  every lane builds the trial vectors of two candidates (D <= 17 coefficients: thread
//            per candidate; any strategy, replayed or Philox draws, bounds) into shared memory;
//   phase 2  A fragments (coefficients of x^n..x^1, zero padded to 4*KS) stay in registers; for
//            every tile of 8 points the warp loads the Vandermonde fragments V[k][i] = x_i^(n-k)
//            once and issues 8*KS mma.sync.m8n8k4.f64 with the accumulator initialised to the
//            constant coefficient; the epilogue (px outside [-1,1] -> (1-px)^2) and the row sum
//            stay in registers, the [NP x M] product is never stored;
//   phase 3  the two end-point terms at x = +-1.2 (Horner), penalty, selection, argmin.
// The same kernel serves generation 0 and mb2_eval_cost for this cost id, so energies are
// self-consistent.  Unlike the Horner path (MB2_COST_CHEBYSHEV) px is not bit-identical to the
// reference's Horner evaluation (different association), so a point within ~1e-13 of the
// threshold |px| = 1 may flip; see DESIGN.md section 5.
#pragma once
#include <type_traits>

#include "de_kernels.cuh"

namespace mb2 {

#ifndef MB2_CHEB_MINBLOCKS
#define MB2_CHEB_MINBLOCKS 2
#endif
constexpr int kChebMaxDim = 17;    // order <= 16
constexpr int kChebBatch = 64;     // candidates per warp batch (8 row tiles x 8)
constexpr int kChebWarps = 8;

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
// first k-step of a tile: D = A.B + C with C left untouched (no register copies of the constant term)
__device__ __forceinline__ void dmma884_init(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// q = 1 - px: is px outside [-1, 1], i.e. q < 0 or q > 2?  Decided on the bit pattern, because the
// FP64 pipe the compares would run on is the one the MMAs need (ncu: 48 % of the stall samples were
// math-pipe throttle with two DSETPs per point): as unsigned, the high word of a negative double
// is >= 0x80000000 and that of a double above 2.0 is >= 0x40000000 (2.0 itself is 0x40000000:0).
// Infinities count as outside, like in the reference; a NaN would too, which the reference's
// `px<-1 or px>1` does not do -- rows whose coefficients are large enough for p(x) to overflow into
// a NaN are therefore evaluated by the Horner fallback of phase 3 (kChebSafeCoeff).
__device__ __forceinline__ bool outside_unit(double q) {
  const unsigned hi = (unsigned)__double2hiint(q);
  return hi > 0x40000000u || (hi == 0x40000000u && __double2loint(q) != 0);
}
constexpr double kChebSafeCoeff = 1e150;  // |coefficients| below this cannot overflow a degree <= 16 polynomial on |x| <= 1.2

// RESID: the epilogue is the squared residual of a polynomial least-squares data fit
// (MB2_COST_POLYFIT_MMA: sum(((p(x_i) - obs_i)/sigma)^2), forward_model.py:241-258 with
// models/poly.py:36-54) instead of the Chebyshev threshold terms; no end-point terms then.
template <int KS, bool RESID = false>  // k-steps of 4 coefficients: 1 (order <= 4), 2 (order <= 8), 4 (order <= 16)
__global__ void __launch_bounds__(kChebWarps * 32, MB2_CHEB_MINBLOCKS)
    de_step_cheb_mma_kernel(const __grid_constant__ StepParams P, int base, int xover, int ndonors) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(16) double smem[];
  __shared__ double sh_e[kChebWarps];
  __shared__ int64_t sh_i[kChebWarps];
  __shared__ unsigned long long sh_ninf[kChebWarps], sh_nacc[kChebWarps];

  const int D = P.dim;
  const int n = D - 1;  // polynomial order: coefficient j multiplies x^(n-j)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  double* s_trial = smem + (size_t)warp * (kChebBatch * D + kChebBatch);  // [64][D]
  double* s_sum = s_trial + kChebBatch * D;                                // [64]

  double w_best_e = CUDART_INF;
  int64_t w_best_idx = INT64_MAX;
  unsigned long long w_ninf = 0, w_nacc = 0;

  const int64_t nbatch = (P.np + kChebBatch - 1) / kChebBatch;
  const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
  const int64_t nw = (int64_t)gridDim.x * wpc;
  const bool uni_b = P.bounds_uniform != 0;

  for (int64_t b = gw; b < nbatch; b += nw) {
    const int64_t c0 = b * kChebBatch;
    unsigned oobmask = 0;  // bit h: this lane's candidate h is out of range
    unsigned bigmask = 0;  // bit h: a coefficient so large (or not finite) that p(x) may overflow: Horner fallback

    // ---------------- phase 1: trial vectors, one thread per candidate (two per lane)
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int cl = lane + 32 * h;
      const int64_t c = c0 + cl;
      double* row = s_trial + cl * D;
      if (c >= P.np) {
        for (int i = 0; i < D; ++i) row[i] = 0.0;
        continue;
      }
      const double* __restrict__ prow = P.pop_cur + c * D;
      bool oob = false;
      if (P.mode == MODE_STEP) {
        int64_t r[5] = {0, 0, 0, 0, 0};
        int nstart;
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
        // crossover decisions as a bit mask over the D <= 17 dimensions
        unsigned cmask = 0;
        const double* u = P.r_uni + c * (int64_t)(D + 1);
        if (xover == XOVER_BIN) {
          u32x4 w = {0, 0, 0, 0};
          for (int i = 0; i < D; ++i) {
            bool hit;
            if (P.replay) {
              hit = __ldg(u + i) < P.CR;
            } else {
              if ((i & 7) == 0) w = philox4x32_10((uint32_t)(i >> 3), cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
              hit = xover_field(w, i & 7) < P.thr16;
            }
            if (hit || i == nstart) cmask |= 1u << i;
          }
        } else {
          const int L = exp_length_thread(P, xword, c, D);
          for (int j = 0; j < L; ++j) {
            int i = nstart + j;
            if (i >= D) i -= D;
            cmask |= 1u << i;
          }
        }
        for (int i = 0; i < D; ++i) {
          double x = ld_stream1(prow + i);
          if (cmask & (1u << i)) {
            double p[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) p[j] = (j < ndonors) ? ld_stream1(P.pop_cur + r[j] * D + i) : 0.0;
            const double bst = __ldg(P.best_x + i);
            switch (base) {
              case BASE_BEST1: x = mutant<BASE_BEST1>(x, bst, p, P.F); break;
              case BASE_RAND1: x = mutant<BASE_RAND1>(x, bst, p, P.F); break;
              case BASE_RANDTOBEST1: x = mutant<BASE_RANDTOBEST1>(x, bst, p, P.F); break;
              case BASE_BEST2: x = mutant<BASE_BEST2>(x, bst, p, P.F); break;
              default: x = mutant<BASE_RAND2>(x, bst, p, P.F); break;
            }
          }
          if (P.bounds_mode != BOUNDS_NONE) {
            const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
            const double hh = uni_b ? P.hi_s : __ldg(P.hi + i);
            if (P.bounds_mode == BOUNDS_CLAMP) {
              x = (x > l) ? x : l;
              x = (x < hh) ? x : hh;
            } else {
              oob |= (x < l) || (x > hh);
            }
          }
          row[i] = x;
        }
      } else {
        for (int i = 0; i < D; ++i) {
          double x = ld_stream1(prow + i);
          if (P.bounds_mode != BOUNDS_NONE) {
            const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
            const double hh = uni_b ? P.hi_s : __ldg(P.hi + i);
            if (P.mode == MODE_GEN0) x = fmin(fmax(x, l), hh);
            oob |= (x < l) || (x > hh);
          }
          row[i] = x;
        }
      }
      if (oob) oobmask |= 1u << h;
      if (!RESID) {
        bool big = false;
        for (int i = 0; i < D; ++i) big |= !(fabs(row[i]) < kChebSafeCoeff);
        if (big) bigmask |= 1u << h;
      }
    }
    __syncwarp();

    // ---------------- phase 2: [64 x 4KS] . [4KS x M] on DMMA, epilogue in registers
    // Chebyshev: the accumulators hold q = 1 - px directly (negated coefficients, constant term
    // 1 - c_n), so the epilogue is one predicated DFMA per point: px > 1 <=> q < 0, px < -1 <=> q > 2.
    // Data fit: accumulators start from c_n - obs_i and hold the residual.
    double a[8][KS], cinit[8], sum[8];
    const int arow = lane >> 2, acol = lane & 3;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const double* row = s_trial + (r * 8 + arow) * D;
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int j = 4 * s + acol;
        const double v = (j < n) ? row[j] : 0.0;
        a[r][s] = RESID ? v : -v;
      }
      cinit[r] = RESID ? row[n] : (1.0 - row[n]);
      sum[r] = 0.0;
    }
    const double* __restrict__ vf = P.cost.cheb_vfrag;
    const int ntiles = P.cost.cheb_ntiles;
    const int M = P.cost.cheb_M;
    const double inv_sigma = RESID ? 1.0 / P.cost.fit_sigma : 1.0;
    const bool unit_sigma = !RESID || P.cost.fit_sigma == 1.0;
    double bnext[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) bnext[s] = __ldg(vf + (size_t)s * 32 + lane);
    // one tile of 8 points; TAIL: the last tile may hold fewer than 8
    auto tile = [&](int t, auto tail) {
      constexpr bool TAIL = decltype(tail)::value;
      double bf[KS];
#pragma unroll
      for (int s = 0; s < KS; ++s) bf[s] = bnext[s];
      if (t + 1 < ntiles) {
#pragma unroll
        for (int s = 0; s < KS; ++s) bnext[s] = __ldg(vf + ((size_t)(t + 1) * KS + s) * 32 + lane);
      }
      const int i0 = t * 8 + 2 * acol;
      const bool v0 = !TAIL || i0 < M, v1 = !TAIL || i0 + 1 < M;
      double o0 = 0.0, o1 = 0.0;
      if (RESID) {
        o0 = v0 ? __ldg(P.cost.fit_obs + i0) : 0.0;
        o1 = v1 ? __ldg(P.cost.fit_obs + i0 + 1) : 0.0;
      }
      double c0v[8], c1v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (RESID)
          dmma884_init(c0v[r], c1v[r], a[r][0], bf[0], cinit[r] - o0, cinit[r] - o1);
        else
          dmma884_init(c0v[r], c1v[r], a[r][0], bf[0], cinit[r], cinit[r]);
      }
#pragma unroll
      for (int s = 1; s < KS; ++s) {
#pragma unroll
        for (int r = 0; r < 8; ++r) dmma884(c0v[r], c1v[r], a[r][s], bf[s]);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double t0 = c0v[r], t1 = c1v[r];
        if (RESID) {
          if (!unit_sigma) {
            t0 *= inv_sigma;
            t1 *= inv_sigma;
          }
          if (v0) sum[r] = fma(t0, t0, sum[r]);
          if (v1) sum[r] = fma(t1, t1, sum[r]);
        } else {
          // _model_helper.py:23-24: if px<-1 or px>1: result += (1-px)**2   (NaN compares false)
          if (v0 && outside_unit(t0)) sum[r] = fma(t0, t0, sum[r]);
          if (v1 && outside_unit(t1)) sum[r] = fma(t1, t1, sum[r]);
        }
      }
    };
    for (int t = 0; t + 1 < ntiles; ++t) tile(t, std::false_type());
    tile(ntiles - 1, std::true_type());
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      double v = sum[r];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (acol == 0) s_sum[r * 8 + arow] = v;
    }
    __syncwarp();

    // ---------------- phase 3: end points, penalty, selection (thread per candidate)
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int cl = lane + 32 * h;
      const int64_t c = c0 + cl;
      if (c >= P.np) continue;
      const double* row = s_trial + cl * D;
      double E = CUDART_INF;
      if (!(oobmask & (1u << h))) {
        E = s_sum[cl];
        if (!RESID && (bigmask & (1u << h))) {
          // the reference's own loop (_model_helper.py:18-25), NaN-safe comparisons
          E = 0.0;
          for (int i = 0; i < P.cost.cheb_M; ++i) {
            const double px = horner(row, D, __ldg(P.cost.cheb_x + i));
            if (px < -1.0 || px > 1.0) {
              const double u = dsub(1.0, px);
              E = dadd(E, dmul(u, u));
            }
          }
        }
        if (!RESID) {
          // _model_helper.py:27-31: px = p_trial(+-1.2) - p_target(+-1.2); if px<0: result += px*px
          double vp = 0.0, vm = 0.0;
          for (int j = 0; j < D; ++j) {
            vp = fma(vp, 1.2, row[j]);
            vm = fma(vm, -1.2, row[j]);
          }
          double px = vp - P.cost.cheb_tp;
          if (px < 0.0) E = fma(px, px, E);
          px = vm - P.cost.cheb_tm;
          if (px < 0.0) E = fma(px, px, E);
        }
      }
      // penalty: t_{n-1} + (... + (t_0 + 0.0)), linear conditions, sequential dot (penalty.py)
      double pen = 0.0;
      for (int t = 0; t < P.pen.n; ++t) {
        const double* g = P.pen.G + (int64_t)t * D;
        double acc = 0.0;
        for (int j = 0; j < D; ++j) acc = dadd(acc, dmul(__ldg(g + j), row[j]));
        const double pf = dsub(acc, __ldg(P.pen.h + t));
        const double v = penalty_term(__ldg(P.pen.ptype + t), __ldg(P.pen.kscale + t), __ldg(P.pen.mult + t), pf);
        pen = dadd(v, pen);
      }
      E = dadd(E, pen);

      if (P.cap_trial != nullptr && P.mode != MODE_EVAL) {
        for (int i = 0; i < D; ++i) P.cap_trial[c * D + i] = row[i];
        P.cap_energy[c] = E;
      }
      if (P.mode == MODE_EVAL) {
        P.e_next[c] = E;
        continue;
      }
      const double e_old = (P.mode == MODE_GEN0) ? CUDART_INF : __ldg(P.e_cur + c);
      const bool acc = E < e_old;
      double* __restrict__ nrow = P.pop_next + c * D;
      if (acc || P.mode == MODE_GEN0) {
        for (int i = 0; i < D; ++i) nrow[i] = row[i];
      } else {
        const double* __restrict__ prow = P.pop_cur + c * D;
        for (int i = 0; i < D; ++i) nrow[i] = ld_stream1(prow + i);
      }
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
    __syncwarp();  // s_trial / s_sum are reused by the next batch
  }

  if (P.mode == MODE_EVAL) return;

  // warp argmin (lowest index on ties) and counters, then per-CTA partials
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
