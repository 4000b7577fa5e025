// The fused differential-evolution generation kernel (K2 of SURVEY.md section 7) and its
// 1-CTA argmin epilogue (K3).  One warp owns one candidate row at a time; a persistent grid
// strides over the shard so that the per-CTA (energy, index) partials stay few.
//
// Reference semantics reproduced (paths relative to the mystic repository root):
//   trial generation   mystic/differential_evolution.py:574-582 + mystic/strategy.py:34-311
//   bounds             mystic/tools.py:420-426 (reject -> inf), mystic/symbolic.py:1139-1144 (clamp)
//   cost + penalty     mystic/differential_evolution.py:515-525, mystic/tools.py:386-389
//   selection + best   mystic/differential_evolution.py:603-613
#pragma once
#include "de_device.cuh"

namespace mb2 {

enum : int { BASE_BEST1 = 0, BASE_RAND1 = 1, BASE_RANDTOBEST1 = 2, BASE_BEST2 = 3, BASE_RAND2 = 4 };
enum : int { XOVER_EXP = 0, XOVER_BIN = 1 };
enum : int { MODE_STEP = 0, MODE_GEN0 = 1, MODE_EVAL = 2 };
enum : int { BOUNDS_NONE = 0, BOUNDS_REJECT = 1, BOUNDS_CLAMP = 2 };

constexpr int kWarpsPerCtaMax = 8;

// Device-resident multi-generation stepping: the termination conditions that
// only read the energy history and the counters (mystic/termination.py:181-240, 320-342, 386-408,
// and the solver's own limits, abstract_solver.py:703-711) are evaluated by the finalize kernel;
// once one is met `stop` is raised and every later launch of the batch returns immediately.
constexpr int kHistMax = 1024;
enum : int { TERM_NONE = 0, TERM_VTR = 1, TERM_COG = 2, TERM_NCOG = 3, TERM_VTRCOG = 4, TERM_LIMITS = 5 };

struct RunState {
  int stop;
  int kind;
  int lookback;  // "generations" argument of the condition
  int pad_;
  double tol, target, gtol;
  long long term_generations, term_evaluations;  // EvaluationLimits condition (<0: none)
  long long max_generations, max_evaluations;    // solver._maxiter / _maxfun (<0: none)
  long long generations;                         // solver.generations
  long long evaluations;                         // solver._fcalls[0]
  long long np_eval;                             // evaluations per generation before skipping inf
  long long hist_count;                          // len(energy_history)
  long long log_count, log_cap;
  double* log;                                   // [log_cap, D+4]: best_e, best_idx, n_inf, n_acc, x[D]
  double hist0;                                  // energy_history[0] (python hist[-0])
  double hist[kHistMax];                         // ring of the last energies
};

struct StepParams {
  const int* stop;  // non-null in batch mode: return at once when *stop != 0
  // population (double buffered) and energies
  const double* pop_cur;
  double* pop_next;
  const double* e_cur;
  double* e_next;
  const double* best_x;  // bestSolution of the previous generation [D]
  const double* lo;      // [D] (bounds_mode != NONE)
  const double* hi;
  int64_t np;             // rows on this shard
  int64_t global_offset;  // global index of row 0
  int dim;
  int dim_pad;  // dim rounded up to a multiple of 4
  int mode;     // MODE_*
  // 1 (MODE_STEP of the staged kernels): only ACCEPTED trials are written to pop_next (e_next gets every row's
  // energy); de_commit_accepted_kernel then moves them into pop_cur in place and the buffers are not swapped --
  // the reference's own layout (population + trialSolution, differential_evolution.py:603-609), and a rejected
  // row costs no write
  int sparse;
  int bounds_mode;
  int bounds_uniform;  // 1: every dimension has the same [lo_s, hi_s] (no per-dimension loads)
  int rows_in_range;   // 1: every row of pop_cur is known to lie inside [lo, hi] (generation 0 clipped it)
  double lo_s, hi_s;
  double F, CR;
  // RNG
  int replay;  // 1: consume recorded draws
  uint32_t key0, key1, generation;
  uint32_t thr16;  // binomial crossover: threshold on the 16-bit fields of stream 1 (philox.cuh, protocol 2)
  ExpLength xlen;  // exponential crossover: length table of the geometric draw
  const int32_t* r_donors;  // [np,5]
  const int32_t* r_nstart;  // [np]
  const double* r_uni;      // [np, dim+1]
  CostParams cost;
  PenaltyParams pen;
  // per-CTA partials
  double* part_e;
  int64_t* part_idx;
  unsigned long long* part_ninf;
  unsigned long long* part_nacc;
  // optional capture
  double* cap_trial;
  double* cap_energy;
};

template <int BASE>
struct NumDonors {
  static constexpr int value = (BASE == BASE_BEST1 || BASE == BASE_RANDTOBEST1) ? 2 : (BASE == BASE_RAND1 ? 3 : (BASE == BASE_BEST2 ? 4 : 5));
};

// mutant value for one dimension, reference operation order, no FMA.
//   parent t, best b, donors p0..p4 (strategy.py formulas)
template <int BASE>
__device__ __forceinline__ double mutant(double t, double b, const double (&p)[5], double F) {
  if (BASE == BASE_BEST1) return dadd(b, dmul(F, dsub(p[0], p[1])));
  if (BASE == BASE_RAND1) return dadd(p[0], dmul(F, dsub(p[1], p[2])));
  if (BASE == BASE_RANDTOBEST1) return dadd(t, dadd(dmul(F, dsub(b, t)), dmul(F, dsub(p[0], p[1]))));
  if (BASE == BASE_BEST2) return dadd(b, dmul(F, dsub(dsub(dadd(p[0], p[1]), p[2]), p[3])));
  return dadd(p[0], dmul(F, dsub(dsub(dadd(p[1], p[2]), p[3]), p[4])));
}

// Exponential crossover length L in [0, D]: number of leading draws < CR, capped at D
// (strategy.py:48-56: `if random.random() >= CR or i == D: break`).  Warp-cooperative.
__device__ __forceinline__ int exp_length_replay(const double* __restrict__ uni, int D, double CR, int lane) {
  for (int base = 0; base < D; base += 32) {
    const int i = base + lane;
    bool fail = true;
    if (i < D) fail = !(__ldg(uni + i) < CR);  // NaN (not drawn) -> stop
    const unsigned m = __ballot_sync(0xffffffffu, fail);
    if (m) return base + __ffs(m) - 1;
  }
  return D;
}

// Same length for ONE thread's own candidate (thread- and group-per-row kernels).  Philox mode: the
// geometric draw from `xword`, the low half of the start-index word (philox.cuh, protocol 2).
__device__ __forceinline__ int exp_length_thread(const StepParams& P, uint32_t xword, int64_t c, int D) {
  if (P.replay) {
    int L = 0;
    const double* u = P.r_uni + c * (int64_t)(D + 1);
    while (L < D && (__ldg(u + L) < P.CR)) ++L;  // NaN (not drawn) -> stop
    return L;
  }
  return exp_length_v2(xword, P.xlen, D);
}

template <int BASE, int XOVER, int COST, bool VEC>
__global__ void __launch_bounds__(kWarpsPerCtaMax * 32)
    de_step_kernel(const __grid_constant__ StepParams P) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(16) double smem[];
  constexpr int K = NumDonors<BASE>::value;
  constexpr bool kNeedsBest = (BASE == BASE_BEST1 || BASE == BASE_RANDTOBEST1 || BASE == BASE_BEST2);
  const int D = P.dim;
  const int Dp = P.dim_pad;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  __shared__ double s_scratch[kWarpsPerCtaMax * 16];
  double* s_best = smem;                              // [Dp]
  double* s_trial = smem + Dp + (size_t)warp * Dp;   // [Dp] per warp
  const Group<1> grp = {lane, 0, lane, 0, s_scratch + warp * 16};  // one warp per row

  if (P.mode == MODE_STEP && kNeedsBest) {
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_best[i] = __ldg(P.best_x + i);
  }
  __syncthreads();

  double w_best_e = CUDART_INF;
  int64_t w_best_idx = INT64_MAX;
  unsigned long long w_ninf = 0, w_nacc = 0;

  const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
  const int64_t nw = (int64_t)gridDim.x * wpc;
  for (int64_t c = gw; c < P.np; c += nw) {
    const double* __restrict__ prow = P.pop_cur + c * D;
    bool oob = false;

    if (P.mode == MODE_STEP) {
      // ---- random draws for this candidate
      int64_t r[5] = {0, 0, 0, 0, 0};
      int nstart;
      uint32_t xword = 0;
      const uint32_t cand_g = (uint32_t)(P.global_offset + c);
      if (P.replay) {
#pragma unroll
        for (int j = 0; j < K; ++j) r[j] = __ldg(P.r_donors + c * 5 + j);
        nstart = __ldg(P.r_nstart + c);
      } else {
        draw_donors<K>(P.key0, P.key1, P.generation, cand_g, c, P.np, D, r, &nstart, &xword);
      }
      int L = 0;
      if (XOVER == XOVER_EXP) {
        L = P.replay ? exp_length_replay(P.r_uni + c * (int64_t)(D + 1), D, P.CR, lane) : exp_length_v2(xword, P.xlen, D);
      }
      const double* __restrict__ d0 = P.pop_cur + r[0] * D;
      const double* __restrict__ d1 = P.pop_cur + r[1] * D;
      const double* __restrict__ d2 = P.pop_cur + r[2] * D;
      const double* __restrict__ d3 = P.pop_cur + r[3] * D;
      const double* __restrict__ d4 = P.pop_cur + r[4] * D;

      // ---- trial vector: 4 consecutive dimensions per lane per trip
      for (int i0 = lane * 4; i0 < D; i0 += 128) {
        double x[4];
        load4<VEC>(prow, i0, D, x);
        bool cross[4];
        if (XOVER == XOVER_BIN) {
          // strategy.py:79-85: i==n or random() < CR, one draw per dimension
          if (P.replay) {
            const double* u = P.r_uni + c * (int64_t)(D + 1) + i0;
#pragma unroll
            for (int e = 0; e < 4; ++e) cross[e] = (i0 + e < D) && ((i0 + e == nstart) || (__ldg(u + e) < P.CR));
          } else {
            // block i0/8 decides dimensions 8q..8q+7; this lane's four are its low or high pair of words
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
          if (K > 2) load4<VEC>(d2, i0, D, q[2]);
          if (K > 3) load4<VEC>(d3, i0, D, q[3]);
          if (K > 4) load4<VEC>(d4, i0, D, q[4]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (cross[e]) {
              double p[5];
#pragma unroll
              for (int j = 0; j < 5; ++j) p[j] = (j < K) ? q[j][e] : 0.0;
              const double b = kNeedsBest ? s_best[i0 + e] : 0.0;
              x[e] = mutant<BASE>(x[e], b, p, P.F);
            }
          }
        }
        // ---- bounds on the trial (constraints applied before the cost, :582)
        if (P.bounds_mode != BOUNDS_NONE) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (i0 + e < D) {
              const double l = P.bounds_uniform ? P.lo_s : __ldg(P.lo + i0 + e);
              const double h = P.bounds_uniform ? P.hi_s : __ldg(P.hi + i0 + e);
              if (P.bounds_mode == BOUNDS_CLAMP) {
                x[e] = (x[e] > l) ? x[e] : l;  // python max(lo, x)
                x[e] = (x[e] < h) ? x[e] : h;  // python min(hi, x)
              } else
                oob |= (x[e] < l) || (x[e] > h);  // after a clamp nothing is out of range
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) s_trial[i0 + e] = x[e];
      }
    } else {
      // MODE_GEN0: evaluate the initial population, clipped into the strict range
      //            (differential_evolution.py:516-520, numpy clip = min(max(x,lo),hi))
      // MODE_EVAL: evaluate arbitrary rows as given
      for (int i0 = lane * 4; i0 < D; i0 += 128) {
        double x[4];
        load4<VEC>(prow, i0, D, x);
        if (P.bounds_mode != BOUNDS_NONE) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (i0 + e < D) {
              const double l = P.bounds_uniform ? P.lo_s : __ldg(P.lo + i0 + e);
              const double h = P.bounds_uniform ? P.hi_s : __ldg(P.hi + i0 + e);
              if (P.mode == MODE_GEN0) x[e] = fmin(fmax(x[e], l), h);
              oob |= (x[e] < l) || (x[e] > h);
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) s_trial[i0 + e] = x[e];
      }
    }
    oob = __any_sync(0xffffffffu, oob);
    __syncwarp();

    // ---- energy = (inf if out of range else cost(x)) + penalty(x)
    double E = CUDART_INF;
    if (!oob) E = group_cost<COST>(grp, s_trial, D, P.cost);
    E = dadd(E, group_penalty(grp, s_trial, D, P.pen));

    if (P.cap_trial != nullptr && P.mode != MODE_EVAL) {
      for (int i = lane; i < D; i += 32) P.cap_trial[c * D + i] = s_trial[i];
      if (lane == 0) P.cap_energy[c] = E;
    }

    if (P.mode == MODE_EVAL) {
      if (lane == 0) P.e_next[c] = E;
    } else {
      // ---- greedy selection (:603-609) and write of the next generation
      const double e_old = (P.mode == MODE_GEN0) ? CUDART_INF : __ldg(P.e_cur + c);
      const bool acc = E < e_old;
      double* __restrict__ nrow = P.pop_next + c * D;
      if (acc || P.mode == MODE_GEN0) {
        for (int i0 = lane * 4; i0 < D; i0 += 128) {
          double x[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = s_trial[i0 + e];
          store4<VEC>(nrow, i0, D, x);
        }
      } else if (!P.sparse) {  // (sparse: a rejected row stays where it is, de_commit_accepted_kernel moves the accepted ones)
        for (int i0 = lane * 4; i0 < D; i0 += 128) {
          double x[4];
          load4<VEC>(prow, i0, D, x);
          store4<VEC>(nrow, i0, D, x);
        }
      }
      if (lane == 0) P.e_next[c] = acc ? E : e_old;
      // generation best among accepted trials; c ascends within the warp, strict '<' keeps
      // the lowest index (:610-613)
      if (acc) {
        ++w_nacc;
        if (E < w_best_e) {
          w_best_e = E;
          w_best_idx = c;
        }
      }
      if (isinf(E)) ++w_ninf;
    }
    __syncwarp();  // s_trial is reused by the next row
  }

  if (P.mode == MODE_EVAL) return;

  // ---- per-CTA (energy, index) argmin, lowest index on ties; counters
  __shared__ double sh_e[kWarpsPerCtaMax];
  __shared__ int64_t sh_i[kWarpsPerCtaMax];
  __shared__ unsigned long long sh_ninf[kWarpsPerCtaMax], sh_nacc[kWarpsPerCtaMax];
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

// device-resident solver state shared by the kernels of one context
struct DeviceState {
  double best_e;       // bestEnergy
  int64_t best_idx;    // GLOBAL index of the member that last became best (-1: none)
  int64_t n_inf;       // of the last generation
  int64_t n_acc;
  int64_t generation;
};

}  // namespace mb2
