// K2e: exponential-crossover generation kernel for long rows, TMA bulk staging.
//
// Nine of the reference's ten strategies run the exponential loop (mystic/strategy.py:48-56): a
// trial differs from its parent only inside a window of L consecutive (wrapping) dimensions that
// starts at `nstart`, E[L] = CR(1-CR^D)/(1-CR) (9 at CR = 0.9).  Per candidate the kernel therefore
// moves the parent row in, the selected row out, and k short donor segments:
//   * a GROUP of GW warps owns one candidate at a time and a private ring of `stages` buffers;
//   * the group's leader warp makes the draws (donors, start index, window length; same Philox
//     counters as every other kernel) and issues `cp.async.bulk` copies onto the stage's mbarrier:
//     the parent row (8*D bytes) and, per donor, the 16-byte aligned segment(s) covering the
//     window (two when it wraps), compacted into a `cap`-element buffer;
//   * the trial is built IN PLACE in the parent buffer; the parent's window values go to an undo
//     buffer, so a rejected trial (the common case) is restored by rewriting L elements;
//   * cost / penalty read the row from shared memory with the canonical reduction of
//     de_device.cuh, and the selected row leaves with one bulk shared->global copy.
// Windows longer than `cap` (probability < 1e-3 by the host's choice of cap) read their donor
// elements with plain loads and restore a rejected trial from the parent row in global memory.
//
// Semantics are those of de_step_kernel<BASE, XOVER_EXP, COST> (de_kernels.cuh); the same device
// functions produce the numbers, so both kernels agree bit for bit.
#pragma once
#include "de_tma.cuh"

namespace mb2 {

struct ExpStageMeta {  // written by the leader before the mbarrier arrive, read by the group after the wait
  int nstart, L, fits, a1, n1;
  int phases;  // rows of odd length: bit 0 the parent's phase, bit 1 + j donor j's (slot of element i = phase + i)
  int pad_[2];
  long long r[5];
};

// Rows of odd length: the 16-byte aligned cover of elements [lo, hi) of a row whose phase (row index & 1 = parity
// of its first element's index in the buffer) is `ph`: it starts at the largest s <= lo and ends at the smallest
// t >= hi with the parity of ph (s may be -1: the last element of the previous row; t may be D + 1: the first of the next)
__device__ __forceinline__ int odd_cover_start(int lo, int ph) { return lo - ((lo - ph) & 1); }
__device__ __forceinline__ int odd_cover_end(int hi, int ph) { return hi + ((hi - ph) & 1); }
static_assert(sizeof(ExpStageMeta) <= 128, "stage meta must fit its 128-byte slot");

__device__ __forceinline__ double mutant_rt(int base, double t, double b, const double (&p)[5], double F) {
  switch (base) {
    case BASE_BEST1: return mutant<BASE_BEST1>(t, b, p, F);
    case BASE_RAND1: return mutant<BASE_RAND1>(t, b, p, F);
    case BASE_RANDTOBEST1: return mutant<BASE_RANDTOBEST1>(t, b, p, F);
    case BASE_BEST2: return mutant<BASE_BEST2>(t, b, p, F);
    default: return mutant<BASE_RAND2>(t, b, p, F);
  }
}

// donors + start index for a run-time donor count; collective over the calling warp, same
// results as draw_donors<K>
__device__ __forceinline__ void draw_donors_warp_rt(const StepParams& P, int kdon, int64_t c, int lane, int64_t (&r)[5],
                                                    int* nstart, uint32_t* xword) {
  if (P.replay) {
#pragma unroll
    for (int j = 0; j < 5; ++j) r[j] = (j < kdon) ? (int64_t)__ldg(P.r_donors + c * 5 + j) : 0;
    *nstart = __ldg(P.r_nstart + c);
    return;
  }
  const u32x4 v = philox4x32_10((uint32_t)lane, (uint32_t)(P.global_offset + c), P.generation, STREAM_DONORS, P.key0, P.key1);
  const uint64_t even = ((uint64_t)v.x << 32) | v.y, odd = ((uint64_t)v.z << 32) | v.w;
  uint64_t w[6];
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    w[2 * b] = __shfl_sync(0xffffffffu, even, b);
    w[2 * b + 1] = __shfl_sync(0xffffffffu, odd, b);
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) r[j] = 0;
  switch (kdon) {
    case 2: resolve_donors<2>(w, c, P.np, P.dim, r, nstart, xword); break;
    case 3: resolve_donors<3>(w, c, P.np, P.dim, r, nstart, xword); break;
    case 4: resolve_donors<4>(w, c, P.np, P.dim, r, nstart, xword); break;
    default: resolve_donors<5>(w, c, P.np, P.dim, r, nstart, xword); break;
  }
}

// requires: cap even, 16-byte aligned population buffers, MODE_STEP or MODE_GEN0.  blockDim = groups * GW * 32 <= MAXT.
// Rows of ODD length (de_tma.cuh explains the phases): the row buffer holds the aligned cover of the parent (D + 1
// elements, the row at slot `phase`), every donor's window cover is aligned for THAT donor's phase (so the donors'
// segments start up to one element apart), and the selected row leaves by plain stores of the group.
template <int COST, int GW, int MAXT, bool ODD = false>  // ODD: rows of odd length (its own build: the even one pays nothing for the phases)
__global__ void __launch_bounds__(MAXT, 1)
    de_step_exp_tma_kernel(const __grid_constant__ StepParams P, int stages, int cap, int base, int kdon) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kTmaMaxGroups * kTmaMaxStages];
  __shared__ double s_scratch[kTmaMaxGroups * 16];
  __shared__ double sh_e[kTmaMaxGroups];
  __shared__ int64_t sh_i[kTmaMaxGroups];
  __shared__ unsigned long long sh_ninf[kTmaMaxGroups], sh_nacc[kTmaMaxGroups];

  const int D = P.dim;
  constexpr bool odd = ODD;  // (the host picks the build by D & 1)
  const unsigned row_bytes = (unsigned)D * 8u;
  const unsigned row_stride = (row_bytes + (odd ? 8u : 0u) + 127u) & ~127u;  // keep every buffer 128-B aligned (32-bit offsets: cheap to rematerialise)
  const unsigned seg_stride = (unsigned)cap * 8u;                        // one donor's (or the undo) buffer
  const unsigned stage_stride = (row_stride + (unsigned)(kdon + 1) * seg_stride + 128u + 127u) & ~127u;
  constexpr int gthreads = GW * 32;
  const int group = threadIdx.x / gthreads;
  const int ngroups = blockDim.x / gthreads;
  const int tig = threadIdx.x - group * gthreads;
  const int lane = threadIdx.x & 31;
  const bool leader = (tig == 0);
  const Group<GW> grp = {tig, tig >> 5, lane, 1 + group, s_scratch + group * 16};
  const bool needs_best = (base == BASE_BEST1 || base == BASE_RANDTOBEST1 || base == BASE_BEST2);

  double* s_best = reinterpret_cast<double*>(smem_raw);
  unsigned char* s_ring = smem_raw + (row_stride + (unsigned)(group * stages) * stage_stride);
  uint64_t* bars = s_bar + group * kTmaMaxStages;

  if (needs_best)
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_best[i] = __ldg(P.best_x + i);
  if (leader) {
    for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t g0 = (int64_t)blockIdx.x * ngroups + group;
  const int64_t ng = (int64_t)gridDim.x * ngroups;

  // stage layout: [row][undo cap][donor 0 cap]...[donor k-1 cap][meta]
  auto stage_row = [&](int s) { return reinterpret_cast<double*>(s_ring + (unsigned)s * stage_stride); };
  auto stage_undo = [&](int s) { return reinterpret_cast<double*>(s_ring + ((unsigned)s * stage_stride + row_stride)); };
  auto stage_don = [&](int s, int j) {
    return reinterpret_cast<double*>(s_ring + ((unsigned)s * stage_stride + row_stride + (unsigned)(1 + j) * seg_stride));
  };
  auto stage_meta = [&](int s) {
    return reinterpret_cast<ExpStageMeta*>(s_ring + ((unsigned)s * stage_stride + row_stride + (unsigned)(kdon + 1) * seg_stride));
  };

  // Draws of candidate c and its loads into stage s.  Called by the whole leader WARP of the group
  // (tig < 32): the draws are collective, lane 0 loads the parent row, lanes 1..k / 9..8+k the first /
  // wrapped segment of donor lane-1 / lane-9.
  // The group's rows are issued in the order g0, g0 + ng, ...: one Philox evaluation of the leader warp makes the
  // donor-stream blocks of the next 10 of them (lane l: block l % 3 of the row l / 3 ahead; three blocks cover
  // every donor count), handed out by shuffle -- same counters and words as draw_donors<K>.
  uint64_t dc_even = 0, dc_odd = 0;
  int dc_pos = 10;  // next cached row (10: the cache is empty)
  auto issue = [&](int64_t c, int s) {
    uint64_t* bar = bars + s;
    // the parent row does not depend on the draws: its copy flies while they are made
    const int php = odd ? (int)(c & 1) : 0;
    if (leader) {
      if (!odd) {
        mbar_expect_tx_only(bar, row_bytes);
        bulk_g2s(stage_row(s), P.pop_cur + c * D, row_bytes, bar);
      } else {
        // the cover of the last row of an odd number of rows would end past the buffer: two elements short + a plain load
        const bool short_cover = (php == 0) && (c == P.np - 1);
        const unsigned bytes = short_cover ? row_bytes - 8u : row_bytes + 8u;
        if (short_cover) stage_row(s)[D - 1] = __ldg(P.pop_cur + c * D + (D - 1));  // (published by the arrive below)
        mbar_expect_tx_only(bar, bytes);
        bulk_g2s(stage_row(s), P.pop_cur + c * D - php, bytes, bar);
      }
    }
    if (P.mode == MODE_GEN0) {  // generation 0: the row is only clipped and evaluated -- no draws, no donors
      if (leader) {
        ExpStageMeta* m = stage_meta(s);
        m->nstart = 0;
        m->L = 0;
        m->fits = 0;
        m->a1 = 0;
        m->n1 = 0;
        m->phases = php;
        mbar_expect_tx(bar, 0u);  // the arrive
      }
      __syncwarp();
      return;
    }
    int64_t r[5];
    int nstart;
    uint32_t xword = 0;
    if (P.replay) {
      draw_donors_warp_rt(P, kdon, c, lane, r, &nstart, &xword);
    } else {
      if (dc_pos == 10) {
        const int64_t cr = c + (int64_t)(lane / 3) * ng;  // rows past the end (and lanes 30, 31): evaluated, never handed out
        const u32x4 v = philox4x32_10((uint32_t)(lane % 3), (uint32_t)(P.global_offset + cr), P.generation, STREAM_DONORS, P.key0, P.key1);
        dc_even = ((uint64_t)v.x << 32) | v.y;
        dc_odd = ((uint64_t)v.z << 32) | v.w;
        dc_pos = 0;
      }
      uint64_t w[6];
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        w[2 * b] = __shfl_sync(0xffffffffu, dc_even, 3 * dc_pos + b);
        w[2 * b + 1] = __shfl_sync(0xffffffffu, dc_odd, 3 * dc_pos + b);
      }
      ++dc_pos;
#pragma unroll
      for (int j = 0; j < 5; ++j) r[j] = 0;
      switch (kdon) {
        case 2: resolve_donors<2>(w, c, P.np, D, r, &nstart, &xword); break;
        case 3: resolve_donors<3>(w, c, P.np, D, r, &nstart, &xword); break;
        case 4: resolve_donors<4>(w, c, P.np, D, r, &nstart, &xword); break;
        default: resolve_donors<5>(w, c, P.np, D, r, &nstart, &xword); break;
      }
    }
    const int L = P.replay ? exp_length_replay(P.r_uni + c * (int64_t)(D + 1), D, P.CR, lane) : exp_length_v2(xword, P.xlen, D);
    // 16-byte aligned cover of the window [nstart, nstart+L) mod D: [a1, a1+n1) and, wrapped, [0, n2)
    int a1 = nstart & ~1;
    const int end = nstart + L;
    int n1 = (end < D ? ((end + 1) & ~1) : D) - a1;
    int n2 = (end > D) ? ((end - D + 1) & ~1) : 0;
    if (L >= D) {  // the whole row
      a1 = 0;
      n1 = D;
      n2 = 0;
    }
    bool fits = (L > 0) && (n1 + n2 <= cap);
    if (odd) {
      // per donor: first segment = cover of [lo1, hi1), wrapped segment = cover of [0, hi2); lo1 is also the
      // threshold `a1` that tells the two apart when an element is looked up
      const int lo1 = (L >= D) ? 0 : nstart, hi1 = (end < D) ? end : D, hi2 = (L < D && end > D) ? end - D : 0;
      int phases = php, total = 0;
      bool ok = L > 0;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        if (j < kdon) {
          const int ph = (int)(r[j] & 1);
          phases |= ph << (1 + j);
          const int t1 = odd_cover_end(hi1, ph);
          const int m1 = t1 - odd_cover_start(lo1, ph), m2 = hi2 > 0 ? odd_cover_end(hi2, ph) + ph : 0;
          total += m1 + m2;
          ok = ok && (m1 + m2 <= cap) && !(r[j] == P.np - 1 && (t1 > D || (hi2 > 0 && odd_cover_end(hi2, ph) > D)));  // (a cover past the end of the buffer)
        }
      }
      fits = ok;
      if (leader) {
        ExpStageMeta* m = stage_meta(s);
        m->nstart = nstart;
        m->L = L;
        m->fits = fits ? 1 : 0;
        m->a1 = lo1;
        m->n1 = hi1;
        m->phases = phases;
#pragma unroll
        for (int j = 0; j < 5; ++j) m->r[j] = r[j];
        mbar_expect_tx(bar, fits ? (unsigned)total * 8u : 0u);  // the arrive: meta is published with it
      }
      __syncwarp();
      if (fits) {
        const int j1 = lane - 1, j2 = lane - 9;
        if (j1 >= 0 && j1 < kdon) {
          const int64_t rj = (j1 == 0) ? r[0] : (j1 == 1) ? r[1] : (j1 == 2) ? r[2] : (j1 == 3) ? r[3] : r[4];
          const int ph = (int)(rj & 1), s1 = odd_cover_start(lo1, ph);
          bulk_g2s(stage_don(s, j1), P.pop_cur + rj * D + s1, (unsigned)(odd_cover_end(hi1, ph) - s1) * 8u, bar);
        }
        if (hi2 > 0 && j2 >= 0 && j2 < kdon) {
          const int64_t rj = (j2 == 0) ? r[0] : (j2 == 1) ? r[1] : (j2 == 2) ? r[2] : (j2 == 3) ? r[3] : r[4];
          const int ph = (int)(rj & 1), m1 = odd_cover_end(hi1, ph) - odd_cover_start(lo1, ph);
          bulk_g2s(stage_don(s, j2) + m1, P.pop_cur + rj * D - ph, (unsigned)(odd_cover_end(hi2, ph) + ph) * 8u, bar);
        }
      }
      return;
    }
    if (leader) {
      ExpStageMeta* m = stage_meta(s);
      m->nstart = nstart;
      m->L = L;
      m->fits = fits ? 1 : 0;
      m->a1 = a1;
      m->n1 = n1;
      m->phases = 0;
#pragma unroll
      for (int j = 0; j < 5; ++j) m->r[j] = r[j];
      mbar_expect_tx(bar, fits ? (unsigned)(kdon * (n1 + n2)) * 8u : 0u);  // the arrive: meta is published with it
    }
    __syncwarp();
    if (fits) {
      const int j1 = lane - 1, j2 = lane - 9;
      if (j1 >= 0 && j1 < kdon) {
        const int64_t rj = (j1 == 0) ? r[0] : (j1 == 1) ? r[1] : (j1 == 2) ? r[2] : (j1 == 3) ? r[3] : r[4];
        bulk_g2s(stage_don(s, j1), P.pop_cur + rj * D + a1, (unsigned)n1 * 8u, bar);
      }
      if (n2 > 0 && j2 >= 0 && j2 < kdon) {
        const int64_t rj = (j2 == 0) ? r[0] : (j2 == 1) ? r[1] : (j2 == 2) ? r[2] : (j2 == 3) ? r[3] : r[4];
        bulk_g2s(stage_don(s, j2) + n1, P.pop_cur + rj * D, (unsigned)n2 * 8u, bar);
      }
    }
  };

  if (tig < 32) {
    for (int s = 0; s < stages; ++s) {
      const int64_t c = g0 + (int64_t)s * ng;
      if (c < P.np) issue(c, s);
    }
  }

  double g_best_e = CUDART_INF;
  int64_t g_best_idx = INT64_MAX;
  unsigned long long g_ninf = 0, g_nacc = 0;
  const bool uni_b = P.bounds_uniform != 0;
  const bool gen0 = (P.mode == MODE_GEN0);
  // rows of this generation are known to lie inside the strict range: only the window can leave it
  const bool window_only = gen0 || (P.bounds_mode == BOUNDS_NONE) || (P.rows_in_range != 0);

  int s = 0;
  unsigned parity = 0;
  for (int64_t c = g0; c < P.np; c += ng) {
    double* s_row = stage_row(s);
    double* s_undo = stage_undo(s);
    const ExpStageMeta* meta = stage_meta(s);
    const double e_old = gen0 ? CUDART_INF : __ldg(P.e_cur + c);  // in flight while the group waits for its rows

    group_wait_stage(grp, bars + s, parity);
    const int nstart = meta->nstart, L = meta->L, a1 = meta->a1, n1 = meta->n1;
    const bool fits = meta->fits != 0;
    // rows of odd length: the row starts at slot `phase` of its buffer; donor j's segments are the covers for ITS
    // phase (a1 / n1 carry the first segment's first element and end: [a1, n1), the wrapped one starts at element 0)
    const int phases = odd ? meta->phases : 0;
    if (odd) s_row += phases & 1;

    // ---- trial vector in place (strategy.py:48-56): window element w is dimension (nstart + w) mod D
    bool oob = false;
    for (int w = tig; w < L; w += gthreads) {
      int i = nstart + w;
      if (i >= D) i -= D;
      const double t = s_row[i];
      double p[5] = {0., 0., 0., 0., 0.};
      if (fits && odd) {
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (j < kdon) {
            const int ph = (phases >> (1 + j)) & 1, s1 = odd_cover_start(a1, ph);
            p[j] = stage_don(s, j)[(i >= a1) ? (i - s1) : (odd_cover_end(n1, ph) - s1 + i + ph)];
          }
      } else if (fits) {
        const int off = (i >= a1) ? (i - a1) : (n1 + i);
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (j < kdon) p[j] = stage_don(s, j)[off];
      } else {
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (j < kdon) p[j] = __ldg(P.pop_cur + meta->r[j] * D + i);
      }
      double x = mutant_rt(base, t, needs_best ? s_best[i] : 0.0, p, P.F);
      if (window_only && P.bounds_mode != BOUNDS_NONE) {
        const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
        const double h = uni_b ? P.hi_s : __ldg(P.hi + i);
        if (P.bounds_mode == BOUNDS_CLAMP) {
          x = (x > l) ? x : l;  // python max(lo, x)
          x = (x < h) ? x : h;  // python min(hi, x)
        } else {
          oob |= (x < l) || (x > h);
        }
      }
      if (fits) s_undo[w] = t;
      s_row[i] = x;
    }
    if (gen0 && P.bounds_mode != BOUNDS_NONE) {
      // generation 0 clips the initial population into the strict range (differential_evolution.py:516-520)
      for (int i = tig; i < D; i += gthreads) {
        const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
        const double h = uni_b ? P.hi_s : __ldg(P.hi + i);
        const double x = fmin(fmax(s_row[i], l), h);
        oob |= (x < l) || (x > h);
        s_row[i] = x;
      }
    }
    if (!window_only) {
      // rows this context has not checked against the range: the reference clamps / tests every
      // element of the trial, untouched ones included
      group_sync(grp);
      for (int i = tig; i < D; i += gthreads) {
        double x = s_row[i];
        const double l = uni_b ? P.lo_s : __ldg(P.lo + i);
        const double h = uni_b ? P.hi_s : __ldg(P.hi + i);
        if (P.bounds_mode == BOUNDS_CLAMP) {
          x = (x > l) ? x : l;
          x = (x < h) ? x : h;
          s_row[i] = x;
        } else {
          oob |= (x < l) || (x > h);
        }
      }
    }
    oob = group_any(grp, oob);  // also the barrier that publishes the trial vector to the group

    double E = CUDART_INF;
    if (!oob) E = group_cost<COST, GW>(grp, s_row, D, P.cost);
    E = dadd(E, group_penalty(grp, s_row, D, P.pen));

    if (P.cap_trial != nullptr) {
      for (int i = tig; i < D; i += gthreads) P.cap_trial[c * D + i] = s_row[i];
      if (leader) P.cap_energy[c] = E;
    }

    const bool acc = E < e_old;
    if (!acc && !gen0 && !P.sparse) {  // (generation 0 keeps the clipped row whatever its energy; sparse: a rejected row is not stored)
      // every thread has read the trial (costs without a group reduction and the capture end
      // without a barrier) before a rejected one is undone
      if (COST == COST_CORANA || COST == COST_ZIMMERMANN || P.cap_trial != nullptr) group_sync(grp);
      // ---- rejected: put the parent back (:603-609 keeps the old row)
      if (!window_only && P.bounds_mode == BOUNDS_CLAMP) {
        for (int i = tig; i < D; i += gthreads) s_row[i] = __ldg(P.pop_cur + c * D + i);
      } else {
        for (int w = tig; w < L; w += gthreads) {
          int i = nstart + w;
          if (i >= D) i -= D;
          s_row[i] = fits ? s_undo[w] : __ldg(P.pop_cur + c * D + i);
        }
      }
    }
    // generic-proxy writes of the row -> visible to the async proxy; all reads of this stage done
    fence_proxy_async();
    group_sync(grp);
    if (acc) {
      ++g_nacc;
      if (E < g_best_e) {
        g_best_e = E;
        g_best_idx = c;
      }
    }
    if (isinf(E)) ++g_ninf;
    if (odd && (acc || gen0 || !P.sparse)) {
      // rows of odd length leave by plain stores of the whole group (8-byte aligned on both sides)
      double* __restrict__ nrow = P.pop_next + c * D;
      for (int i = tig; i < D; i += gthreads) nrow[i] = s_row[i];
      group_sync(grp);  // the stage has been read before it is refilled
    }
    if (tig < 32) {
      if (leader) {
        const bool store = !odd && (acc || gen0 || !P.sparse);
        if (store) {
          bulk_s2g(P.pop_next + c * D, s_row, row_bytes);
          bulk_commit();
        }
        P.e_next[c] = acc ? E : e_old;
        if (store) bulk_wait_read0();  // the store has finished reading the stage
      }
      __syncwarp();
      // refill this stage with the row it serves next
      const int64_t cn = c + (int64_t)stages * ng;
      if (cn < P.np) issue(cn, s);
    }
    if (++s == stages) {
      s = 0;
      parity ^= 1u;
    }
  }
  if (leader) {
    bulk_wait_all0();
    sh_e[group] = g_best_e;
    sh_i[group] = g_best_idx;
    sh_ninf[group] = g_ninf;
    sh_nacc[group] = g_nacc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double be = sh_e[0];
    int64_t bi = sh_i[0];
    unsigned long long ni = sh_ninf[0], na = sh_nacc[0];
    for (int w = 1; w < ngroups; ++w) {
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
