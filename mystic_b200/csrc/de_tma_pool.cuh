// K2p: K2a's generation step (Best1Bin, whole donor rows, de_tma.cuh) for IN-PLACE generations of long rows, with the
// row buffers of an SM in ONE pool and a producer warp.
//
// K2a gives every group its own stage of three row buffers, refilled when the group is done with its row; at BASELINE
// config 2's shape (8 KB rows) shared memory holds one stage per group, so a group's buffers idle while it computes and
// its loads' latency is fully exposed: the memory system delivers this read pattern in 0.199 ms (tools/gather_peak ...
// rows), the kernel needs 0.27 ms.  In place a rejected parent is never stored, which makes the buffers' lifetimes
// unequal: the parent's and the second donor's rows are dead once the trial vector stands in the first donor's buffer,
// that one lives until the row is selected.  So here
//   * the buffers form two rings -- P PAIRS (parent + second donor) and Q first-donor buffers, 2P + Q = what fits --
//     and row k of the CTA (rows blockIdx.x, + gridDim.x, ...) uses pair k mod P and buffer k mod Q;
//   * a PRODUCER warp walks the CTA's rows in order: draws (the same Philox counters as every kernel), waits until
//     pair k mod P and buffer k mod Q have been released by the rows that used them last (mbarriers the consumers
//     arrive on), and issues the three `cp.async.bulk` copies onto the row's `full` mbarrier;
//   * consumer GROUP g (GW warps) takes rows g, g + G, ...: waits for `full`, builds the trial in the first donor's
//     buffer, RELEASES THE PAIR, evaluates cost / penalty, selects, stores an accepted trial, releases the buffer.
// Rows in flight are bounded by the pool, not by the number of groups, and the consumers carry no draws or requests.
// Semantics and numbers are K2a's (the same device functions): bit-identical populations and energies.
#pragma once
#include <type_traits>

#include "de_tma.cuh"

#ifndef MB2_POOL_EARLY_DRAWS
#define MB2_POOL_EARLY_DRAWS 1
#endif

namespace mb2 {

constexpr int kPoolMaxGroups = 24;   // consumer groups (groups of several warps: named barriers 1..15)
constexpr int kPoolMaxRing = 64;     // P, Q <= 64

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// requires: D even, MODE_STEP, StepParams::sparse (in place), 16-byte aligned buffers.  blockDim = (G * GW + nprod) * 32: the last nprod (1 or 2)
// warps are producers.
template <int COST, int GW, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
    de_step_best1bin_pool_kernel(const __grid_constant__ StepParams P, int npairs_ring, int nd0_ring, int nprod) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kPoolMaxRing], s_pair_free[kPoolMaxRing], s_d0_free[kPoolMaxRing];
  __shared__ double s_scratch[kPoolMaxGroups * 16];
  __shared__ int2 s_nstart[kPoolMaxRing];  // {crossover start index, row number k} of the row in each first-donor slot
  __shared__ double sh_e[kPoolMaxGroups];
  __shared__ int64_t sh_i[kPoolMaxGroups];
  __shared__ unsigned long long sh_ninf[kPoolMaxGroups], sh_nacc[kPoolMaxGroups];

  const int D = P.dim;
  const unsigned row_bytes = (unsigned)D * 8u;
  const unsigned row_stride = (row_bytes + 127u) & ~127u;
  constexpr int gthreads = GW * 32;
  const int ngroups = (blockDim.x / 32 - nprod) / GW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp >= ngroups * GW;
  const int PR = npairs_ring, QR = nd0_ring;

  double* s_best = reinterpret_cast<double*>(smem_raw);
  unsigned char* s_pairs = smem_raw + row_stride;                              // [PR][2] rows
  unsigned char* s_d0s = s_pairs + (unsigned)(2 * PR) * row_stride;            // [QR] rows

  for (int i = threadIdx.x; i < D; i += blockDim.x) s_best[i] = __ldg(P.best_x + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < QR; ++i) {
      mbar_init(s_full + i, nprod);  // every producer warp arrives with the bytes of its copies
      mbar_init(s_d0_free + i, 1);
    }
    for (int i = 0; i < PR; ++i) mbar_init(s_pair_free + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t c0 = blockIdx.x, cstep = gridDim.x;   // the CTA's rows: c0 + k * cstep
  const int64_t nrows = (P.np > c0) ? (P.np - c0 + cstep - 1) / cstep : 0;

  if (producer) {
    // ---- producer warp: draws + copy requests, in row order.  A batch of 16 rows at a time: ONE Philox evaluation of
    // the warp makes their donor / start-index words (K2a's cache: lane l evaluates block l & 1 of row l >> 1), lane j
    // gathers the three words of row j and resolves ITS row's donors -- all rows of the batch in parallel -- and then
    // lane 0 waits for each row's buffers in turn and issues its copies, so the serial chain per row is five shuffles, two
    // mbarrier tests, one expect_tx and three requests (the per-row resolve of a warp-uniform loop made the producer the
    // bottleneck: its ~150 instructions per row compete with 16 consumer warps for issue slots).
    // Two producer warps (short rows, where one warp's serial chain per row would bound the kernel) split the COPIES, not
    // the rows: the first serves the pair ring (parent + second donor), the second the first-donor ring and the row's
    // tag; every ring then has ONE sequential writer, which is what keeps a parity test from passing a phase early.  The
    // first also waits for the row's first-donor slot (a non-destructive test): its arrive on `full` must not fall into
    // the slot's previous phase.
    const int pw = warp - ngroups * GW;
    const bool do_pairs = (pw == 0), do_d0 = (pw == nprod - 1);
    const unsigned my_bytes = (do_pairs ? 2u * row_bytes : 0u) + (do_d0 ? row_bytes : 0u);
    int p = 0, q = 0;
    unsigned pphase = 0, qphase = 0;  // parity of the release a re-used ring slot waits for: ((k / ring) - 1) & 1
    for (int64_t kb = 0; kb < nrows; kb += 16) {
      const int nb = (nrows - kb < 16) ? (int)(nrows - kb) : 16;
      int64_t r[2] = {0, 0};
      int nstart = 0;
      const int64_t c_mine = c0 + (kb + (lane & 15)) * cstep;  // (rows past the end: evaluated, never issued)
      if (!P.replay) {
        const int64_t cr = c0 + (kb + (lane >> 1)) * cstep;
        const u32x4 v = philox4x32_10((uint32_t)(lane & 1), (uint32_t)(P.global_offset + cr), P.generation, STREAM_DONORS, P.key0, P.key1);
        const uint64_t dc_even = ((uint64_t)v.x << 32) | v.y, dc_odd = ((uint64_t)v.z << 32) | v.w;
        uint64_t w[3];
        w[0] = __shfl_sync(0xffffffffu, dc_even, (2 * lane) & 31);
        w[1] = __shfl_sync(0xffffffffu, dc_odd, (2 * lane) & 31);
        w[2] = __shfl_sync(0xffffffffu, dc_even, (2 * lane + 1) & 31);
        resolve_donors<2>(w, c_mine, P.np, D, r, &nstart);
      } else if (lane < nb) {
        r[0] = __ldg(P.r_donors + c_mine * 5 + 0);
        r[1] = __ldg(P.r_donors + c_mine * 5 + 1);
        nstart = __ldg(P.r_nstart + c_mine);
      }
      for (int j = 0; j < nb; ++j) {
        const int64_t k = kb + j;
        // row j's draws to every lane (warp-uniform from here: one lane issues, the others fall through to the
        // __syncwarp -- a lane-divergent spin on the mbarriers must not be left to the scheduler)
        const int64_t rj = __shfl_sync(0xffffffffu, do_d0 ? r[0] : r[1], j);
        const int nsj = __shfl_sync(0xffffffffu, nstart, j);
        const int64_t rj1 = (nprod == 1) ? __shfl_sync(0xffffffffu, r[1], j) : rj;
        if (lane == 0) {
          const int64_t c = c0 + k * cstep;
          if (do_pairs && k >= PR) mbar_wait(s_pair_free + p, pphase);
          if (k >= QR) mbar_wait(s_d0_free + q, qphase);
          if (do_d0) s_nstart[q] = make_int2(nsj, (int)k);
          mbar_expect_tx(s_full + q, my_bytes);  // (the arrive publishes s_nstart)
          if (do_pairs) {
            unsigned char* pb = s_pairs + (unsigned)(2 * p) * row_stride;
            bulk_g2s(pb, P.pop_cur + c * D, row_bytes, s_full + q);
            bulk_g2s(pb + row_stride, P.pop_cur + rj1 * D, row_bytes, s_full + q);
          }
          if (do_d0) bulk_g2s(s_d0s + (unsigned)q * row_stride, P.pop_cur + rj * D, row_bytes, s_full + q);
        }
        __syncwarp();
        if (++p == PR) {
          p = 0;
          if (k >= PR) pphase ^= 1u;
        }
        if (++q == QR) {
          q = 0;
          if (k >= QR) qphase ^= 1u;
        }
      }
    }
  } else {
    // ---- consumer groups
    const int group = threadIdx.x / gthreads;
    const int tig = threadIdx.x - group * gthreads;
    const bool leader = (tig == 0);
    const Group<GW> grp = {tig, tig >> 5, lane, 1 + group, s_scratch + group * 16};
    double g_best_e = CUDART_INF;
    int64_t g_best_idx = INT64_MAX;
    unsigned long long g_ninf = 0, g_nacc = 0;
    const int npairs = D >> 1;  // double2 chunks per row
    const bool uni_b = P.bounds_uniform != 0;
    const int tq = tig & 3;
    const bool tq0 = (tq & 1) != 0, tq1 = (tq & 2) != 0;

    // ring slots of the group's rows k = group, group + G, ...: k mod P, k mod Q and the parity of (k / Q), kept incrementally
    int p = group % PR, q = group % QR;
    unsigned fparity = (unsigned)((group / QR) & 1);
    const int dp = ngroups % PR, dq = ngroups % QR;
    const unsigned dpar = (unsigned)((ngroups / QR) & 1);
    for (int64_t k = group; k < nrows; k += ngroups) {
      const int64_t c = c0 + k * cstep;
      const double2* s_par = reinterpret_cast<const double2*>(s_pairs + (unsigned)(2 * p) * row_stride);
      const double2* s_d1 = reinterpret_cast<const double2*>(s_pairs + (unsigned)(2 * p + 1) * row_stride);
      double2* s_d0 = reinterpret_cast<double2*>(s_d0s + (unsigned)q * row_stride);  // becomes the trial vector
      const double2* s_b2 = reinterpret_cast<const double2*>(s_best);
      const uint32_t cand_g = (uint32_t)(P.global_offset + c);
      const double e_old = __ldg(P.e_cur + c);  // in flight while the group waits for its rows

      // ---- the row's crossover decisions, BEFORE the wait for its buffers: they depend on the row number only (Philox
      // counters / the recorded uniforms), so the ten Philox rounds and the quad transposes of every trip run while the
      // copies fly instead of after them.  8 bits per trip (4 chunks of 2 dimensions per thread), up to 8 trips.
      auto draw_bits = [&](int base, auto guard) -> uint32_t {
        constexpr bool GUARD = decltype(guard)::value;
        uint32_t bits = 0;
        if (P.replay) {
          const double* u = P.r_uni + c * (int64_t)(D + 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int pp = base + j * gthreads + tig;
            if (!GUARD || pp < npairs) {
              bits |= (__ldg(u + 2 * pp) < P.CR ? 1u : 0u) << (2 * j);
              bits |= (__ldg(u + 2 * pp + 1) < P.CR ? 1u : 0u) << (2 * j + 1);
            }
          }
        } else {
          const int qq = ((base + tq * gthreads + tig) >> 2);
          const u32x4 w = philox4x32_10((uint32_t)qq, cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
          const uint32_t w01 = tq0 ? w.y : w.x, w23 = tq0 ? w.w : w.z;
          const uint32_t x01 = tq0 ? w.x : w.y, x23 = tq0 ? w.z : w.w;
          uint32_t recv[4];
          recv[0] = tq1 ? w23 : w01;
          recv[1] = __shfl_xor_sync(0xffffffffu, tq1 ? x23 : x01, 1);
          recv[2] = __shfl_xor_sync(0xffffffffu, tq1 ? w01 : w23, 2);
          recv[3] = __shfl_xor_sync(0xffffffffu, tq1 ? x01 : x23, 3);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool b0 = ((j & 1) != 0) != tq0, b1 = ((j & 2) != 0) != tq1;
            const uint32_t lo = b0 ? recv[1] : recv[0], hi = b0 ? recv[3] : recv[2];
            const uint32_t word = b1 ? hi : lo;
            bits |= (field_lo_hit(word, P.thr16) ? 1u : 0u) << (2 * j);
            bits |= (field_hi_hit(word, P.thr16) ? 1u : 0u) << (2 * j + 1);
          }
        }
        return bits;
      };
      uint64_t xmask = 0;
      auto draw_row = [&]() {
        if (npairs % (4 * gthreads) == 0) {
          for (int base = 0, t = 0; base < npairs; base += 4 * gthreads, ++t) xmask |= (uint64_t)draw_bits(base, std::false_type()) << (8 * t);
        } else {
          for (int base = 0, t = 0; base < npairs; base += 4 * gthreads, ++t) xmask |= (uint64_t)draw_bits(base, std::true_type()) << (8 * t);
        }
      };
      if (MB2_POOL_EARLY_DRAWS) draw_row();

      // A parity test tells phase n from phase n - 1 only: were the slot still loading its PREVIOUS row (k - Q, issued
      // before any row this group has finished, but not necessarily landed), the test would pass one phase early.  The
      // producer therefore publishes the row number with the start index, and a group that finds another row's number
      // waits again (never seen in practice: a row's copies would have to outlast the whole processing of a later row).
      if (GW == 1 || grp.wig == 0) {  // (the group's first warp polls, the others block on the named barrier: group_wait_stage)
        int seen;
        do {
          mbar_wait(s_full + q, fparity);
          seen = reinterpret_cast<volatile int2*>(s_nstart + q)->y;
        } while (seen != (int)k);
      }
      if (GW > 1) group_sync(grp);
      const int nstart = s_nstart[q].x;
      if (!MB2_POOL_EARLY_DRAWS) draw_row();  // (A/B build: the draws after the wait, as K2a makes them)

      // ---- trial vector (strategy.py:77-85): K2a's loop (de_tma.cuh), rows of even length
      bool oob = false;
      auto trip = [&](int base, uint32_t bits, auto guard, auto uclamp) {
        constexpr bool GUARD = decltype(guard)::value;
        constexpr bool UCLAMP = decltype(uclamp)::value;
        bool cx[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pp = base + j * gthreads + tig;
          cx[2 * j] = (!GUARD || pp < npairs) && ((2 * pp == nstart) || ((bits >> (2 * j)) & 1u) != 0);
          cx[2 * j + 1] = (!GUARD || pp < npairs) && ((2 * pp + 1 == nstart) || ((bits >> (2 * j + 1)) & 1u) != 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int pp = base + j * gthreads + tig;
          if (!GUARD || pp < npairs) {
            double2 x = s_par[pp];
            const double2 a = s_d0[pp], b = s_d1[pp], bs = s_b2[pp];
            if (cx[2 * j]) x.x = dadd(bs.x, dmul(P.F, dsub(a.x, b.x)));
            if (cx[2 * j + 1]) x.y = dadd(bs.y, dmul(P.F, dsub(a.y, b.y)));
            if (UCLAMP) {
              x.x = (x.x > P.lo_s) ? x.x : P.lo_s;
              x.x = (x.x < P.hi_s) ? x.x : P.hi_s;
              x.y = (x.y > P.lo_s) ? x.y : P.lo_s;
              x.y = (x.y < P.hi_s) ? x.y : P.hi_s;
            } else if (P.bounds_mode != BOUNDS_NONE) {
              const double l0 = uni_b ? P.lo_s : __ldg(P.lo + 2 * pp), l1 = uni_b ? P.lo_s : __ldg(P.lo + 2 * pp + 1);
              const double h0 = uni_b ? P.hi_s : __ldg(P.hi + 2 * pp), h1 = uni_b ? P.hi_s : __ldg(P.hi + 2 * pp + 1);
              if (P.bounds_mode == BOUNDS_CLAMP) {
                x.x = (x.x > l0) ? x.x : l0;
                x.x = (x.x < h0) ? x.x : h0;
                x.y = (x.y > l1) ? x.y : l1;
                x.y = (x.y < h1) ? x.y : h1;
              } else {
                oob |= (x.x < l0) || (x.x > h0) || (x.y < l1) || (x.y > h1);
              }
            }
            s_d0[pp] = x;
          }
        }
      };
      if (npairs % (4 * gthreads) == 0) {
        if (uni_b && P.bounds_mode == BOUNDS_CLAMP) {
          for (int base = 0, t = 0; base < npairs; base += 4 * gthreads, ++t)
            trip(base, (uint32_t)(xmask >> (8 * t)) & 0xffu, std::false_type(), std::true_type());
        } else {
          for (int base = 0, t = 0; base < npairs; base += 4 * gthreads, ++t)
            trip(base, (uint32_t)(xmask >> (8 * t)) & 0xffu, std::false_type(), std::false_type());
        }
      } else {
        for (int base = 0, t = 0; base < npairs; base += 4 * gthreads, ++t)
          trip(base, (uint32_t)(xmask >> (8 * t)) & 0xffu, std::true_type(), std::false_type());
      }
      oob = group_any(grp, oob);  // also the barrier that publishes the trial vector to the group
      // every thread of the group has read the parent and the second donor: their buffers go back to the producer
      if (leader) mbar_arrive(s_pair_free + p);

      const double* s_trial = reinterpret_cast<const double*>(s_d0);
      double E = CUDART_INF;
      if (!oob) E = group_cost<COST, GW>(grp, s_trial, D, P.cost);
      E = dadd(E, group_penalty(grp, s_trial, D, P.pen));

      if (P.cap_trial != nullptr) {
        for (int i = tig; i < D; i += gthreads) P.cap_trial[c * D + i] = s_trial[i];
        if (leader) P.cap_energy[c] = E;
      }

      const bool acc = E < e_old;
      // generic-proxy writes of the trial -> visible to the async proxy (the store below, the producer's next copy
      // into this buffer); all reads of the buffer done
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
      if (leader) {
        if (acc) {
          bulk_s2g(P.pop_next + c * D, s_d0, row_bytes);
          bulk_commit();
        }
        P.e_next[c] = acc ? E : e_old;
        if (acc) bulk_wait_read0();  // the store has finished reading the buffer
        mbar_arrive(s_d0_free + q);
      }
      p += dp;
      if (p >= PR) p -= PR;
      q += dq;
      fparity ^= dpar;
      if (q >= QR) {
        q -= QR;
        fparity ^= 1u;
      }
    }
    if (leader) {
      bulk_wait_all0();
      sh_e[group] = g_best_e;
      sh_i[group] = g_best_idx;
      sh_ninf[group] = g_ninf;
      sh_nacc[group] = g_nacc;
    }
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
