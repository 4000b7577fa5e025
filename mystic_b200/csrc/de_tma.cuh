// Full-row (binomial crossover) generation kernel with TMA bulk staging.
//
// Best1Bin reads whole rows: the parent and two donors, 8*D bytes each, contiguous in HBM.  A
// GROUP of G warps (G in {1,2,4,8}) owns one candidate row at a time and a private ring of
// `stages` buffers in shared memory; the group's leader thread issues three `cp.async.bulk`
// global->shared copies per candidate (SASS: UBLKCP) that complete on the group's mbarrier, so
// (groups x stages x 3 rows) are in flight per SM without holding registers.  The trial vector
// is built in place over the first donor's buffer, the cost is evaluated from shared memory with
// the canonical reduction of de_device.cuh, and the selected row (trial or parent) leaves with
// one bulk shared->global copy.  Groups synchronise on their own named barrier; there is no
// CTA-wide barrier after start-up.
//
// Semantics are those of de_step_kernel<BASE_BEST1, XOVER_BIN, COST> (de_kernels.cuh); the same
// device functions produce the numbers, so both kernels agree bit for bit.
#pragma once
#include <type_traits>

#include "de_kernels.cuh"

namespace mb2 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// add to the transaction count of the current phase WITHOUT arriving: lets a copy that does not depend on
// the draws (the parent row) start before the draws are made; the arrive follows with the rest of the bytes
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#ifndef MB2_TMA_WAIT_HINT_NS
#define MB2_TMA_WAIT_HINT_NS 1000000u
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      // the hint lets the hardware keep the thread suspended until the phase completes instead of
      // returning to this loop every few hundred cycles: fewer issue slots and less power spent polling
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@!p bra WAIT_LOOP;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(MB2_TMA_WAIT_HINT_NS)
      : "memory");
}
// A group's wait for its stage: with MB2_TMA_LEADER_WAIT the group's first warp alone polls the
// mbarrier and the other warps block on the group's named barrier (bar.sync does not issue while it
// waits; a try_wait loop does).  The barrier orders the leader's acquire before the others' reads.
#ifndef MB2_TMA_LEADER_WAIT
#define MB2_TMA_LEADER_WAIT 1
#endif
template <int GW>
__device__ __forceinline__ void group_wait_stage(const Group<GW>& g, uint64_t* bar, unsigned parity) {
  if (GW == 1 || !MB2_TMA_LEADER_WAIT) {
    mbar_wait(bar, parity);
  } else {
    if (g.wig == 0) mbar_wait(bar, parity);
    group_sync(g);
  }
}
// global -> shared bulk copy, completion counted in bytes on `bar` (size multiple of 16, 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kTmaMaxGroups = 32;  // groups per CTA (named barriers 1..15 serve groups of more than one warp)
constexpr int kTmaMaxStages = 4;
constexpr int kTmaMaxThreads = 1024;

// donors + start index of candidate c (replayed or Philox); collective over the calling warp
// (the two Philox blocks are computed by two lanes in parallel)
__device__ __forceinline__ void best1_draws(const StepParams& P, int64_t c, int lane, int64_t (&r)[2], int* nstart) {
  if (P.replay) {
    r[0] = __ldg(P.r_donors + c * 5 + 0);
    r[1] = __ldg(P.r_donors + c * 5 + 1);
    *nstart = __ldg(P.r_nstart + c);
  } else {
    draw_donors_warp<2>(P.key0, P.key1, P.generation, (uint32_t)(P.global_offset + c), c, P.np, P.dim, lane, r, nstart);
  }
}

// requires: MODE_STEP, 16-byte aligned population buffers.  blockDim = groups * GW * 32 <= MAXT.
// Rows of ODD length (8*D is not a multiple of 16, every other row starts 8 bytes off a 16-byte boundary): a buffer
// holds the 16-byte aligned COVER of its row, D + 1 elements starting `ph` = (row index & 1) elements before the row
// (the last element of the previous row; past the row's end for even rows: the first element of the next one), so
// element i of the row sits at slot ph + i.  The three rows of a stage have their own phases: the trial loop then
// uses 8-byte shared-memory accesses, and the selected row leaves by plain stores of the whole group.  The one
// cover that would reach past the end of the buffer (the last row when the row count is odd) is copied two
// elements short and completed by a plain load of the leader.
template <int COST, int GW, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
    de_step_best1bin_tma_kernel(const __grid_constant__ StepParams P, int stages) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kTmaMaxGroups * kTmaMaxStages];
  __shared__ double s_scratch[kTmaMaxGroups * 16];
  __shared__ int s_nstart[kTmaMaxGroups * kTmaMaxStages];  // crossover start index of the row in each stage
  __shared__ double sh_e[kTmaMaxGroups];
  __shared__ int64_t sh_i[kTmaMaxGroups];
  __shared__ unsigned long long sh_ninf[kTmaMaxGroups], sh_nacc[kTmaMaxGroups];

  const int D = P.dim;
  const bool odd = (D & 1) != 0;
  const unsigned row_bytes = (unsigned)D * 8u;
  const unsigned row_stride = (row_bytes + (odd ? 8u : 0u) + 127u) & ~127u;  // keep every buffer 128-B aligned (32-bit offsets: cheap to rematerialise)
  constexpr int gthreads = GW * 32;
  const int group = threadIdx.x / gthreads;
  const int ngroups = blockDim.x / gthreads;
  const int tig = threadIdx.x - group * gthreads;
  const int lane = threadIdx.x & 31;
  const bool leader = (tig == 0);
  const Group<GW> grp = {tig, tig >> 5, lane, 1 + group, s_scratch + group * 16};

  double* s_best = reinterpret_cast<double*>(smem_raw);
  unsigned char* s_ring = smem_raw + (row_stride + (unsigned)(group * stages) * 3u * row_stride);
  uint64_t* bars = s_bar + group * kTmaMaxStages;

  for (int i = threadIdx.x; i < D; i += blockDim.x) s_best[i] = __ldg(P.best_x + i);
  if (leader) {
    for (int s = 0; s < stages; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t g0 = (int64_t)blockIdx.x * ngroups + group;
  const int64_t ng = (int64_t)gridDim.x * ngroups;

  // issue the three row loads of candidate c into stage s (group leader only).  The draws are made
  // once, here; the start index travels to the group through shared memory (published by the
  // release of the mbarrier arrive, acquired by the group's wait on the same barrier).
  // called by the whole leader WARP of the group (tig < 32): collective draws, lane 0 issues
  // The group's rows are issued in the order g0, g0 + ng, g0 + 2 ng, ...: ONE Philox evaluation of the leader
  // warp makes the donor / start-index words of the next 16 of them (lane l: block l & 1 of row (l >> 1) ahead),
  // kept in two registers per lane and handed out by shuffle -- instead of one evaluation per row in which 30 of
  // the 32 lanes idle.  Same counters, same words as draw_donors<2>.
  uint64_t dc_even = 0, dc_odd = 0;
  int dc_pos = 16;  // next cached row (16: the cache is empty)
  auto issue = [&](int64_t c, int s) {
    int64_t r[2];
    int nstart;
    if (P.replay) {
      best1_draws(P, c, lane, r, &nstart);
    } else {
      if (dc_pos == 16) {
        const int64_t cr = c + (int64_t)(lane >> 1) * ng;  // rows past the end: evaluated, never handed out
        const u32x4 v = philox4x32_10((uint32_t)(lane & 1), (uint32_t)(P.global_offset + cr), P.generation, STREAM_DONORS, P.key0, P.key1);
        dc_even = ((uint64_t)v.x << 32) | v.y;
        dc_odd = ((uint64_t)v.z << 32) | v.w;
        dc_pos = 0;
      }
      uint64_t w[3];
      w[0] = __shfl_sync(0xffffffffu, dc_even, 2 * dc_pos);
      w[1] = __shfl_sync(0xffffffffu, dc_odd, 2 * dc_pos);
      w[2] = __shfl_sync(0xffffffffu, dc_even, 2 * dc_pos + 1);
      resolve_donors<2>(w, c, P.np, D, r, &nstart);
      ++dc_pos;
    }
    if (leader) {
      unsigned char* buf = s_ring + (unsigned)s * 3u * row_stride;
      if (!odd) {
        s_nstart[group * kTmaMaxStages + s] = nstart;
        mbar_expect_tx(bars + s, 3u * row_bytes);
        bulk_g2s(buf, P.pop_cur + c * D, row_bytes, bars + s);
        bulk_g2s(buf + row_stride, P.pop_cur + r[0] * D, row_bytes, bars + s);
        bulk_g2s(buf + 2 * row_stride, P.pop_cur + r[1] * D, row_bytes, bars + s);
      } else {
        const int64_t rr[3] = {c, r[0], r[1]};
        unsigned bytes[3], total = 0;
        int phases = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int ph = (int)(rr[k] & 1);  // (row * D) & 1 with D odd
          phases |= ph << k;
          const bool short_cover = (ph == 0) && (rr[k] == P.np - 1);  // the cover would end one element past the buffer
          bytes[k] = short_cover ? row_bytes - 8u : row_bytes + 8u;
          total += bytes[k];
          if (short_cover)  // its last element by a plain load (at most one row of a population)
            reinterpret_cast<double*>(buf + (unsigned)k * row_stride)[D - 1] = __ldg(P.pop_cur + rr[k] * D + (D - 1));
        }
        s_nstart[group * kTmaMaxStages + s] = nstart | (phases << 16);
        mbar_expect_tx(bars + s, total);  // (the arrive also publishes the plain store above and s_nstart)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          bulk_g2s(buf + (unsigned)k * row_stride, P.pop_cur + rr[k] * D - (rr[k] & 1), bytes[k], bars + s);
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
  const int npairs = (D + 1) >> 1;  // double2 chunks per row (the last one of an odd row holds one element)
  const bool uni_b = P.bounds_uniform != 0;

  int s = 0;
  unsigned parity = 0;
  for (int64_t c = g0; c < P.np; c += ng) {
    unsigned char* buf = s_ring + (unsigned)s * 3u * row_stride;
    const double2* s_par = reinterpret_cast<const double2*>(buf);
    double2* s_d0 = reinterpret_cast<double2*>(buf + row_stride);  // becomes the trial vector
    const double2* s_d1 = reinterpret_cast<const double2*>(buf + 2 * row_stride);
    const double2* s_b2 = reinterpret_cast<const double2*>(s_best);

    const uint32_t cand_g = (uint32_t)(P.global_offset + c);
    const double e_old = __ldg(P.e_cur + c);  // in flight while the group waits for its rows

    group_wait_stage(grp, bars + s, parity);
    const int nstart_ph = s_nstart[group * kTmaMaxStages + s];
    const int nstart = nstart_ph & 0xffff;
    // odd rows: element i of the parent / first donor (trial) / second donor sits at slot ph + i of its buffer
    const double* o_par = reinterpret_cast<const double*>(buf) + ((nstart_ph >> 16) & 1);
    double* o_d0 = reinterpret_cast<double*>(buf + row_stride) + ((nstart_ph >> 17) & 1);
    const double* o_d1 = reinterpret_cast<const double*>(buf + 2 * row_stride) + ((nstart_ph >> 18) & 1);

    // ---- trial vector (strategy.py:77-85), four conflict-free double2 chunks per thread per trip: chunk
    //      p_j = base + j*gthreads + tig, j = 0..3.  Crossover decisions (philox.cuh, protocol 2): Philox block q
    //      holds eight 16-bit fields = dims 8q..8q+7 = chunks 4q..4q+3, i.e. word t of block q decides chunk 4q+t.
    //      The four threads of a quad (4m..4m+3) own, for every j, the four chunks of ONE block q_j; thread t
    //      evaluates block q_t and the quad transposes the 4x4 words by three xor-shuffles, so one Philox
    //      evaluation per thread covers its eight dimensions and no evaluation is duplicated.
    bool oob = false;
    const int tq = tig & 3;
    const bool tq0 = (tq & 1) != 0, tq1 = (tq & 2) != 0;
    auto trip = [&](int base, auto guard, auto uclamp, auto oddrow) {
      constexpr bool GUARD = decltype(guard)::value;
      constexpr bool UCLAMP = decltype(uclamp)::value;  // clamp to one [lo_s, hi_s] for every dimension, known outside the loop
      constexpr bool ODD = decltype(oddrow)::value;     // 8-byte accesses at the buffers' phases; the last chunk holds one element
      bool cx[8];  // crossover decisions of dims 2p_j, 2p_j+1, j = 0..3
      if (P.replay) {
        const double* u = P.r_uni + c * (int64_t)(D + 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = base + j * gthreads + tig;
          cx[2 * j] = (!GUARD || p < npairs) && ((2 * p == nstart) || (__ldg(u + 2 * p) < P.CR));
          cx[2 * j + 1] = (!GUARD || p < npairs) && ((2 * p + 1 == nstart) || (__ldg(u + 2 * p + 1) < P.CR));
        }
      } else {
        const int q = ((base + tq * gthreads + tig) >> 2);
        const u32x4 w = philox4x32_10((uint32_t)q, cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
        // 4x4 transpose over the quad: in round k thread t hands partner t^k the word of the partner's chunk
        // (word t^k of its block) and receives word t of the partner's block
        const uint32_t w01 = tq0 ? w.y : w.x, w23 = tq0 ? w.w : w.z;   // words (t&1), 2+(t&1)
        const uint32_t x01 = tq0 ? w.x : w.y, x23 = tq0 ? w.z : w.w;   // words (t&1)^1, 2+((t&1)^1)
        uint32_t recv[4];
        recv[0] = tq1 ? w23 : w01;                                                // word t
        recv[1] = __shfl_xor_sync(0xffffffffu, tq1 ? x23 : x01, 1);              // send word t^1
        recv[2] = __shfl_xor_sync(0xffffffffu, tq1 ? w01 : w23, 2);              // send word t^2
        recv[3] = __shfl_xor_sync(0xffffffffu, tq1 ? x01 : x23, 3);              // send word t^3
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // block q_j was evaluated by quad member j = partner t^(j^t): the word is recv[j ^ t]
          const bool b0 = ((j & 1) != 0) != tq0, b1 = ((j & 2) != 0) != tq1;
          const uint32_t lo = b0 ? recv[1] : recv[0], hi = b0 ? recv[3] : recv[2];
          const uint32_t word = b1 ? hi : lo;
          const int p = base + j * gthreads + tig;
          cx[2 * j] = (!GUARD || p < npairs) && ((2 * p == nstart) || field_lo_hit(word, P.thr16));
          cx[2 * j + 1] = (!GUARD || p < npairs) && ((2 * p + 1 == nstart) || field_hi_hit(word, P.thr16));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = base + j * gthreads + tig;
        if (!GUARD || p < npairs) {
          double2 x, a, b, bs;
          const bool second = !ODD || 2 * p + 1 < D;
          if (ODD) {
            x.x = o_par[2 * p], a.x = o_d0[2 * p], b.x = o_d1[2 * p], bs.x = s_best[2 * p];
            x.y = a.y = b.y = bs.y = 0.0;
            if (second) x.y = o_par[2 * p + 1], a.y = o_d0[2 * p + 1], b.y = o_d1[2 * p + 1], bs.y = s_best[2 * p + 1];
          } else {
            x = s_par[p];
            a = s_d0[p], b = s_d1[p], bs = s_b2[p];
          }
          if (cx[2 * j]) x.x = dadd(bs.x, dmul(P.F, dsub(a.x, b.x)));
          if (cx[2 * j + 1]) x.y = dadd(bs.y, dmul(P.F, dsub(a.y, b.y)));
          if (UCLAMP) {
            x.x = (x.x > P.lo_s) ? x.x : P.lo_s;
            x.x = (x.x < P.hi_s) ? x.x : P.hi_s;
            x.y = (x.y > P.lo_s) ? x.y : P.lo_s;
            x.y = (x.y < P.hi_s) ? x.y : P.hi_s;
          } else if (P.bounds_mode != BOUNDS_NONE) {
            const double l0 = uni_b ? P.lo_s : __ldg(P.lo + 2 * p), l1 = (uni_b || !second) ? P.lo_s : __ldg(P.lo + 2 * p + 1);
            const double h0 = uni_b ? P.hi_s : __ldg(P.hi + 2 * p), h1 = (uni_b || !second) ? P.hi_s : __ldg(P.hi + 2 * p + 1);
            if (P.bounds_mode == BOUNDS_CLAMP) {
              x.x = (x.x > l0) ? x.x : l0;
              x.x = (x.x < h0) ? x.x : h0;
              x.y = (x.y > l1) ? x.y : l1;
              x.y = (x.y < h1) ? x.y : h1;
            } else {
              oob |= (x.x < l0) || (x.x > h0) || (second && ((x.y < l1) || (x.y > h1)));  // nothing is out of range after a clamp
            }
          }
          if (ODD) {
            o_d0[2 * p] = x.x;
            if (second) o_d0[2 * p + 1] = x.y;
          } else {
            s_d0[p] = x;
          }
        }
      }
    };
    if (odd) {
      for (int base = 0; base < npairs; base += 4 * gthreads) trip(base, std::true_type(), std::false_type(), std::true_type());
    } else if (npairs % (4 * gthreads) == 0) {  // whole trips only: no per-chunk guards
      if (uni_b && P.bounds_mode == BOUNDS_CLAMP) {
        for (int base = 0; base < npairs; base += 4 * gthreads) trip(base, std::false_type(), std::true_type(), std::false_type());
      } else {
        for (int base = 0; base < npairs; base += 4 * gthreads) trip(base, std::false_type(), std::false_type(), std::false_type());
      }
    } else {
      for (int base = 0; base < npairs; base += 4 * gthreads) trip(base, std::true_type(), std::false_type(), std::false_type());
    }
    oob = group_any(grp, oob);  // also the barrier that publishes the trial vector to the group

    const double* s_trial = o_d0;  // (even rows: phase 0, the first donor's buffer)
    double E = CUDART_INF;
    if (!oob) E = group_cost<COST, GW>(grp, s_trial, D, P.cost);
    E = dadd(E, group_penalty(grp, s_trial, D, P.pen));

    if (P.cap_trial != nullptr) {
      for (int i = tig; i < D; i += gthreads) P.cap_trial[c * D + i] = s_trial[i];
      if (leader) P.cap_energy[c] = E;
    }

    const bool acc = E < e_old;
    // generic-proxy writes of the trial -> visible to the async proxy; all reads of this stage done
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
    if (odd && (acc || !P.sparse)) {
      // rows of odd length leave by plain stores of the whole group (8-byte aligned on both sides)
      const double* src = acc ? s_trial : o_par;
      double* __restrict__ nrow = P.pop_next + c * D;
      for (int i = tig; i < D; i += gthreads) nrow[i] = src[i];
      group_sync(grp);  // the stage has been read before it is refilled
    }
    if (tig < 32) {
      if (leader) {
        const bool store = !odd && (acc || !P.sparse);  // sparse: a rejected row stays where it is (pop_cur is committed in place)
        if (store) {
          bulk_s2g(P.pop_next + c * D, acc ? (const void*)s_d0 : (const void*)s_par, row_bytes);
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
