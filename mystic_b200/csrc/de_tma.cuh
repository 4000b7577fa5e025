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

// requires: dim even (rows are 16-byte multiples), MODE_STEP.  blockDim = groups * GW * 32 <= MAXT.
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
  const unsigned row_bytes = (unsigned)D * 8u;
  const unsigned row_stride = (row_bytes + 127u) & ~127u;  // keep every buffer 128-B aligned (32-bit offsets: cheap to rematerialise)
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
  auto issue = [&](int64_t c, int s) {
    int64_t r[2];
    int nstart;
    best1_draws(P, c, lane, r, &nstart);
    if (leader) {
      s_nstart[group * kTmaMaxStages + s] = nstart;
      unsigned char* buf = s_ring + (unsigned)s * 3u * row_stride;
      mbar_expect_tx(bars + s, 3u * row_bytes);
      bulk_g2s(buf, P.pop_cur + c * D, row_bytes, bars + s);
      bulk_g2s(buf + row_stride, P.pop_cur + r[0] * D, row_bytes, bars + s);
      bulk_g2s(buf + 2 * row_stride, P.pop_cur + r[1] * D, row_bytes, bars + s);
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
  const int npairs = D >> 1;  // double2 chunks per row
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
    const int nstart = s_nstart[group * kTmaMaxStages + s];

    // ---- trial vector (strategy.py:77-85), two conflict-free double2 chunks per thread per trip:
    //      chunk pA = base + tig and pB = pA + gthreads.  Crossover words: Philox block q covers
    //      dims 4q..4q+3, i.e. chunks 2q and 2q+1, so the thread pair (2m, 2m+1) shares blocks qA and
    //      qB; the even thread computes qA, the odd thread qB, and they swap halves by shuffle.
    bool oob = false;
    auto trip = [&](int base, auto guard) {
      constexpr bool GUARD = decltype(guard)::value;
        const int pA = base + tig, pB = base + gthreads + tig;
        bool cx[4];  // crossover decisions for dims 2pA, 2pA+1, 2pB, 2pB+1
        if (P.replay) {
          const double* u = P.r_uni + c * (int64_t)(D + 1);
          cx[0] = (!GUARD || pA < npairs) && ((2 * pA == nstart) || (__ldg(u + 2 * pA) < P.CR));
          cx[1] = (!GUARD || pA < npairs) && ((2 * pA + 1 == nstart) || (__ldg(u + 2 * pA + 1) < P.CR));
          cx[2] = (!GUARD || pB < npairs) && ((2 * pB == nstart) || (__ldg(u + 2 * pB) < P.CR));
          cx[3] = (!GUARD || pB < npairs) && ((2 * pB + 1 == nstart) || (__ldg(u + 2 * pB + 1) < P.CR));
        } else {
          const int q = ((tig & 1) ? pB : pA) >> 1;  // even thread: qA, odd thread: qB
          const u32x4 w = philox4x32_10((uint32_t)q, cand_g, P.generation, STREAM_XOVER, P.key0, P.key1);
          // even thread keeps (x,y) of qA for pA and needs (x,y) of qB from its odd neighbour;
          // odd thread keeps (z,w) of qB for pB and needs (z,w) of qA from its even neighbour.
          const uint32_t send0 = (tig & 1) ? w.x : w.z;
          const uint32_t send1 = (tig & 1) ? w.y : w.w;
          const uint32_t got0 = __shfl_xor_sync(0xffffffffu, send0, 1);
          const uint32_t got1 = __shfl_xor_sync(0xffffffffu, send1, 1);
          const uint32_t a0 = (tig & 1) ? got0 : w.x, a1 = (tig & 1) ? got1 : w.y;  // words for chunk pA
          const uint32_t b0 = (tig & 1) ? w.z : got0, b1 = (tig & 1) ? w.w : got1;  // words for chunk pB
          cx[0] = (!GUARD || pA < npairs) && ((2 * pA == nstart) || ((uint64_t)a0 < P.thr));
          cx[1] = (!GUARD || pA < npairs) && ((2 * pA + 1 == nstart) || ((uint64_t)a1 < P.thr));
          cx[2] = (!GUARD || pB < npairs) && ((2 * pB == nstart) || ((uint64_t)b0 < P.thr));
          cx[3] = (!GUARD || pB < npairs) && ((2 * pB + 1 == nstart) || ((uint64_t)b1 < P.thr));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int p = h ? pB : pA;
          if (!GUARD || p < npairs) {
            double2 x = s_par[p];
            const double2 a = s_d0[p], b = s_d1[p], bs = s_b2[p];
            if (cx[2 * h]) x.x = dadd(bs.x, dmul(P.F, dsub(a.x, b.x)));
            if (cx[2 * h + 1]) x.y = dadd(bs.y, dmul(P.F, dsub(a.y, b.y)));
            if (P.bounds_mode != BOUNDS_NONE) {
              const double l0 = uni_b ? P.lo_s : __ldg(P.lo + 2 * p), l1 = uni_b ? P.lo_s : __ldg(P.lo + 2 * p + 1);
              const double h0 = uni_b ? P.hi_s : __ldg(P.hi + 2 * p), h1 = uni_b ? P.hi_s : __ldg(P.hi + 2 * p + 1);
              if (P.bounds_mode == BOUNDS_CLAMP) {
                x.x = (x.x > l0) ? x.x : l0;
                x.x = (x.x < h0) ? x.x : h0;
                x.y = (x.y > l1) ? x.y : l1;
                x.y = (x.y < h1) ? x.y : h1;
              } else {
                oob |= (x.x < l0) || (x.x > h0) || (x.y < l1) || (x.y > h1);  // nothing is out of range after a clamp
              }
            }
            s_d0[p] = x;
          }
        }
    };
    if (npairs % (2 * gthreads) == 0) {  // whole trips only: no per-chunk guards
#pragma unroll 2
      for (int base = 0; base < npairs; base += 2 * gthreads) trip(base, std::false_type());
    } else {
      for (int base = 0; base < npairs; base += 2 * gthreads) trip(base, std::true_type());
    }
    oob = group_any(grp, oob);  // also the barrier that publishes the trial vector to the group

    const double* s_trial = reinterpret_cast<const double*>(s_d0);
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
    if (tig < 32) {
      if (leader) {
        bulk_s2g(P.pop_next + c * D, acc ? (const void*)s_d0 : (const void*)s_par, row_bytes);
        bulk_commit();
        P.e_next[c] = acc ? E : e_old;
        bulk_wait_read0();  // the store has finished reading the stage
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
