// K2f: the short-row generation step (K2d's row lengths: D a multiple of 4, 16 <= D <= 128) with every
// byte moved by the bulk-copy engine, for the exponential-crossover strategies.
//
// `tools/gather_peak.cu` shows the copy engine moves the access pattern of BASELINE config 4 (a 512-byte
// parent row and three ~96-byte donor segments in, a row out) in 0.240-0.245 ms at NP = 1 M with almost no
// instructions (K2d, de_rows.cuh: 0.341-0.350 ms; the same pattern with plain loads: 0.26-0.27 ms).  Here a
// warp owns a ring of `stages` buffers (one by default) of 32/LPR rows:
//   * issue: lane 0 requests the stage's parent rows with ONE copy (they are contiguous in global memory and
//     the tile is unpadded) before anything else, so they fly while the draws are made (rows_draw,
//     de_rows.cuh: the same Philox blocks as every kernel); the window chunks q0 .. q0+nq-1 (mod D/4) of
//     each row get a slot range in the stage's donor pools by a prefix sum over the rows, and one SIMT copy
//     instruction has lane 1 + j of a row's group request the window's cover of donor j (a second one when
//     the window wraps), all onto the stage's mbarrier (`cp.async.bulk`, SASS UBLKCP: a uniform-datapath
//     instruction, so that "one instruction" is an elect loop of ~10 instructions per request);
//   * compute (when the stage has landed): K2d's window mutation in place, with donor chunks read from
//     the pools and the parent's chunks saved in an undo pool; K2d's cost / penalty; a rejected trial is
//     undone in the tile; ONE bulk shared->global copy stores the stage's rows.
// Rows whose window does not fit the pools read their donors with plain loads and are restored from
// global memory when rejected (the pools are sized from what shared memory is left: 36 chunks per stage
// at C4, mean demand 25).  Results are bit-identical to K2d's: same draws, same mutation and cost code.
// B200 (profiles/README.md): the step is bound by the warps' own dependent chains, so 24 resident warps at
// 80 registers (0.30 ms at C4) beat 12 warps with two stages each (0.40 ms).
#pragma once
#include "de_rows.cuh"
#include "de_tma.cuh"

namespace mb2 {

#ifndef MB2_RB_MAXWARPS
#define MB2_RB_MAXWARPS 24
#endif
constexpr int kBulkRowsMaxWarps = MB2_RB_MAXWARPS;  // 24 warps: 80 registers per thread (no spills)
constexpr int kBulkRowsMaxStages = 4;

struct BulkRowMeta {  // one per row of a stage, written at issue time
  int nstart, L, q0, nq, off;  // off: first slot of the row in the donor / undo pools, -1: not staged
  int r[5];
  int pad_[2];
};
static_assert(sizeof(BulkRowMeta) == 48, "meta slot");

// shared-memory layout, computed once on the host (de_abi.cu: rows_bulk_geometry) so that the kernel reads
// it from the constant bank instead of re-deriving it with 64-bit arithmetic inside the loop:
// best member | per warp: undo pool, then per stage: rows, K donor pools, meta
struct BulkRowsLayout {
  int stages;            // stages per warp
  int slots;             // 4-element chunks per pool
  unsigned best_bytes;   // bestSolution [D], rounded to 128
  unsigned rows_bytes;   // the stage's rows
  unsigned pool_bytes;   // one pool (slots * 32)
  unsigned meta_off;     // rows_bytes + K * pool_bytes
  unsigned stage_bytes;  // rounded to 128
  unsigned undo_bytes;   // the warp's undo pool, rounded to 128
  unsigned warp_bytes;   // undo_bytes + stages * stage_bytes
  int donors_ldgsts;     // 1: donor windows by per-lane 16-byte cp.async (LDGSTS), 0: by cp.async.bulk requests
};

// 16-byte asynchronous global -> shared copy through the LSU path (SASS LDGSTS.E.BYPASS.128): no registers held,
// completion by cp.async.wait_all of the issuing thread
__device__ __forceinline__ void ldgsts16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ldgsts_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int COST, int LPR, int K>
__global__ void __launch_bounds__(kBulkRowsMaxWarps * 32, 1)
    de_step_rows_bulk_kernel(const __grid_constant__ StepParams P, int base, const __grid_constant__ BulkRowsLayout G) {
  if (P.stop != nullptr && *P.stop) return;  // batch mode: a termination condition was met earlier
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kBulkRowsMaxWarps * kBulkRowsMaxStages];
  __shared__ double sh_e[kBulkRowsMaxWarps];
  __shared__ int64_t sh_i[kBulkRowsMaxWarps];
  __shared__ unsigned long long sh_ninf[kBulkRowsMaxWarps], sh_nacc[kBulkRowsMaxWarps];
  constexpr int RPW = 32 / LPR;  // rows per warp
  constexpr unsigned LMASK = (1u << LPR) - 1u;
  static_assert(LPR >= 4 && LPR <= 16 && (K + 2) / 2 < LPR, "group too narrow for the draws");

  const int D = P.dim;
  const int NQ = D >> 2;
  // The tile is unpadded (K2d pads its rows by 16 bytes against bank conflicts): the rows of a stage are
  // contiguous in global memory, so ONE copy brings all parents in and ONE takes the next generation out
  // (a copy request costs ~8 instructions of an elect loop; two-way conflicts on the 16-byte tile reads cost less)
  const int TS = D;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpc = blockDim.x >> 5;
  const int grp = lane / LPR;
  const int sub = lane % LPR;
  const int gbase = grp * LPR;
  const bool needs_best = (base == BASE_BEST1 || base == BASE_RANDTOBEST1 || base == BASE_BEST2);
  const unsigned row_bytes = (unsigned)D * 8u;

  const int stages = G.stages, wq_cap = G.slots;
  double* s_best = reinterpret_cast<double*>(smem_raw);
  unsigned char* w_base = smem_raw + G.best_bytes + warp * G.warp_bytes;
  double* s_undo = reinterpret_cast<double*>(w_base);
  unsigned char* s_stages = w_base + G.undo_bytes;
  uint64_t* bars = s_bar + warp * kBulkRowsMaxStages;

  if (needs_best)
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_best[i] = __ldg(P.best_x + i);
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bars + s, RPW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t nchunks = (P.np + RPW - 1) / RPW;
  const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
  const int64_t nw = (int64_t)gridDim.x * wpc;

  // ---- draws + loads of the rows of chunk `ch` into the stage at `st` (whole warp)
  auto issue = [&](int64_t ch, unsigned char* st, uint64_t* bar) {
    const int64_t c = ch * RPW + grp;
    const int64_t cc = (c < P.np) ? c : (P.np - 1);  // idle groups shadow the last row
    const int64_t left = P.np - ch * RPW;
    const unsigned nlive = left < RPW ? (unsigned)left : (unsigned)RPW;
    // the parent rows do not depend on the draws: their copy flies while the draws are made
    if (lane == 0) {
      mbar_expect_tx_only(bar, nlive * row_bytes);
      bulk_g2s(st, P.pop_cur + ch * RPW * D, nlive * row_bytes, bar);
    }
    RowDraws<K> dr;
    rows_draw<LPR, K>(P, XOVER_EXP, cc, sub, gbase, dr);
    const int q0 = dr.nstart >> 2;
    int nq = (dr.L > 0) ? (((dr.nstart + dr.L - 1) >> 2) - q0 + 1) : 0;
    if (nq > NQ) nq = NQ;  // a full-length window that starts inside a chunk: each chunk once
    if (c >= P.np) nq = 0;   // idle groups mutate nothing (their tile rows hold stale data, results are discarded)
    // slots of the row in the pools: exclusive prefix sum of nq over the rows of the warp
    int incl = (sub == 0) ? nq : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int off = __shfl_sync(0xffffffffu, incl, gbase) - nq;
    const bool fits = (nq > 0) && (off + nq <= wq_cap);
    if (!fits) off = -1;
    if (sub == 0) {
      BulkRowMeta* m = reinterpret_cast<BulkRowMeta*>(st + G.meta_off) + grp;
      m->nstart = dr.nstart;
      m->L = dr.L;
      m->q0 = q0;
      m->nq = nq;
      m->off = off;
#pragma unroll
      for (int j = 0; j < K; ++j) m->r[j] = dr.r[j];
      // the arrive of this row: meta is published with it
      mbar_expect_tx(bar, (fits && !G.donors_ldgsts) ? (unsigned)(K * nq) * 32u : 0u);
    }
    __syncwarp();
    if (G.donors_ldgsts) {
      // Donor windows through the LSU path: the bulk-copy engine takes ~20 cycles per request whatever its size
      // (the C4 pattern is ~3 requests of ~100 bytes per row: request-rate bound, ncu shows the warps queueing on
      // UBLKCP), and a request costs an elect loop of ~10 instructions.  Here lane `sub` of the row's group copies
      // the 16-byte halves sub, sub+LPR, ... of the window's chunks, donor by donor; one instruction serves the
      // RPW rows of the warp.  Completion: cp.async.wait_all + __syncwarp before the stage is read.
      if (fits) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const double* srow = P.pop_cur + (int64_t)dr.r[j] * D;
          double* pool = reinterpret_cast<double*>(st + G.rows_bytes + (unsigned)j * G.pool_bytes) + 4 * off;
          for (int t = sub; t < 2 * nq; t += LPR) {
            int q = q0 + (t >> 1);
            if (q >= NQ) q -= NQ;
            ldgsts16(pool + 2 * t, srow + 4 * q + 2 * (t & 1));
          }
        }
      }
      return;
    }
    // one copy instruction serves the whole warp: lane 1 + j of a row's group requests the window's cover
    // of donor j (a second instruction for the wrapped parts; further rounds when K > LPR - 1)
    const int n1 = (q0 + nq <= NQ) ? nq : (NQ - q0);  // chunks before the wrap
#pragma unroll
    for (int jb = 0; jb < K; jb += LPR - 1) {
      const int j = jb + sub - 1;
      const bool donor = fits && sub >= 1 && j < K;
      int rj = 0;
#pragma unroll
      for (int t = 0; t < K; ++t)
        if (t == j) rj = dr.r[t];
      const double* src = P.pop_cur + (int64_t)rj * D + 4 * q0;
      double* dst = reinterpret_cast<double*>(st + G.rows_bytes + (unsigned)(donor ? j : 0) * G.pool_bytes) + 4 * (donor ? off : 0);
      const unsigned bytes = donor ? (unsigned)n1 * 32u : 0u;
      if (bytes) bulk_g2s(dst, src, bytes, bar);
      if (__any_sync(0xffffffffu, donor && n1 < nq)) {
        if (donor && n1 < nq) bulk_g2s(dst + 4 * n1, P.pop_cur + (int64_t)rj * D, (unsigned)(nq - n1) * 32u, bar);
      }
    }
  };

  for (int s = 0; s < stages; ++s) {
    const int64_t ch = gw + (int64_t)s * nw;
    if (ch < nchunks) issue(ch, s_stages + (unsigned)s * G.stage_bytes, bars + s);
  }

  double w_best_e = CUDART_INF;
  int64_t w_best_idx = INT64_MAX;
  unsigned long long w_ninf = 0, w_nacc = 0;

  int s = 0;
  unsigned parity = 0;
  for (int64_t ch = gw; ch < nchunks; ch += nw) {
    unsigned char* st = s_stages + (unsigned)s * G.stage_bytes;
    const int64_t c = ch * RPW + grp;
    const bool live = c < P.np;
    const int64_t cc = live ? c : (P.np - 1);
    double* t_row = reinterpret_cast<double*>(st) + (size_t)grp * TS;
    const double e_old = __ldg(P.e_cur + cc);  // in flight while the warp waits for its stage

    mbar_wait(bars + s, parity);
    if (G.donors_ldgsts) {
      ldgsts_wait_all();  // this lane's donor pieces have landed ...
      __syncwarp();       // ... and so have the other lanes'
    }
    const BulkRowMeta* m = reinterpret_cast<const BulkRowMeta*>(st + G.meta_off) + grp;
    RowDraws<K> dr;
    dr.nstart = m->nstart;
    dr.L = m->L;
    const int q0 = m->q0, nq = m->nq, off = m->off;
    const bool fits = off >= 0;
    if (!fits) {
#pragma unroll
      for (int j = 0; j < K; ++j) dr.r[j] = m->r[j];
    }

    // ---- crossover window: chunks q0 .. q0+nq-1 (mod NQ), lane `sub` takes sub, sub+LPR, ...
    bool oob = false;
    for (int w = sub; w < nq; w += LPR) {
      int q = q0 + w;
      if (q >= NQ) q -= NQ;
      DonorChunk<K> dc;
      if (fits) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const double* dp = reinterpret_cast<const double*>(st + G.rows_bytes + (unsigned)j * G.pool_bytes) + 4 * (off + w);
          dc.v[j][0] = *reinterpret_cast<const double2*>(dp);
          dc.v[j][1] = *reinterpret_cast<const double2*>(dp + 2);
        }
      } else {
        rows_load_donors<K>(P.pop_cur, dr, D, q << 2, dc);
      }
      oob |= rows_mutate_chunk<K>(P, base, XOVER_EXP, dr, cc, q, dc, s_best, t_row, (fits && !P.sparse) ? s_undo + 4 * (off + w) : nullptr);
    }
    oob = ((__ballot_sync(0xffffffffu, oob) >> gbase) & LMASK) != 0;
    __syncwarp();

    // ---- energy = (inf if out of range else cost(x)) + penalty(x); idle groups evaluate their shadow
    // row because the shuffles inside need every lane
    double E = CUDART_INF;
    {
      const double cost = rows_cost<COST, LPR>(t_row, NQ, sub, gbase, P.cost);
      if (!oob) E = cost;
      E = dadd(E, rows_penalty<LPR>(t_row, D, NQ, sub, P.pen));
    }
    if (live && P.cap_trial != nullptr) {
      for (int i = sub; i < D; i += LPR) P.cap_trial[c * D + i] = t_row[i];
      if (sub == 0) P.cap_energy[c] = E;
    }
    __syncwarp();  // every lane has read the trial before a rejected one is undone

    // ---- selection (differential_evolution.py:603-613): a rejected trial is undone in the tile
    const bool acc = E < e_old;
    if (!acc && !P.sparse) {
      for (int w = sub; w < nq; w += LPR) {
        int q = q0 + w;
        if (q >= NQ) q -= NQ;
        double2 a, b;
        if (fits) {
          a = *reinterpret_cast<const double2*>(s_undo + 4 * (off + w));
          b = *reinterpret_cast<const double2*>(s_undo + 4 * (off + w) + 2);
        } else {
          a = ld_stream2(P.pop_cur + cc * D + 4 * q);
          b = ld_stream2(P.pop_cur + cc * D + 4 * q + 2);
        }
        *reinterpret_cast<double2*>(t_row + 4 * q) = a;
        *reinterpret_cast<double2*>(t_row + 4 * q + 2) = b;
      }
    }
    fence_proxy_async();  // generic-proxy writes of the tile -> visible to the bulk store; reads of the stage done
    __syncwarp();
    // sparse: only accepted rows leave, one store each (they are rare when this mode is chosen)
    const bool whole = !P.sparse && (ch + 1) * RPW <= P.np;  // every row of the stage is live: one store for all of them
    if (lane == 0 && whole) {
      bulk_s2g(P.pop_next + ch * RPW * D, st, RPW * row_bytes);
      bulk_commit();
    }
    if (sub == 0 && live) {
      if (!whole && (acc || !P.sparse)) {
        bulk_s2g(P.pop_next + c * D, t_row, row_bytes);
        bulk_commit();
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
      bulk_wait_read0();  // the store has finished reading the stage
    }
    __syncwarp();
    const int64_t chn = ch + (int64_t)stages * nw;
    if (chn < nchunks) issue(chn, st, bars + s);
    if (++s == stages) {
      s = 0;
      parity ^= 1u;
    }
  }
  if (sub == 0) bulk_wait_all0();

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
