// The non-templated kernels of the engine: argmin epilogue (K3) with the device-side termination
// test, population initialisation (K0), multi-GPU best exchange records, self tests.  Included by
// exactly one translation unit (de_abi.cu).
#pragma once
#include "de_kernels.cuh"

namespace mb2 {

// ---- multi-GPU best exchange over peer memory (NVLink), fused into the argmin epilogue -----------
// Every rank owns a mailbox [2 parities][world slots][stride doubles] + flags [2][world]; the
// mailboxes of all ranks are mapped into every rank's address space (CUDA IPC).  After its local
// argmin a rank stores its record [best_e, best_idx(global), n_inf, n_acc, x[D]] into slot `rank`
// of EVERY mailbox (remote stores), fences, raises the slot's flag to the exchange epoch, waits
// until the `world` flags of its own mailbox carry the epoch and merges the records it received:
// lowest energy wins, ties to the lowest global index (differential_evolution.py:611 strict '<'
// in ascending candidate order).  Parities alternate per exchange, so a rank that is one exchange
// ahead never overwrites a record a slower rank is still reading.
constexpr int kMaxWorld = 16;
struct ExchangeParams {
  int world, rank;
  int stride;                     // doubles per slot (>= D + 4)
  int pad_;
  unsigned long long epoch;       // of this exchange, > 0, the same on every rank
  unsigned long long timeout_ns;  // give up waiting for a peer after this long
  double* box[kMaxWorld];         // box[r]: rank r's mailbox as mapped here
  int* error;                     // raised on a timeout
};

__device__ __forceinline__ unsigned long long* exchange_flags(double* box, int world, int stride) {
  return reinterpret_cast<unsigned long long*>(box + (size_t)2 * world * stride);
}

// Returns false when a peer did not deliver within the timeout: *X.error is raised, the mailbox is NOT merged
// (its records are stale) and this shard keeps its local best; the host fails the next read of the result.
__device__ __forceinline__ bool exchange_best(const ExchangeParams& X, int D, double* __restrict__ best_x,
                                              DeviceState* __restrict__ st) {
  __shared__ int sh_w;
  __shared__ int sh_late;
  if (threadIdx.x == 0) sh_late = 0;
  const int par = (int)(X.epoch & 1ull);
  // ---- my record into slot `rank` of every mailbox
  for (int r = 0; r < X.world; ++r) {
    double* rec = X.box[r] + ((size_t)par * X.world + X.rank) * X.stride;
    if (threadIdx.x == 0) {
      rec[0] = st->best_e;
      rec[1] = (double)st->best_idx;
      rec[2] = (double)st->n_inf;
      rec[3] = (double)st->n_acc;
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) rec[4 + i] = best_x[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < X.world) {
    unsigned long long* f = exchange_flags(X.box[threadIdx.x], X.world, X.stride) + (size_t)par * X.world + X.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(X.epoch) : "memory");
    // ---- wait for the record of rank threadIdx.x in my own mailbox
    const unsigned long long* mine = exchange_flags(X.box[X.rank], X.world, X.stride) + (size_t)par * X.world + threadIdx.x;
    unsigned long long t0 = 0, now = 0, v = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v == X.epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > X.timeout_ns) {
        atomicExch(X.error, 1);
        sh_late = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (sh_late || *reinterpret_cast<volatile int*>(X.error) != 0) return false;  // an earlier exchange failed: ranks are out of step
  // ---- merge
  const double* recs = X.box[X.rank] + (size_t)par * X.world * X.stride;
  if (threadIdx.x == 0) {
    int w = -1;
    double be = CUDART_INF, bi = 0.0, ninf = 0.0, nacc = 0.0;
    for (int r = 0; r < X.world; ++r) {
      const volatile double* rec = recs + (size_t)r * X.stride;
      ninf += rec[2];
      nacc += rec[3];
      if (rec[1] < 0.0) continue;  // shard has no best yet
      if (w < 0 || rec[0] < be || (rec[0] == be && rec[1] < bi)) {
        w = r;
        be = rec[0];
        bi = rec[1];
      }
    }
    if (w >= 0) {
      st->best_e = be;
      st->best_idx = (int64_t)bi;
    }
    st->n_inf = (int64_t)ninf;
    st->n_acc = (int64_t)nacc;
    sh_w = w;
  }
  __syncthreads();
  const int w = sh_w;
  if (w >= 0 && w != X.rank) {
    const volatile double* rec = recs + (size_t)w * X.stride;
    for (int i = threadIdx.x; i < D; i += blockDim.x) best_x[i] = rec[4 + i];
  }
  return true;
}

// K3: reduce the per-CTA partials, update bestEnergy/bestSolution if the generation best beats
// the previous best (strict '<', :610), publish counters; with X.world > 1 the shards' bests are
// exchanged over peer memory in the same kernel.  One CTA.
__device__ __forceinline__ void finalize_body(const double* __restrict__ part_e, const int64_t* __restrict__ part_idx,
                                              const unsigned long long* __restrict__ part_ninf,
                                              const unsigned long long* __restrict__ part_nacc, int nparts,
                                              const double* __restrict__ pop_next, int D, int64_t global_offset, int gen0,
                                              int64_t generation, double* __restrict__ best_x, DeviceState* __restrict__ st,
                                              RunState* __restrict__ rs, const ExchangeParams& X, long long* hint = nullptr) {
  if (rs != nullptr && rs->stop) return;  // batch mode: this generation was not run
  __shared__ double sh_e[256];
  __shared__ int64_t sh_i[256];
  __shared__ unsigned long long sh_a[256], sh_b[256];
  __shared__ int64_t sh_winner;
  double be = CUDART_INF;
  int64_t bi = INT64_MAX;
  unsigned long long ni = 0, na = 0;
  for (int k = threadIdx.x; k < nparts; k += blockDim.x) {
    const double e = part_e[k];
    const int64_t i = part_idx[k];
    if (e < be || (e == be && i < bi)) {
      be = e;
      bi = i;
    }
    ni += part_ninf[k];
    na += part_nacc[k];
  }
  sh_e[threadIdx.x] = be;
  sh_i[threadIdx.x] = bi;
  sh_a[threadIdx.x] = ni;
  sh_b[threadIdx.x] = na;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double e = sh_e[threadIdx.x + s];
      const int64_t i = sh_i[threadIdx.x + s];
      if (e < sh_e[threadIdx.x] || (e == sh_e[threadIdx.x] && i < sh_i[threadIdx.x])) {
        sh_e[threadIdx.x] = e;
        sh_i[threadIdx.x] = i;
      }
      sh_a[threadIdx.x] += sh_a[threadIdx.x + s];
      sh_b[threadIdx.x] += sh_b[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double prev = gen0 ? CUDART_INF : st->best_e;
    int64_t winner = -1;
    if (sh_i[0] != INT64_MAX && sh_e[0] < prev) {
      winner = sh_i[0];
      st->best_e = sh_e[0];
      st->best_idx = global_offset + winner;
    } else if (gen0) {
      // nothing finite yet: bestSolution stays population[0], bestEnergy = popEnergy[0] = inf (:559-567)
      st->best_e = CUDART_INF;
      st->best_idx = -1;
      winner = -2;  // copy row 0
    }
    st->n_inf = (int64_t)sh_a[0];
    st->n_acc = (int64_t)sh_b[0];
    st->generation = generation;
    sh_winner = winner;
  }
  __syncthreads();
  const int64_t w = sh_winner;
  if (w != -1) {
    const double* src = pop_next + (w == -2 ? 0 : w) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) best_x[i] = src[i];
  }
  if (X.world > 1) {  // sharded population: install the global best on every shard
    __syncthreads();
    if (!exchange_best(X, D, best_x, st)) {
      // batch mode: the remaining launches of the batch become no-ops; mb2_run_result reports the failure
      if (rs != nullptr && threadIdx.x == 0) rs->stop = 3;
      return;
    }
  }
  // the accepted count of this generation for the host's choice of the commit mode (pinned host memory, no fence:
  // a stale value is as good as a fresh one)
  if (hint != nullptr && threadIdx.x == 0) *reinterpret_cast<volatile long long*>(hint) = st->n_acc;  // (thread 0 wrote st->n_acc)
  if (rs == nullptr) return;
  __syncthreads();
  // ---- batch mode: log the generation, then the host part of Step(): counters and termination
  const long long slot = rs->log_count;
  if (slot < rs->log_cap) {
    double* rec = rs->log + slot * (D + 4);
    for (int i = threadIdx.x; i < D; i += blockDim.x) rec[4 + i] = best_x[i];
    if (threadIdx.x == 0) {
      rec[0] = st->best_e;
      rec[1] = (double)st->best_idx;
      rec[2] = (double)st->n_inf;
      rec[3] = (double)st->n_acc;
    }
  }
  if (threadIdx.x == 0) {
    rs->log_count = slot + 1;
    rs->evaluations += rs->np_eval - st->n_inf;        // differential_evolution.py:601
    if (!gen0) rs->generations += 1;                    // (generation 0 is logged by batches of solvers only)
    const double e = st->best_e;                        // stepmon(bestSolution, bestEnergy) (:617)
    if (rs->hist_count == 0) rs->hist0 = e;
    rs->hist[rs->hist_count % kHistMax] = e;
    rs->hist_count += 1;
    const long long lg = rs->hist_count;
    const int g = rs->lookback;
    // python hist[-g]: g == 0 addresses hist[0]
    const double hg = (g == 0) ? rs->hist0 : rs->hist[(lg - g + (long long)kHistMax * 4) % kHistMax];
    const double h1 = e;
    bool met = false;
    switch (rs->kind) {
      case TERM_VTR:  // termination.py:181-196
        met = fabs(h1 - rs->target) <= rs->tol;
        break;
      case TERM_COG:  // termination.py:198-217
        met = (lg > g) && ((hg - h1) <= rs->tol || hg == h1);
        break;
      case TERM_NCOG:  // termination.py:219-240
        met = (lg > g) && (hg == h1 || 2.0 * (hg - h1) <= rs->tol * (fabs(hg) + fabs(h1)) + 1e-20);
        break;
      case TERM_VTRCOG:  // termination.py:320-342
        met = ((lg > g) && ((hg - h1) <= rs->gtol || hg == h1)) || (fabs(h1 - rs->target) <= rs->tol);
        break;
      case TERM_LIMITS:  // termination.py:386-408
        met = (rs->term_evaluations >= 0 && rs->evaluations >= rs->term_evaluations) ||
              (rs->term_generations >= 0 && rs->generations >= rs->term_generations);
        break;
      default:
        break;
    }
    // the solver's own limits (abstract_solver.py:703-711)
    if (rs->max_evaluations >= 0 && rs->evaluations >= rs->max_evaluations) met = true;
    if (rs->max_generations >= 0 && rs->generations >= rs->max_generations) met = true;
    if (met || rs->log_count >= rs->log_cap) rs->stop = met ? 1 : 2;  // 2: log full (cannot happen: one batch = log_cap launches)
  }
}

__global__ void __launch_bounds__(256)
    de_finalize_kernel(const double* __restrict__ part_e, const int64_t* __restrict__ part_idx,
                       const unsigned long long* __restrict__ part_ninf,
                       const unsigned long long* __restrict__ part_nacc, int nparts,
                       const double* __restrict__ pop_next, int D, int64_t global_offset, int gen0,
                       int64_t generation, double* __restrict__ best_x, DeviceState* __restrict__ st,
                       RunState* __restrict__ rs, const __grid_constant__ ExchangeParams X, long long* hint) {
  finalize_body(part_e, part_idx, part_ninf, part_nacc, nparts, pop_next, D, global_offset, gen0, generation, best_x, st, rs, X, hint);
}

// Batch of independent solvers: one CTA per island (blockIdx.x), each with its own partials
// [isl*nparts, (isl+1)*nparts), rows, best member and DeviceState.
__global__ void __launch_bounds__(256)
    de_finalize_islands_kernel(const double* __restrict__ part_e, const int64_t* __restrict__ part_idx,
                               const unsigned long long* __restrict__ part_ninf,
                               const unsigned long long* __restrict__ part_nacc, int nparts,
                               const double* __restrict__ pop_next, int64_t island_np, int D, int64_t global_offset,
                               int gen0, int64_t generation, double* __restrict__ best_x, DeviceState* __restrict__ st,
                               RunState* __restrict__ rs) {
  const int64_t isl = blockIdx.x;
  ExchangeParams none;
  none.world = 1;
  finalize_body(part_e + isl * nparts, part_idx + isl * nparts, part_ninf + isl * nparts, part_nacc + isl * nparts, nparts,
                pop_next + isl * island_np * D, D, global_offset + isl * island_np, gen0, generation, best_x + isl * D,
                st + isl, rs != nullptr ? rs + isl : nullptr, none);
}

// batch of solvers with device-side termination: point every island's log at its block of the batch log
// [n_islands, max_steps, D+4] and clear a "log full" stop of the previous batch
__global__ void islands_arm_kernel(RunState* __restrict__ rs, int n, double* log, long long max_steps, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rs[i].log = log + (size_t)i * max_steps * (D + 4);
  rs[i].log_count = 0;
  rs[i].log_cap = max_steps;
  if (rs[i].stop == 2) rs[i].stop = 0;
}

// (generations logged in this batch, stop flag) of every island
__global__ void islands_status_kernel(const RunState* __restrict__ rs, int n, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[2 * i] = (int)rs[i].log_count;
  out[2 * i + 1] = rs[i].stop;
}

// Islands that share their best member within groups of `gsize` consecutive islands (blockIdx.x = group): after the
// per-island epilogue the group's winner -- lowest energy, ties to the lowest global index, the rule of the
// sharded best exchange above -- is installed in every island of the group.  A group is then EXACTLY a population
// sharded over `gsize` ranks (shard-local donors, global best), all on one GPU: what the statistics tests use to
// validate that deviation from the reference's global donor pool over many seeds in one launch per generation.
__global__ void __launch_bounds__(256)
    islands_share_best_kernel(int gsize, int D, double* __restrict__ best_x, DeviceState* __restrict__ st) {
  __shared__ int sh_w;
  const int first = blockIdx.x * gsize;
  if (threadIdx.x == 0) {
    int w = -1;
    double be = CUDART_INF;
    int64_t bi = 0;
    for (int r = 0; r < gsize; ++r) {
      const DeviceState& s = st[first + r];
      if (s.best_idx < 0) continue;  // island has no best yet
      if (w < 0 || s.best_e < be || (s.best_e == be && s.best_idx < bi)) {
        w = r;
        be = s.best_e;
        bi = s.best_idx;
      }
    }
    if (w >= 0)
      for (int r = 0; r < gsize; ++r) {
        st[first + r].best_e = be;
        st[first + r].best_idx = bi;
      }
    sh_w = w;
  }
  __syncthreads();
  const int w = sh_w;
  if (w < 0) return;
  const double* src = best_x + (size_t)(first + w) * D;
  for (int r = 0; r < gsize; ++r) {
    if (r == w) continue;
    double* dst = best_x + (size_t)(first + r) * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) dst[i] = src[i];
  }
}

// K4 (sparse generations, StepParams::sparse): the fused kernel wrote only the accepted trials (rows of `trial`,
// the other population buffer) and every row's new energy (`e_new`); this moves the accepted rows into the
// population IN PLACE -- the reference's `self.population[candidate][:] = self.trialSolution[candidate]`
// (differential_evolution.py:603-609) -- after every row of the generation has read its donors from the old
// one.  A row was accepted iff its new energy is lower (the fused kernel stores `acc ? E : e_old`), so a second
// run over the same buffers changes nothing: the no-op launches after a batch's stop flag are harmless.
// A CTA scans 256 energies at a time (a warp 32), queues the accepted rows in shared memory and copies them with all
// its warps, round robin -- a warp that copied only the rows it found itself would make the kernel as slow as the
// unluckiest warp (binomial tail: 43 us instead of ~15 at BASELINE config 2's shape with 13 % of the trials accepted).
__device__ __forceinline__ void commit_copy_row(double* __restrict__ pop, const double* __restrict__ trial, int64_t r, int D, int vec,
                                                int lane) {
  if (vec) {  // D even, 16-byte aligned buffers: eight 16-byte pieces per lane in flight (4 KB per warp)
    const double2* __restrict__ src = reinterpret_cast<const double2*>(trial + r * D);
    double2* __restrict__ dst = reinterpret_cast<double2*>(pop + r * D);
    const int n2 = D >> 1;
    for (int i = lane; i < n2; i += 256) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i + 32 * u < n2) v[u] = src[i + 32 * u];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i + 32 * u < n2) dst[i + 32 * u] = v[u];
    }
  } else {
    for (int i = lane; i < D; i += 128) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + 32 * u < D) v[u] = trial[r * D + i + 32 * u];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + 32 * u < D) pop[r * D + i + 32 * u] = v[u];
    }
  }
}

__global__ void __launch_bounds__(256)
    de_commit_accepted_kernel(const int* __restrict__ stop, double* __restrict__ pop, const double* __restrict__ trial,
                              double* __restrict__ e, const double* __restrict__ e_new, int64_t np, int D, int vec) {
  if (stop != nullptr && *stop) return;
  __shared__ int s_rows[256];  // accepted rows of the CTA's current 256 (offsets from its first row)
  __shared__ int s_count;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (vec && D <= 128) {
    // short rows (<= 64 16-byte pieces): every warp copies the rows it finds among its own 32, four at a time so
    // that four rows' loads are in flight before the first store needs its data (no CTA-wide barriers: with 2 M
    // rows of 32 dimensions the queue below costs more than the imbalance it removes)
    const int n2 = D >> 1;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = gw * 32; base < np; base += nw * 32) {
      const int64_t c = base + lane;
      bool a = false;
      double en = 0.0;
      if (c < np) {
        en = e_new[c];
        a = en < e[c];
      }
      unsigned m = __ballot_sync(0xffffffffu, a);
      if (a) e[c] = en;
      while (m) {
        double2 v[4][2];
        int64_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          r[k] = -1;
          if (m) {
            r[k] = base + (__ffs(m) - 1);
            m &= m - 1;
            const double2* __restrict__ src = reinterpret_cast<const double2*>(trial + r[k] * D);
            if (lane < n2) v[k][0] = src[lane];
            if (lane + 32 < n2) v[k][1] = src[lane + 32];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (r[k] >= 0) {
            double2* __restrict__ dst = reinterpret_cast<double2*>(pop + r[k] * D);
            if (lane < n2) dst[lane] = v[k][0];
            if (lane + 32 < n2) dst[lane + 32] = v[k][1];
          }
        }
      }
    }
    return;
  }
  for (int64_t first = (int64_t)blockIdx.x * 256; first < np; first += (int64_t)gridDim.x * 256) {
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int64_t c = first + threadIdx.x;
    bool a = false;
    double en = 0.0;
    if (c < np) {
      en = e_new[c];
      a = en < e[c];
    }
    const unsigned m = __ballot_sync(0xffffffffu, a);
    if (a) e[c] = en;
    int slot = 0;
    if (lane == 0 && m) slot = atomicAdd(&s_count, __popc(m));
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (a) s_rows[slot + __popc(m & ((1u << lane) - 1u))] = threadIdx.x;
    __syncthreads();
    const int n = s_count;
    for (int k = warp; k < n; k += 8) commit_copy_row(pop, trial, first + s_rows[k], D, vec, lane);
    __syncthreads();  // the queue is reused by the CTA's next 256 rows
  }
}

// K0: device-side SetRandomInitialPoints (abstract_solver.py:532-534): pop[c][j] = lo + (hi-lo)*u
__global__ void __launch_bounds__(256)
    de_init_population_kernel(double* __restrict__ pop, double* __restrict__ energy, int64_t np, int D,
                              int64_t global_offset, const double* __restrict__ lo, const double* __restrict__ hi,
                              uint32_t k0, uint32_t k1) {
  const int64_t nblk = (D + 1) / 2;  // philox blocks per row (2 doubles each)
  const int64_t total = np * nblk;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / nblk;
    const int b = (int)(t - c * nblk);
    const u32x4 v = philox4x32_10((uint32_t)b, (uint32_t)(global_offset + c), 0u, STREAM_INIT, k0, k1);
    const int j = 2 * b;
    const double u0 = u64_to_unit(((uint64_t)v.x << 32) | v.y);
    const double u1 = u64_to_unit(((uint64_t)v.z << 32) | v.w);
    pop[c * D + j] = dadd(lo[j], dmul(dsub(hi[j], lo[j]), u0));
    if (j + 1 < D) pop[c * D + j + 1] = dadd(lo[j + 1], dmul(dsub(hi[j + 1], lo[j + 1]), u1));
    if (b == 0) energy[c] = CUDART_INF;
  }
}

// self test: div_by_table(c, sqrt(i+1), 1/sqrt(i+1)) against the IEEE divide on n Philox-random
// operands; counts bit mismatches
__global__ void selftest_division_kernel(int64_t n, uint32_t k0, uint32_t k1, double span, const double2* tab, int D,
                                         unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const u32x4 v = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 7u, 9u, k0, k1);
    const double u = u64_to_unit(((uint64_t)v.x << 32) | v.y);
    const double c = (2.0 * u - 1.0) * span;
    const int i = (int)(v.z % (uint32_t)D);
    const double2 sr = tab[i];
    const double want = c / sqrt((double)i + 1.0);
    const double got = div_by_table(c, sr.x, sr.y);
    if (__double_as_longlong(want) != __double_as_longlong(got)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

// max |cos_bounded(a) - cos(a)| over n arguments uniform in [-span, span] (bits of the maximum,
// positive doubles order like their bit patterns)
__global__ void selftest_cosine_kernel(int64_t n, uint32_t k0, uint32_t k1, double span, unsigned long long* max_bits) {
  double worst = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const u32x4 v = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), 11u, 13u, k0, k1);
    const double u = u64_to_unit(((uint64_t)v.x << 32) | v.y);
    const double a = (2.0 * u - 1.0) * span;
    const double e = fabs(cos_bounded(a) - cos(a));
    if (!(e <= worst)) worst = e;  // NaN sticks
  }
  atomicMax(max_bits, (unsigned long long)__double_as_longlong(worst));
}

// does every element of the rows lie inside the strict range?  (*flag |= 1 when one does not; NaN counts as outside)
__global__ void rows_outside_range_kernel(const double* __restrict__ rows, int64_t n, int D, const double* __restrict__ lo,
                                          const double* __restrict__ hi, int uniform, double lo_s, double hi_s, int* flag) {
  bool bad = false;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const double x = rows[t];
    const int j = (int)(t % D);
    const double l = uniform ? lo_s : __ldg(lo + j), h = uniform ? hi_s : __ldg(hi + j);
    bad |= !(x >= l && x <= h);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void fill_kernel(double* __restrict__ p, int64_t n, double v) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = v;
}

// multi-GPU best exchange records: [best_e, best_idx(global), n_inf, n_acc, x[D]]
__global__ void pack_best_kernel(const DeviceState* __restrict__ st, const double* __restrict__ best_x, int D,
                                 double* __restrict__ rec) {
  if (threadIdx.x == 0) {
    rec[0] = st->best_e;
    rec[1] = (double)st->best_idx;
    rec[2] = (double)st->n_inf;
    rec[3] = (double)st->n_acc;
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) rec[4 + i] = best_x[i];
}

// exchange alone (after a restored state, or whenever the host asks for it)
__global__ void __launch_bounds__(256)
    exchange_best_kernel(int D, double* __restrict__ best_x, DeviceState* __restrict__ st, const __grid_constant__ ExchangeParams X) {
  (void)exchange_best(X, D, best_x, st);
}

__global__ void merge_best_kernel(const double* __restrict__ recs, int world, int D, double* __restrict__ best_x,
                                  DeviceState* __restrict__ st) {
  __shared__ int sh_w;
  if (threadIdx.x == 0) {
    int w = -1;
    double be = CUDART_INF;
    double bi = 0.0;
    double ninf = 0.0, nacc = 0.0;
    for (int r = 0; r < world; ++r) {
      const double* rec = recs + (int64_t)r * (D + 4);
      ninf += rec[2];
      nacc += rec[3];
      if (rec[1] < 0.0) continue;  // shard has no best yet
      if (w < 0 || rec[0] < be || (rec[0] == be && rec[1] < bi)) {
        w = r;
        be = rec[0];
        bi = rec[1];
      }
    }
    if (w >= 0) {
      st->best_e = be;
      st->best_idx = (int64_t)bi;
    }
    st->n_inf = (int64_t)ninf;
    st->n_acc = (int64_t)nacc;
    sh_w = w;
  }
  __syncthreads();
  const int w = sh_w;
  if (w >= 0) {
    const double* rec = recs + (int64_t)w * (D + 4);
    for (int i = threadIdx.x; i < D; i += blockDim.x) best_x[i] = rec[4 + i];
  }
}

}  // namespace mb2
