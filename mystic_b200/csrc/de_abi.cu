// C ABI of the engine (include/mystic_b200.h): context management, launch geometry and the
// dispatch of the fused generation kernel over <base vector, crossover, cost, vector width>.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/mystic_b200.h"
#include "de_kernels.cuh"
#include "de_misc.cuh"
#include "de_tma.cuh"
#include "de_tma_pool.cuh"
#include "de_cheb_mma.cuh"
#include "de_small.cuh"
#include "de_subwarp.cuh"
#include "de_rows.cuh"
#include "de_rows_bulk.cuh"

using namespace mb2;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(MB2_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Every entry point runs on its context's device and leaves the caller's current device as it found it
// (PyTorch, or another engine of the process, may be working on a different GPU).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    else if (err == cudaSuccess) prev = -1;  // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ENTER_DEVICE(dev)                                                                                  \
  DeviceGuard _device_guard(dev);                                                                          \
  if (_device_guard.err != cudaSuccess)                                                                    \
    return fail(MB2_ECUDA, "cudaSetDevice(%d) failed: %s", (int)(dev), cudaGetErrorString(_device_guard.err))

struct StrategyInfo {
  int base, xover, k;
};

bool strategy_info(int s, StrategyInfo* out) {
  switch (s) {
    case MB2_BEST1EXP: *out = {BASE_BEST1, XOVER_EXP, 2}; return true;
    case MB2_BEST1BIN: *out = {BASE_BEST1, XOVER_BIN, 2}; return true;
    case MB2_RAND1EXP:
    case MB2_RAND1BIN: *out = {BASE_RAND1, XOVER_EXP, 3}; return true;
    case MB2_RANDTOBEST1EXP:
    case MB2_RANDTOBEST1BIN: *out = {BASE_RANDTOBEST1, XOVER_EXP, 2}; return true;
    case MB2_BEST2EXP:
    case MB2_BEST2BIN: *out = {BASE_BEST2, XOVER_EXP, 4}; return true;
    case MB2_RAND2EXP:
    case MB2_RAND2BIN: *out = {BASE_RAND2, XOVER_EXP, 5}; return true;
  }
  return false;
}

}  // namespace

struct mb2_ctx {
  int device = 0;
  int sm_count = 0;
  int64_t np = 0, np_global = 0, offset = 0;
  int dim = 0, dim_pad = 0;
  double* pop[2] = {nullptr, nullptr};
  double* en[2] = {nullptr, nullptr};
  int cur = 0;
  int strategy = MB2_BEST1BIN;
  StrategyInfo sinfo = {BASE_BEST1, XOVER_BIN, 2};
  double F = 0.8, CR = 0.9;
  // exponential crossover length table of the current CR (philox.cuh, protocol 2): T[1..nmax] on the device
  uint32_t* xlen_tab = nullptr;
  int xlen_nmax = -1;
  float xlen_c = 0.f;
  uint64_t xlen_thr = ~0ull;  // threshold the table was built for
  int bounds_mode = BOUNDS_NONE;
  int bounds_uniform = 0;
  // every row of the current population lies inside [lo, hi]: generation 0 clipped it with these
  // bounds and nothing has replaced rows or bounds since (lets kernels skip re-checking parents)
  int rows_in_range = 0;
  int rows_checked = 0;   // the rows have been tested against the current bounds since either last changed
  int* range_flag = nullptr;
  double lo_s = 0.0, hi_s = 0.0;
  double* lo = nullptr;
  double* hi = nullptr;
  int cost = MB2_COST_ROSEN;
  CostParams cp = {0, nullptr, 0.0, 0.0, nullptr, 0, nullptr, 0, nullptr, 1.0, 0, 0, nullptr, 0, {0.0, 0.0, 0.0}};
  double* fit_obs = nullptr;
  size_t cap_fit_obs = 0;
  double* fit_sigma_v = nullptr;
  size_t cap_fit_sigma_v = 0;
  double bounds_absmax = 0.0;  // max(|lo_i|, |hi_i|) of the strict range
  double* cheb_vfrag = nullptr;
  int cheb_ks = 0;
  double* cheb_x = nullptr;
  double2* grw_tab = nullptr;
  PenaltyParams pp = {0, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* pen_type = nullptr;
  double* pen_mult = nullptr;
  double* pen_G = nullptr;
  double* pen_h = nullptr;
  double* pen_k = nullptr;
  uint64_t seed = 0;
  int replay = 0;
  const int32_t* r_donors = nullptr;
  const int32_t* r_nstart = nullptr;
  const double* r_uni = nullptr;
  double* best_x = nullptr;
  DeviceState* st = nullptr;
  double* part_e = nullptr;
  int64_t* part_idx = nullptr;
  unsigned long long* part_ninf = nullptr;
  unsigned long long* part_nacc = nullptr;
  int parts_cap = 0;
  double* cap_trial = nullptr;
  double* cap_energy = nullptr;
  DeviceState* h_st = nullptr;  // pinned
  double* h_best_x = nullptr;   // pinned
  int64_t generation = -1;      // last generation evaluated
  int64_t launches = 0;
  // launch geometry of the fused kernel
  int wpc = 8, grid = 0, smem = 0;
  int family = MB2_KERNEL_NONE;  // kernel family of the last fused launch
  // small device / pinned-host buffers come out of a caller-provided workspace when there is one
  // (torch's caching allocators make that free; cudaMalloc / cudaMallocHost cost milliseconds
  // each next to gigabytes of live allocations) and from cudaMalloc otherwise
  char* ws_dev = nullptr;
  size_t ws_dev_cap = 0, ws_dev_off = 0;
  char* ws_host = nullptr;
  size_t ws_host_cap = 0, ws_host_off = 0;
  std::vector<void*> owned_dev, owned_host;
  size_t cap_cheb_x = 0, cap_vfrag = 0, cap_pen = 0;
  double* init_lo = nullptr;
  double* init_hi = nullptr;
  // batch mode (mb2_run): device-resident termination state, its host description and mirror
  RunState* run = nullptr;
  RunState* h_run = nullptr;  // pinned
  int run_active = 0;         // launches carry the stop flag / run state
  int run_cur0 = 0;
  int64_t run_gen0 = 0;
  int run_enqueued = 0;
  // peer-memory best exchange (sharded population): this rank's mailbox, the peers' mailboxes as
  // mapped here, the epoch of the last exchange
  ExchangeParams xp = {};
  double* mailbox = nullptr;
  size_t mailbox_bytes = 0;
  void* peer_maps[kMaxWorld] = {};
  // batch of independent solvers (mb2_set_islands): n_islands populations of np / n_islands rows
  // side by side in the buffers, each with its own best member and DeviceState
  int n_islands = 1;
  int island_group = 1;  // islands share their best member within groups of this many consecutive islands
  // device-side termination of a batch of solvers: one RunState per island
  RunState* run_isl = nullptr;
  int run_isl_cap = 0;
  int run_isl_set = 0;      // mb2_set_termination has installed a condition for the islands
  int* isl_status = nullptr;    // device [2 * n_islands]
  int* h_isl_status = nullptr;  // pinned
  DeviceState* st_isl = nullptr;
  double* best_x_isl = nullptr;
  DeviceState* h_st_isl = nullptr;  // pinned
  double* h_best_isl = nullptr;     // pinned
  int xchg_connected = 0;
  int counts_global = 0;  // DeviceState counters are sums over the shards (peer exchange or mb2_merge_best)
  // how a generation leaves the fused kernel: every row into the other buffer + swap (double buffer), or accepted
  // rows only + de_commit_accepted_kernel (in place); auto = in place while the accepted fraction the host last
  // saw is below kInPlaceMaxAccept (a rejected row then costs no write, an accepted one three row transfers)
  int commit_mode = MB2_COMMIT_AUTO;
  double acc_hint = -1.0;  // accepted fraction of the last generation read back (-1: none yet)
  // the epilogue kernel also writes the accepted count of every generation into pinned host memory (no
  // synchronisation: the value is a few launches old at worst and only steers a choice between two equivalent paths),
  // so `auto` follows the run even when the host never reads a result (mb2_step loops)
  long long* h_hint = nullptr;  // pinned, mapped: accepted count of the last generation finalised, -1 = none yet
  long long* d_hint = nullptr;  // the same location as the device sees it
  int batch_in_place = -1;      // mb2_run: the decision taken for the whole batch (-1: none)
  int last_sparse = 0;     // the last fused launch wrote accepted rows only
  int run_sparse = 0;      // ... and so did the launches of the current batch (mb2_run)
  int xchg_no_adopt = 0;  // after mb2_exchange_reset: open a fresh mailbox
  int* h_xerr = nullptr;  // pinned
};

namespace {

// Mailboxes of destroyed contexts are parked here with their peer mappings: the next context of
// this process with the same (device, world, rank, slot size) adopts them instead of repeating
// cudaMalloc + the IPC handle round trip + cudaIpcOpenMemHandle per peer (milliseconds).  Every
// rank runs the same sequence of solvers, so all ranks adopt or none does.
struct ParkedChannel {
  int device;
  ExchangeParams xp;
  double* mailbox;
  size_t bytes;
  void* peer_maps[kMaxWorld];
};
std::mutex g_park_mutex;
std::vector<ParkedChannel> g_parked;

int dev_alloc(mb2_ctx* c, size_t bytes, void** out) {
  const size_t need = (bytes + 255) & ~(size_t)255;
  if (c->ws_dev != nullptr && c->ws_dev_off + need <= c->ws_dev_cap) {
    *out = c->ws_dev + c->ws_dev_off;
    c->ws_dev_off += need;
    return MB2_OK;
  }
  CUDA_TRY(cudaMalloc(out, need));
  c->owned_dev.push_back(*out);
  return MB2_OK;
}

int host_alloc(mb2_ctx* c, size_t bytes, void** out) {
  const size_t need = (bytes + 63) & ~(size_t)63;
  if (c->ws_host != nullptr && c->ws_host_off + need <= c->ws_host_cap) {
    *out = c->ws_host + c->ws_host_off;
    c->ws_host_off += need;
    return MB2_OK;
  }
  CUDA_TRY(cudaMallocHost(out, need));
  c->owned_host.push_back(*out);
  return MB2_OK;
}

// exponential crossover: the length table of the geometric draw for CrossProbability `cross`
// (philox.cuh, protocol 2), rebuilt when the probability changes
int ensure_exp_length_table(mb2_ctx* c, double cross) {
  const uint64_t thr = crossover_threshold(cross);
  if (thr == c->xlen_thr) return MB2_OK;
  ENTER_DEVICE(c->device);
  std::vector<uint32_t> tab((size_t)c->dim);
  const int nmax = build_exp_length_table(thr, c->dim, tab.data());
  if (c->xlen_tab == nullptr) {
    void* p = nullptr;
    if (int rc = dev_alloc(c, sizeof(uint32_t) * (size_t)c->dim, &p)) return rc;
    c->xlen_tab = (uint32_t*)p;
  }
  if (nmax > 0) {
    if (c->generation >= 0) CUDA_TRY(cudaDeviceSynchronize());  // kernels in flight may still read the old table
    CUDA_TRY(cudaMemcpy(c->xlen_tab, tab.data(), sizeof(uint32_t) * (size_t)nmax, cudaMemcpyHostToDevice));
  }
  c->xlen_nmax = nmax;
  // first guess of the length: log2(2^32 / w) / -log2(CR), CR as the kernels see it (thr / 2^32)
  const double l2 = (thr > 0) ? -std::log2((double)thr / 4294967296.0) : INFINITY;
  c->xlen_c = (thr > 0 && l2 > 0.0) ? (float)(1.0 / l2) : 0.f;
  c->xlen_thr = thr;
  return MB2_OK;
}

// kernel instantiations live in their own translation units (inst_*.cu) so that they compile in
// parallel; these getters return the host stubs
typedef void (*StepKernel)(const StepParams);
typedef void (*TmaKernel)(const StepParams, int);
typedef void (*ThreadRowKernel)(const StepParams, int, int, int);
typedef void (*ExpTmaKernel)(const StepParams, int, int, int, int);
}  // namespace
namespace mb2 {
StepKernel get_ldg_kernel(int base, int xover, int cost, bool vec);
TmaKernel get_tma_kernel(int cost, int gwarps, int threads);
typedef void (*PoolKernel)(const StepParams, int, int, int);
PoolKernel get_pool_kernel(int cost, int gwarps);  // K2p: de_tma_pool.cuh
ExpTmaKernel get_exp_tma_kernel_even(int cost, int gwarps, int threads);
ExpTmaKernel get_exp_tma_kernel_odd(int cost, int gwarps, int threads);  // rows of odd length
inline ExpTmaKernel get_exp_tma_kernel(int cost, int gwarps, int threads, int dim) {
  return (dim & 1) ? get_exp_tma_kernel_odd(cost, gwarps, threads) : get_exp_tma_kernel_even(cost, gwarps, threads);
}
ThreadRowKernel get_small_kernel(int cost);
ThreadRowKernel get_cheb_mma_kernel(int ksteps, bool resid);
StepKernel get_datafit_ldg_kernel(int base, int xover, int cost);
StepKernel get_separable_ldg_kernel(int base, int xover, bool vec);
ThreadRowKernel get_separable_small_kernel();
ThreadRowKernel get_separable_subwarp_kernel(int lanes_per_row, bool vec);
ThreadRowKernel get_subwarp_islands_kernel(int cost, int lanes_per_row, bool vec);
ThreadRowKernel get_datafit_subwarp_islands_kernel(int cost);
ThreadRowKernel get_separable_subwarp_islands_kernel(int lanes_per_row, bool vec);
ThreadRowKernel get_datafit_small_kernel(int cost);
ThreadRowKernel get_datafit_subwarp_kernel(int cost);
ThreadRowKernel get_subwarp_kernel(int cost, int lanes_per_row, bool vec);
ThreadRowKernel get_rows_kernel_rosen(int lanes_per_row, int ndonors);
ThreadRowKernel get_rows_kernel_corana(int lanes_per_row, int ndonors);
ThreadRowKernel get_rows_kernel_griewangk(int lanes_per_row, int ndonors);
typedef void (*BulkRowsKernel)(const StepParams, int, const BulkRowsLayout);
BulkRowsKernel get_rows_bulk_kernel_rosen(int lanes_per_row, int ndonors);
BulkRowsKernel get_rows_bulk_kernel_corana(int lanes_per_row, int ndonors);
BulkRowsKernel get_rows_bulk_kernel_griewangk(int lanes_per_row, int ndonors);
}  // namespace mb2
namespace {

bool is_datafit(int cost) { return cost == MB2_COST_POLYFIT || cost == MB2_COST_LORENTZIAN; }

bool is_separable(int cost) { return cost >= MB2_COST_SPHERE && cost <= MB2_COST_POWERS; }

StepKernel pick_kernel(const StrategyInfo& s, int cost, bool vec) {
  if (is_datafit(cost)) return get_datafit_ldg_kernel(s.base, s.xover, cost);
  if (is_separable(cost)) return get_separable_ldg_kernel(s.base, s.xover, vec);
  return get_ldg_kernel(s.xover == XOVER_BIN ? BASE_BEST1 : s.base, s.xover, cost, vec);
}

// warps per CTA and dynamic shared memory for a row length
void geometry(int dim_pad, int* wpc, int* smem) {
  const size_t row = (size_t)dim_pad * sizeof(double);
  int w = kWarpsPerCtaMax;
  // keep >= 2 CTAs per SM resident when the rows allow it (227 KB per SM)
  while (w > 1 && (size_t)(w + 1) * row > 100 * 1024) w >>= 1;
  *wpc = w;
  *smem = (int)((size_t)(w + 1) * row);
}

int ensure_parts(mb2_ctx* c, int n) {
  if (n <= c->parts_cap) return MB2_OK;
  int cap = c->sm_count * 32;  // at most 32 resident CTAs per SM: persistent grids never exceed this
  if (cap < n) cap = n;
  void* p = nullptr;
  if (int rc = dev_alloc(c, sizeof(double) * cap, &p)) return rc;
  c->part_e = (double*)p;
  if (int rc = dev_alloc(c, sizeof(int64_t) * cap, &p)) return rc;
  c->part_idx = (int64_t*)p;
  if (int rc = dev_alloc(c, sizeof(unsigned long long) * cap, &p)) return rc;
  c->part_ninf = (unsigned long long*)p;
  if (int rc = dev_alloc(c, sizeof(unsigned long long) * cap, &p)) return rc;
  c->part_nacc = (unsigned long long*)p;
  c->parts_cap = cap;
  return MB2_OK;
}

TmaKernel pick_tma_kernel(int cost, int gwarps, int threads) { return get_tma_kernel(cost, gwarps, threads); }

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// in-place commit pays while fewer than a third of the trials are accepted: per row it moves
// (1 + k + a) rows + 2a in the commit against (2 + k) with the double buffer
constexpr double kInPlaceMaxAccept = 0.30;
// ... for long rows; rows of up to 128 dimensions leave their kernels a stage of several rows at a time with ONE
// copy request, which in place becomes a request per accepted row: measured break-even near 0.15 (bench.py g128 at
// 0.145: 0.678 ms either way; r32 at 0.178: 0.440 ms in place, 0.417 ms with the double buffer)
constexpr double kInPlaceMaxAcceptShortRows = 0.12;
bool want_in_place(const mb2_ctx* c) {
  int mode = env_int("MB2_COMMIT_MODE", -1);
  if (mode < MB2_COMMIT_AUTO || mode > MB2_COMMIT_IN_PLACE) mode = c->commit_mode;
  if (mode == MB2_COMMIT_DOUBLE_BUFFER) return false;
  if (mode == MB2_COMMIT_IN_PLACE) return true;
  if (c->batch_in_place >= 0) return c->batch_in_place != 0;  // one decision per batch of generations (mb2_run)
  double hint = c->acc_hint;
  if (c->h_hint != nullptr) {
    const long long n = *reinterpret_cast<volatile long long*>(c->h_hint);
    const int64_t denom = c->xchg_connected ? c->np_global : c->np;  // (the fused exchange sums the shards' counts)
    if (n >= 0 && denom > 0) hint = (double)n / (double)denom;
  }
  return hint < (c->dim > 128 ? kInPlaceMaxAccept : kInPlaceMaxAcceptShortRows);  // (no generation finalised yet: in place)
}
void reset_hint(mb2_ctx* c) {
  c->acc_hint = -1.0;
  if (c->h_hint != nullptr) *c->h_hint = -1;
}
void note_accepted(mb2_ctx* c, int64_t n_acc) {
  const int64_t denom = (c->xchg_connected || c->counts_global) ? c->np_global : c->np;
  c->acc_hint = (denom > 0 && n_acc >= 0) ? (double)n_acc / (double)denom : -1.0;
}

// Ring geometry of the TMA kernel: groups per CTA, warps per group, stages per group, dynamic
// smem.  Returns false when a row does not fit (the LDG kernel is used instead).
// Measured on B200 at D=1024 (profiles/README.md): concurrency across rows matters most (8 groups
// beat 4 groups with deeper rings), then warps per row, then stages.
bool tma_geometry(int dim, bool in_place, int* groups, int* gwarps, int* stages, int* smem) {
  const size_t row = (((size_t)(dim + (dim & 1)) * 8) + 127) & ~(size_t)127;  // (rows of odd length: the 16-byte aligned cover, dim + 1 elements)
  const size_t budget = 220 * 1024;
  if (row * 4 > budget) return false;
  const size_t slots = (budget - row) / (3 * row);  // (group, stage) buffers that fit on one SM
  int g = kTmaMaxGroups;
  while (g > 1 && (size_t)g > slots) g >>= 1;
  // warps per row: about 8 dimensions per thread (B200 sweeps: D=1024 -> 4 warps, D=256 -> 1 warp;
  // more warps per row only add barrier and per-warp overhead once the CTA is full)
  int gw = 1;
  while (gw < 8 && g * (gw * 2) * 32 <= kTmaMaxThreads && (gw * 2) * 32 * 8 <= dim) gw <<= 1;
  // Generations committed in place are no longer bound by HBM but by the groups' own instruction streams: half as
  // many warps per row when that brings the CTA down to 512 threads, i.e. into the 128-register build (BASELINE
  // config 2, 8 groups: 2 warps per row 0.295 / 0.273 ms sustained / burst against 0.305 / 0.278 with 4)
  if (in_place && gw >= 2 && g * gw * 32 > 512 && g * (gw / 2) * 32 <= 512) gw >>= 1;
  if (gw > 1 && g > 15) g = 15;  // groups of several warps synchronise on named barriers 1..15
  // One-warp groups (D < 512): 16 groups in a 512-thread CTA -- the 128-register build; the 64 registers of a
  // 1024-thread CTA make the griewangk body spill -- with the ring as deep as shared memory allows.  B200 sweep at
  // BASELINE config 5 (profiles/README.md): 16 x 1 warp x 2 stages 3.25-3.32 ms, 24 x 1 x 1 (768 threads, 80 registers)
  // 3.35-3.37, 32 x 1 x 1 (round 1's choice) 3.57-3.59.
  if (gw == 1 && g > 16) g = 16;
  int st = 1;
  while (st < kTmaMaxStages && (size_t)g * (st * 2) <= slots) st *= 2;
  const int eg = env_int("MB2_TMA_GROUPS", 0), ew = env_int("MB2_TMA_GWARPS", 0), es = env_int("MB2_TMA_STAGES", 0);
  if (eg >= 1 && eg <= kTmaMaxGroups && (ew == 1 || (ew == 2 || ew == 4 || ew == 8) && eg <= 15) && es >= 1 &&
      es <= kTmaMaxStages && (size_t)eg * es <= slots && eg * ew * 32 <= kTmaMaxThreads) {
    g = eg;
    gw = ew;
    st = es;
  }
  *groups = g;
  *gwarps = gw;
  *stages = st;
  *smem = (int)(row + (size_t)g * st * 3 * row);
  return true;
}

// Pool geometry of K2p (de_tma_pool.cuh: in-place Best1Bin generations of rows whose K2a ring has one stage): P pairs +
// Q first-donor buffers out of the row buffers that fit beside bestSolution; 8 consumer groups of 2 warps.
bool pool_geometry(int dim, int* pairs, int* d0s, int* groups, int* gwarps, int* nprod, int* smem) {
  if (env_int("MB2_TMA_POOL", 1) == 0 || dim % 2 != 0 || dim > 1536) return false;
  // rows of 256 .. 766 dimensions (K2a: one-warp groups with rings of two stages, BASELINE config 5): one-warp consumer
  // groups and two producer warps
  const bool short_rows = dim < 768;
  if (short_rows && (dim < 256 || env_int("MB2_POOL_SHORT", 0) == 0)) return false;
  const size_t row = (((size_t)dim * 8) + 127) & ~(size_t)127;
  const size_t budget = 218 * 1024;
  int slots = (int)((budget - row) / row);
  if (slots > 2 * kPoolMaxRing) slots = 2 * kPoolMaxRing;
  // a pair lives from its request to the end of the trial loop, a first-donor buffer until its row is selected; B200 sweep
  // at 1024 dimensions (26 buffers, tools/k2p_ab.sh): 7 groups, 7 pairs + 12 buffers 0.237 ms, 8 / 7 + 12 0.240, 7 / 8 + 10
  // 0.242, 8 / 8 + 10 0.246, 8 / 6 + 14 0.251, 6 / 7 + 12 0.257 (K2a: 0.274); groups of 4 warps: 4 / 9 + 8 0.290
  int q = (slots * 12 + 13) / 26, p = (slots - q) / 2;
  int gw = short_rows ? 1 : 2;
  const int egw = env_int("MB2_POOL_GW", 0);
  if (egw == 1 || egw == 2 || egw == 4) gw = egw;
  int np_ = 2;   // (config 2: 0.2384 ms with two producer warps, 0.2438 with one)
  const int enp = env_int("MB2_POOL_NPROD", 0);
  if (enp == 1 || enp == 2) np_ = enp;
  const int max_warps = (gw == 1) ? 20 : 17;   // (the builds' __launch_bounds__: 640 / 544 threads; 5 warps per scheduler = 96 registers)
  int g = gw == 4 ? 4 : gw == 2 ? 7 : 17;
  const int ep = env_int("MB2_POOL_P", 0), eq = env_int("MB2_POOL_Q", 0), egp = env_int("MB2_POOL_G", 0);
  if (ep >= 1 && eq >= 1 && 2 * ep + eq <= slots && ep <= kPoolMaxRing && eq <= kPoolMaxRing) p = ep, q = eq;
  if (egp >= 1) g = egp;
  while (g > 1 && g * gw + np_ > max_warps) --g;
  if (gw > 1 && g > 15) g = 15;   // named barriers 1..15
  if (g > kPoolMaxGroups) g = kPoolMaxGroups;
  // a consumer group looks at the `full` barrier of its NEXT row (k + G, slot (k + G) mod Q) as soon as it is done with
  // row k; the parity test only tells phase n from phase n - 1, so that slot's previous row (k + G - Q) should have been
  // issued by then (Q >= G; the row tag in the slot catches what is left)
  if (g > q) g = q;
  if (p > kPoolMaxRing) p = kPoolMaxRing;
  if (q > kPoolMaxRing) q = kPoolMaxRing;
  if (p < 2 || q < 2) return false;
  *pairs = p;
  *d0s = q;
  *groups = g;
  *gwarps = gw;
  *nprod = np_;
  *smem = (int)(row * (size_t)(1 + 2 * p + q));
  return true;
}

// Ring geometry of K2e (exponential crossover, de_tma_exp.cuh).  A stage holds the parent row, an
// undo buffer and k donor-segment buffers of `cap` elements each; cap covers the crossover window
// (plus alignment slack) with probability >= 0.999: P(L > n) = CR^n.  In-flight bytes per SM are what
// matter (a stage carries one row), so the ring is as deep as the shared memory allows.
bool exp_tma_geometry(int dim_in, int k, double CR, bool in_place, int* groups, int* gwarps, int* stages, int* cap, int* smem) {
  // rows of odd length: buffers hold 16-byte aligned covers -- dim + 1 elements for a row, up to two more than the
  // window for a donor segment
  const int dim = dim_in + (dim_in & 1);
  const size_t row = (((size_t)dim * 8) + 127) & ~(size_t)127;
  const size_t budget = 220 * 1024;
  int want = dim;
  if (CR < 1.0) {
    const double n = (CR > 0.0) ? std::ceil(std::log(1e-3) / std::log(CR)) : 1.0;
    if (n + 3.0 + 2.0 * (dim_in & 1) < (double)dim) want = (int)n + 3 + 2 * (dim_in & 1);
  }
  want = (want + 1) & ~1;
  if (want < 8) want = 8;
  if (want > dim) want = dim;
  const int ec = env_int("MB2_XTMA_CAP", 0);
  if (ec >= 2 && ec <= dim) want = ec & ~1;
  size_t stage = (row + (size_t)(k + 1) * want * 8 + 128 + 127) & ~(size_t)127;
  if (row + 2 * stage > budget) {  // whole-row donor buffers do not fit twice: short windows only
    want = 64 < dim ? 64 : dim;
    stage = (row + (size_t)(k + 1) * want * 8 + 128 + 127) & ~(size_t)127;
    if (row + 2 * stage > budget) return false;
  }
  const size_t slots = (budget - row) / stage;  // (group, stage) buffers that fit on one SM
  // B200 sweep at D=1024 (profiles/README.md): rows in flight matter most -- 15 groups x 2 warps x 1 stage
  // 0.210 ms, 8 groups x 4 warps x 2 stages 0.301 ms.  As many groups as fit, then >= 16 dimensions per thread.
  // Generations committed in place (no row leaves a stage unless its trial is accepted) are bound by the rows in
  // flight alone: >= 32 dimensions per thread, i.e. 20 one-warp groups at D=1024 -- 0.183 ms against 0.207 ms with
  // 15 x 2 (with the double buffer the same geometry loses: 0.217 against 0.210).
  const int per_thread = in_place ? 32 : 16;
  int g = (int)slots;
  if (g > kTmaMaxGroups) g = kTmaMaxGroups;
  int gw = 1;
  while (gw < 8 && (gw * 2) * 32 * per_thread <= dim && (g > 15 ? 15 : g) * (gw * 2) * 32 <= kTmaMaxThreads) gw <<= 1;
  if (gw > 1 && g > 15) g = 15;  // groups of several warps synchronise on named barriers 1..15
  if (g * gw * 32 > kTmaMaxThreads) g = kTmaMaxThreads / (gw * 32);
  int st = (int)(slots / g);
  if (st > kTmaMaxStages) st = kTmaMaxStages;
  const int eg = env_int("MB2_XTMA_GROUPS", 0), ew = env_int("MB2_XTMA_GWARPS", 0), es = env_int("MB2_XTMA_STAGES", 0);
  if (eg >= 1 && eg <= kTmaMaxGroups && (ew == 1 || (ew == 2 || ew == 4 || ew == 8) && eg <= 15) && es >= 1 &&
      es <= kTmaMaxStages && (size_t)eg * es <= slots && eg * ew * 32 <= kTmaMaxThreads) {
    g = eg;
    gw = ew;
    st = es;
  }
  *groups = g;
  *gwarps = gw;
  *stages = st;
  *cap = want;
  *smem = (int)(row + (size_t)g * st * stage);
  return true;
}

// Geometry of K2f (de_rows_bulk.cuh): warps per CTA (one CTA per SM), stages per warp, slots (4-element
// chunks) of a stage's donor pools = what shared memory is left after the row tiles.  False when the
// pools would be too small to hold typical windows.
bool rows_bulk_geometry(int dim, int lpr, int k, int* warps, BulkRowsLayout* L, int* smem) {
  const int rpw = 32 / lpr;
  const size_t best = (((size_t)dim * 8) + 127) & ~(size_t)127;
  const size_t rows = (size_t)rpw * dim * 8, meta = (size_t)rpw * sizeof(BulkRowMeta);
  const size_t budget = 220 * 1024 - best;
  // B200 sweep at C4 (profiles/README.md): the step is bound by the warps' own dependent chains, so the
  // number of resident warps decides (8 warps 0.53 ms, 12: 0.40, 16: 0.35-0.37, 20: 0.34, 24: 0.32); a
  // second stage per warp does not pay for the pool space it takes.
  int w = env_int("MB2_RB_WARPS", kBulkRowsMaxWarps), st = env_int("MB2_RB_STAGES", 1);
  if (w < 1 || w > kBulkRowsMaxWarps) w = kBulkRowsMaxWarps;
  if (st < 1 || st > kBulkRowsMaxStages) st = 1;
  const size_t avail = (budget / w) & ~(size_t)127;
  const size_t fixed = (size_t)st * (rows + meta + 128) + 128;
  if (avail <= fixed) return false;
  int slots = (int)((avail - fixed) / (32 * (size_t)(1 + st * k)));
  const int all = rpw * (dim >> 2);  // every chunk of every row
  if (slots > all) slots = all;
  const int es = env_int("MB2_RB_POOL", 0);
  if (es >= 1 && es < slots) slots = es;
  if (slots < 2 * rpw && es == 0) return false;
  const size_t pool_b = (size_t)slots * 32;
  L->stages = st;
  L->slots = slots;
  L->best_bytes = (unsigned)best;
  L->rows_bytes = (unsigned)rows;
  L->pool_bytes = (unsigned)pool_b;
  L->meta_off = (unsigned)(rows + (size_t)k * pool_b);
  L->stage_bytes = (unsigned)((rows + (size_t)k * pool_b + meta + 127) & ~(size_t)127);
  L->undo_bytes = (unsigned)((pool_b + 127) & ~(size_t)127);
  L->warp_bytes = L->undo_bytes + (unsigned)st * L->stage_bytes;
  L->donors_ldgsts = env_int("MB2_RB_LDGSTS", 0) != 0 ? 1 : 0;  // measured on B200 at C4: 0.303 ms vs 0.297 ms with bulk requests (profiles/README.md)
  *warps = w;
  *smem = (int)(best + (size_t)w * L->warp_bytes);
  return true;
}

int ensure_parts(mb2_ctx* c, int n);
void fill_params(mb2_ctx* c, int mode, const double* in_rows, double* out_rows, const double* e_in, double* e_out, int64_t np,
                 StepParams* out);

// batch of independent solvers: generation 0 / one generation of every island in ONE launch
// (grid.y = island) of the rows-per-warp kernel with run-time strategy (rows of <= 128 dimensions)
int launch_step_islands(mb2_ctx* c, int mode, cudaStream_t stream) {
  const int64_t inp = c->np / c->n_islands;
  if (c->cost == MB2_COST_CHEBYSHEV_MMA || c->cost == MB2_COST_POLYFIT_MMA)
    return fail(MB2_EINVAL, "batched solvers use the Horner / elementwise cost variants");
  const double* in_rows = c->pop[c->cur];
  double* out_rows = c->pop[c->cur ^ 1];
  const bool vec = (c->dim % 2 == 0) && ((reinterpret_cast<uintptr_t>(in_rows) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out_rows) & 15) == 0);
  int lpr = 4;
  while (lpr < 32 && lpr * 8 < c->dim) lpr <<= 1;
  ThreadRowKernel k = is_datafit(c->cost) ? get_datafit_subwarp_islands_kernel(c->cost)
                                          : (is_separable(c->cost) ? get_separable_subwarp_islands_kernel(lpr, vec)
                                                                   : get_subwarp_islands_kernel(c->cost, lpr, vec));
  if (k == nullptr) return fail(MB2_EINVAL, "no batched kernel for cost %d", c->cost);
  const int rows_per_cta = kSubWarps * (32 / lpr);
  const size_t smem_need = sizeof(double) * (size_t)(1 + kSubWarps * (32 / lpr)) * c->dim_pad;
  // rows longer than 256 dimensions take a whole warp each (lpr = 32): the best member and 8 rows live in shared memory
  if (smem_need > 227 * 1024) return fail(MB2_EINVAL, "batched solvers support rows of at most %d dimensions", (227 * 1024 / 72) & ~3);
  const int smem = (int)smem_need;
  CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int64_t gx = (inp + rows_per_cta - 1) / rows_per_cta;
  if (gx > 64) gx = 64;  // the CTAs of an island stride over its rows
  if (int rc = ensure_parts(c, (int)(gx * c->n_islands))) return rc;
  StepParams P;
  fill_params(c, mode, in_rows, out_rows, c->en[c->cur], c->en[c->cur ^ 1], inp, &P);
  P.best_x = c->best_x_isl;
  P.rows_in_range = 0;
  P.stop = (c->run_active && c->run_isl != nullptr) ? &c->run_isl[0].stop : nullptr;  // per island: stride sizeof(RunState)
  c->wpc = kSubWarps;
  c->grid = (int)gx;
  c->smem = smem;
  c->family = MB2_KERNEL_ISLANDS;
  c->last_sparse = 0;
  const StrategyInfo s = c->sinfo;
  k<<<dim3((unsigned)gx, (unsigned)c->n_islands), kSubWarps * 32, smem, stream>>>(P, s.base, s.xover, s.k);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  return MB2_OK;
}

// configure + launch the fused kernel in `mode` over `np` rows
int launch_step(mb2_ctx* c, int mode, const double* in_rows, double* out_rows, const double* e_in, double* e_out,
                int64_t np, cudaStream_t stream) {
  if (c->n_islands > 1 && mode != MODE_EVAL) return launch_step_islands(c, mode, stream);
  // Rows that were installed by the caller (restart files, SetPopulation) or bounds that changed after
  // generation 0 (SetStrictRanges mid-run): test the population once; when every row lies inside the range
  // the kernels may treat the untouched elements of a trial as in range, as after a generation-0 clip
  if (mode == MODE_STEP && c->bounds_mode != BOUNDS_NONE && !c->rows_in_range && !c->rows_checked && in_rows == c->pop[c->cur]) {
    if (c->range_flag == nullptr) {
      void* p = nullptr;
      if (int rc = dev_alloc(c, sizeof(int), &p)) return rc;
      c->range_flag = (int*)p;
    }
    CUDA_TRY(cudaMemsetAsync(c->range_flag, 0, sizeof(int), stream));
    const int64_t n = np * c->dim;
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)c->sm_count * 16) g = (int64_t)c->sm_count * 16;
    rows_outside_range_kernel<<<(unsigned)g, 256, 0, stream>>>(in_rows, n, c->dim, c->lo, c->hi, c->bounds_uniform, c->lo_s, c->hi_s,
                                                                c->range_flag);
    CUDA_TRY(cudaGetLastError());
    ++c->launches;
    int outside = 1;
    CUDA_TRY(cudaMemcpyAsync(&outside, c->range_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    c->rows_checked = 1;
    c->rows_in_range = outside ? 0 : 1;
  }
  const bool al16 = ((reinterpret_cast<uintptr_t>(in_rows) & 15) == 0) && (out_rows == nullptr || (reinterpret_cast<uintptr_t>(out_rows) & 15) == 0);
  const bool vec = (c->dim % 2 == 0) && al16;
  // the staged long-row kernels also take rows of odd length (16-byte aligned covers, de_tma.cuh); MB2_TMA_ODD=0
  // sends those to the warp-per-row kernel as before
  const bool bulk_ok = al16 && (c->dim % 2 == 0 || env_int("MB2_TMA_ODD", 1) != 0);
  StrategyInfo s = c->sinfo;
  if (mode != MODE_STEP) s = {BASE_RAND1, XOVER_EXP, 3};  // no best/donors needed; any instance
  const bool cheb_mma = (c->cost == MB2_COST_CHEBYSHEV_MMA || c->cost == MB2_COST_POLYFIT_MMA);
  const bool datafit = is_datafit(c->cost);
  const bool separable = is_separable(c->cost);
  StepKernel k = cheb_mma ? nullptr : pick_kernel(s, c->cost, vec);
  if (k == nullptr && !cheb_mma) return fail(MB2_EINVAL, "no kernel for strategy %d cost %d", c->strategy, c->cost);
  ThreadRowKernel kc = cheb_mma ? get_cheb_mma_kernel(c->cheb_ks, c->cost == MB2_COST_POLYFIT_MMA) : nullptr;
  // short rows: one thread per candidate (all modes, so energies of a context share one summation order)
  ThreadRowKernel ks = nullptr;
  int small_max = env_int("MB2_SMALL_MAXD", kSmallMaxDim);
  if (small_max > kSmallMaxDim) small_max = kSmallMaxDim;
  // MB2_SMALL_KERNEL: 3 = K2d rows kernel where it applies, else as 1 (default); 1 = several rows
  // per warp (generic); 2 = one thread per row; 0 = none of them
  const int small_kind = env_int("MB2_SMALL_KERNEL", 3);
  ThreadRowKernel kw = nullptr, kr = nullptr;
  BulkRowsKernel kb = nullptr;
  BulkRowsLayout blay = {};
  int bw = 0, bsmem = 0;
  int lpr = 4;
  while (lpr < 32 && lpr * 8 < c->dim) lpr <<= 1;  // about 8 dimensions per lane
  s = c->sinfo;  // the short-row kernels take the strategy at run time in every mode
  if (!cheb_mma && c->dim <= small_max) {
    if (small_kind == 3 && c->dim % 4 == 0 && c->dim >= kRowsMinDim && c->dim <= kRowsMaxDim && vec && np <= 0x7fffffff) {
      lpr = (c->dim <= 64) ? 4 : 8;  // 16 dimensions per lane
      if (c->cost == MB2_COST_ROSEN) kr = get_rows_kernel_rosen(lpr, s.k);
      else if (c->cost == MB2_COST_CORANA) kr = get_rows_kernel_corana(lpr, s.k);
      else if (c->cost == MB2_COST_GRIEWANGK) kr = get_rows_kernel_griewangk(lpr, s.k);
    }
    // exponential crossover over rows known to lie in the strict range: K2f moves the same tiles with the
    // bulk-copy engine, a stage ahead of the arithmetic
    if (kr != nullptr && mode == MODE_STEP && s.xover == XOVER_EXP && env_int("MB2_ROWS_BULK", 1) != 0 &&
        (c->bounds_mode == BOUNDS_NONE || (in_rows == c->pop[c->cur] && c->rows_in_range)) &&
        rows_bulk_geometry(c->dim, lpr, s.k, &bw, &blay, &bsmem)) {
      if (c->cost == MB2_COST_ROSEN) kb = get_rows_bulk_kernel_rosen(lpr, s.k);
      else if (c->cost == MB2_COST_CORANA) kb = get_rows_bulk_kernel_corana(lpr, s.k);
      else if (c->cost == MB2_COST_GRIEWANGK) kb = get_rows_bulk_kernel_griewangk(lpr, s.k);
    }
    if (kr != nullptr) kw = kr;
    else if (small_kind == 2)
      ks = datafit ? get_datafit_small_kernel(c->cost) : (separable ? get_separable_small_kernel() : get_small_kernel(c->cost));
    else if (small_kind == 1 || small_kind == 3)
      kw = datafit ? get_datafit_subwarp_kernel(c->cost)
                   : (separable ? get_separable_subwarp_kernel(lpr, vec) : get_subwarp_kernel(c->cost, lpr, vec));
  }

  // a generation of the context's own buffers may be committed in place (decided before the geometries: K2a / K2e / K2p
  // size their rings for it)
  const bool in_place_req = mode == MODE_STEP && in_rows == c->pop[c->cur] && out_rows == c->pop[c->cur ^ 1] && e_in == c->en[c->cur] &&
                            e_out == c->en[c->cur ^ 1] && want_in_place(c);
  // Best1Bin reads whole donor rows: stage them with TMA bulk copies when rows are 16-B multiples
  TmaKernel kt = nullptr;
  PoolKernel kp = nullptr;
  int tg = 0, tw = 0, ts = 0, tsmem = 0, pool_p = 0, pool_q = 0, pool_np = 1;
  if (mode == MODE_STEP && s.xover == XOVER_BIN && bulk_ok && env_int("MB2_TMA", 1) != 0 && !datafit &&
      !cheb_mma && ks == nullptr && kw == nullptr && tma_geometry(c->dim, in_place_req, &tg, &tw, &ts, &tsmem))
    kt = pick_tma_kernel(c->cost, tw, tg * tw * 32);
  // in place, one-stage rings: the pooled producer / consumer kernel
  if (kt != nullptr && in_place_req && (ts == 1 || c->dim < 768) && pool_geometry(c->dim, &pool_p, &pool_q, &tg, &tw, &pool_np, &tsmem)) {
    kp = get_pool_kernel(c->cost, tw);
    if (kp == nullptr) tma_geometry(c->dim, in_place_req, &tg, &tw, &ts, &tsmem);  // (no instantiation for this cost: K2a's geometry again)
  }
  // the other strategies read a short window of the donors: K2e keeps long rows in a TMA ring
  ExpTmaKernel kx = nullptr;
  int xcap = 0;
  if (mode == MODE_STEP && s.xover == XOVER_EXP && bulk_ok && env_int("MB2_TMA_EXP", 1) != 0 && !datafit && !cheb_mma &&
      ks == nullptr && kw == nullptr && exp_tma_geometry(c->dim, s.k, c->CR, in_place_req, &tg, &tw, &ts, &xcap, &tsmem))
    kx = get_exp_tma_kernel(c->cost, tw, tg * tw * 32, c->dim);
  // generation 0 of long rows is a read-clip-evaluate-write stream: the same staged kernel without draws or donors
  // (every strategy; the warp-per-row kernel needed 0.31 ms for it at BASELINE config 2)
  int gen0_k = 0;
  if (mode == MODE_GEN0 && bulk_ok && env_int("MB2_TMA_GEN0", 1) != 0 && !datafit && !cheb_mma && ks == nullptr && kw == nullptr &&
      out_rows != nullptr && exp_tma_geometry(c->dim, 2, 0.0, false, &tg, &tw, &ts, &xcap, &tsmem)) {
    kx = get_exp_tma_kernel(c->cost, tw, tg * tw * 32, c->dim);
    gen0_k = 2;
  }

  int wpc, smem;
  int occ = 0;
  int rows_per_cta;
  if (kc != nullptr) {
    wpc = kChebWarps;
    rows_per_cta = kChebWarps * kChebBatch;
    smem = (int)(sizeof(double) * (size_t)kChebWarps * (kChebBatch * c->dim + kChebBatch));
    CUDA_TRY(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kc, wpc * 32, smem));
  } else if (kb != nullptr) {
    wpc = bw;
    rows_per_cta = bw * (32 / lpr);
    smem = bsmem;
    CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kb, wpc * 32, smem));
  } else if (kr != nullptr) {
    wpc = kRowsWarps;
    rows_per_cta = kRowsWarps * (32 / lpr);
    smem = (int)(sizeof(double) * ((size_t)c->dim + (size_t)kRowsWarps * 2 * (32 / lpr) * (c->dim + 2)));
    CUDA_TRY(cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kr, wpc * 32, smem));
  } else if (kw != nullptr) {
    wpc = kSubWarps;
    rows_per_cta = kSubWarps * (32 / lpr);
    smem = (int)(sizeof(double) * (size_t)(1 + kSubWarps * (32 / lpr)) * c->dim_pad);
    CUDA_TRY(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kw, wpc * 32, smem));
  } else if (ks != nullptr) {
    wpc = kSmallWarps;
    rows_per_cta = kSmallWarps * 32;
    smem = (int)(sizeof(double) * (size_t)kSmallWarps * 32 * (c->dim | 1));
    CUDA_TRY(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ks, wpc * 32, smem));
  } else if (kp != nullptr) {
    wpc = tg * tw + pool_np;  // (+ the producer warps)
    rows_per_cta = tg;
    smem = tsmem;
    CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kp, wpc * 32, smem));
  } else if (kt != nullptr) {
    wpc = tg * tw;
    rows_per_cta = tg;
    smem = tsmem;
    CUDA_TRY(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kt, wpc * 32, smem));
  } else if (kx != nullptr) {
    wpc = tg * tw;
    rows_per_cta = tg;
    smem = tsmem;
    CUDA_TRY(cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kx, wpc * 32, smem));
  } else {
    geometry(c->dim_pad, &wpc, &smem);
    rows_per_cta = wpc;
    if (smem > 227 * 1024) return fail(MB2_EINVAL, "dim %d too large for one warp-resident row", c->dim);
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, wpc * 32, smem));
  }
  if (occ < 1) return fail(MB2_EINVAL, "kernel does not fit on an SM (smem %d)", smem);
  int64_t need = (np + rows_per_cta - 1) / rows_per_cta;
  int64_t grid = (int64_t)c->sm_count * occ;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (int rc = ensure_parts(c, (int)grid)) return rc;

  StepParams P;
  fill_params(c, mode, in_rows, out_rows, e_in, e_out, np, &P);
  // the staged kernels can leave rejected rows where they are (the caller then commits the accepted ones in place)
  // (kernels that honour StepParams::sparse: K2f, K2d, K2a, K2e and the warp-per-row kernel K2b; the launch order
  // below decides which one runs)
  const bool by_rows = kc == nullptr && (kb != nullptr || (kw != nullptr && kw == kr));
  const bool by_groups = kc == nullptr && kb == nullptr && kw == nullptr && ks == nullptr;   // kt, kx or k
  const bool staged = by_rows || (by_groups && !(kt == nullptr && kx != nullptr && gen0_k));
  const bool sparse = in_place_req && staged;
  P.sparse = sparse ? 1 : 0;
  c->last_sparse = P.sparse;
  c->wpc = wpc;
  c->grid = (int)grid;
  c->smem = smem;
  c->family = kc ? MB2_KERNEL_DMMA : kb ? MB2_KERNEL_ROWS_BULK : kr ? MB2_KERNEL_ROWS : kw ? MB2_KERNEL_ROWS_PER_WARP : ks ? MB2_KERNEL_THREAD_PER_ROW
              : kp ? MB2_KERNEL_TMA_POOL : kt ? MB2_KERNEL_TMA_BIN : kx ? MB2_KERNEL_TMA_EXP : MB2_KERNEL_WARP_PER_ROW;
  if (kc != nullptr)
    kc<<<(unsigned)grid, wpc * 32, smem, stream>>>(P, s.base, s.xover, s.k);
  else if (kb != nullptr)
    kb<<<(unsigned)grid, wpc * 32, smem, stream>>>(P, s.base, blay);
  else if (kw != nullptr)
    kw<<<(unsigned)grid, wpc * 32, smem, stream>>>(P, s.base, s.xover, s.k);
  else if (ks != nullptr)
    ks<<<(unsigned)grid, wpc * 32, smem, stream>>>(P, s.base, s.xover, s.k);
  else if (kp != nullptr)
    kp<<<(unsigned)grid, wpc * 32, smem, stream>>>(P, pool_p, pool_q, pool_np);
  else if (kt != nullptr)
    kt<<<(unsigned)grid, wpc * 32, smem, stream>>>(P, ts);
  else if (kx != nullptr)
    kx<<<(unsigned)grid, wpc * 32, smem, stream>>>(P, ts, xcap, gen0_k ? BASE_RAND1 : s.base, gen0_k ? gen0_k : s.k);
  else
    k<<<(unsigned)grid, wpc * 32, smem, stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  return MB2_OK;
}

ExchangeParams next_exchange(mb2_ctx* c) {
  ExchangeParams X = c->xp;
  if (c->xchg_connected) {
    X.epoch = ++c->xp.epoch;
  } else {
    X.world = 1;
  }
  return X;
}

void fill_params(mb2_ctx* c, int mode, const double* in_rows, double* out_rows, const double* e_in, double* e_out, int64_t np,
                 StepParams* out) {
  StepParams P;
  memset(&P, 0, sizeof(P));
  P.pop_cur = in_rows;
  P.pop_next = out_rows;
  P.e_cur = e_in;
  P.e_next = e_out;
  P.best_x = c->best_x;
  P.lo = c->lo;
  P.hi = c->hi;
  P.np = np;
  P.global_offset = c->offset;
  P.dim = c->dim;
  P.dim_pad = c->dim_pad;
  P.mode = mode;
  P.bounds_mode = c->bounds_mode;
  P.bounds_uniform = c->bounds_uniform;
  P.rows_in_range = (mode == MODE_STEP && in_rows == c->pop[c->cur]) ? c->rows_in_range : 0;
  P.lo_s = c->lo_s;
  P.hi_s = c->hi_s;
  P.F = c->F;
  P.CR = c->CR;
  P.replay = (mode == MODE_STEP) ? c->replay : 0;
  P.key0 = (uint32_t)(c->seed & 0xffffffffu);
  P.key1 = (uint32_t)(c->seed >> 32);
  P.generation = (uint32_t)(c->generation + 1);
  P.thr16 = crossover_threshold16(c->CR);
  P.xlen = {c->xlen_tab, c->xlen_nmax, c->xlen_c};
  P.r_donors = c->r_donors;
  P.r_nstart = c->r_nstart;
  P.r_uni = c->r_uni;
  P.cost = c->cp;
  // rows inside a finite strict range (generation 0 clips, trials are clamped or rejected; the
  // energy of an out-of-range row is inf whatever its cost evaluates to): guard-free griewangk factors
  P.cost.grw_bounded = (c->bounds_mode != BOUNDS_NONE && c->bounds_absmax <= kCosFastMax) ? 1 : 0;
  P.pen = c->pp;
  P.part_e = c->part_e;
  P.part_idx = c->part_idx;
  P.part_ninf = c->part_ninf;
  P.part_nacc = c->part_nacc;
  P.cap_trial = c->cap_trial;
  P.cap_energy = c->cap_energy;
  P.stop = (c->run_active && mode == MODE_STEP) ? &c->run->stop : nullptr;

  *out = P;
}

// sparse generation: accepted rows and energies from the other buffers into the current ones, in place
int launch_commit(mb2_ctx* c, cudaStream_t stream) {
  double* pop = c->pop[c->cur];
  const double* trial = c->pop[c->cur ^ 1];
  const int vec = ((c->dim % 2 == 0) && ((reinterpret_cast<uintptr_t>(pop) & 15) == 0) && ((reinterpret_cast<uintptr_t>(trial) & 15) == 0)) ? 1 : 0;
  int64_t grid = (c->np + 255) / 256;  // a CTA per 256 rows
  const int64_t cap = (int64_t)c->sm_count * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  de_commit_accepted_kernel<<<(unsigned)grid, 256, 0, stream>>>(c->run_active ? &c->run->stop : nullptr, pop, trial, c->en[c->cur],
                                                                c->en[c->cur ^ 1], c->np, c->dim, vec);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  return MB2_OK;
}

int launch_finalize(mb2_ctx* c, int gen0, cudaStream_t stream) {
  if (c->n_islands > 1) {
    de_finalize_islands_kernel<<<(unsigned)c->n_islands, 256, 0, stream>>>(c->part_e, c->part_idx, c->part_ninf, c->part_nacc, c->grid,
                                                                           c->pop[c->cur ^ 1], c->np / c->n_islands, c->dim, c->offset, gen0,
                                                                           c->generation + 1, c->best_x_isl, c->st_isl,
                                                                           c->run_active ? c->run_isl : nullptr);
    CUDA_TRY(cudaGetLastError());
    ++c->launches;
    if (c->island_group > 1) {
      islands_share_best_kernel<<<(unsigned)(c->n_islands / c->island_group), 256, 0, stream>>>(c->island_group, c->dim, c->best_x_isl,
                                                                                               c->st_isl);
      CUDA_TRY(cudaGetLastError());
      ++c->launches;
    }
    return MB2_OK;
  }
  de_finalize_kernel<<<1, 256, 0, stream>>>(c->part_e, c->part_idx, c->part_ninf, c->part_nacc, c->grid,
                                            c->pop[c->cur ^ 1], c->dim, c->offset, gen0, c->generation + 1,
                                            c->best_x, c->st, (c->run_active && !gen0) ? c->run : nullptr,
                                            next_exchange(c), gen0 ? nullptr : c->d_hint);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  return MB2_OK;
}

}  // namespace

extern "C" {

int mb2_abi_version(void) { return MB2_ABI_VERSION; }

const char* mb2_last_error(void) { return g_err; }

int64_t mb2_workspace_bytes(int32_t dim, int64_t* host_bytes) {
  if (host_bytes) *host_bytes = 4096 + (int64_t)sizeof(double) * dim;
  // state, best, bounds (x4 rows), partials for 148*32 CTAs, tables for the common costs
  return 262144 + 72 * (int64_t)dim;
}

int mb2_create_ws(int device, int64_t np_local, int32_t dim, int64_t np_global, int64_t global_offset,
                  void* dev_workspace, int64_t dev_bytes, void* host_workspace, int64_t host_bytes, mb2_ctx** out) {
  if (out == nullptr) return fail(MB2_EINVAL, "out is NULL");
  *out = nullptr;
  if (np_local < 1 || dim < 1 || np_global < np_local || global_offset < 0 || global_offset + np_local > np_global)
    return fail(MB2_EINVAL, "bad shape np_local=%lld dim=%d np_global=%lld offset=%lld", (long long)np_local, dim,
                (long long)np_global, (long long)global_offset);
  if (np_global > 0xffffffffll) return fail(MB2_EINVAL, "np_global exceeds the 32-bit Philox candidate counter");
  ENTER_DEVICE(device);
  mb2_ctx* c = new (std::nothrow) mb2_ctx();
  if (c == nullptr) return fail(MB2_EINVAL, "out of host memory");
  c->device = device;
  c->np = np_local;
  c->np_global = np_global;
  c->offset = global_offset;
  c->dim = dim;
  c->dim_pad = (dim + 3) & ~3;
  if (dev_workspace != nullptr && dev_bytes > 0) {
    const uintptr_t a = (reinterpret_cast<uintptr_t>(dev_workspace) + 255) & ~(uintptr_t)255;
    c->ws_dev = reinterpret_cast<char*>(a);
    c->ws_dev_cap = (size_t)dev_bytes - (size_t)(a - reinterpret_cast<uintptr_t>(dev_workspace));
  }
  if (host_workspace != nullptr && host_bytes > 0) {
    const uintptr_t a = (reinterpret_cast<uintptr_t>(host_workspace) + 63) & ~(uintptr_t)63;
    c->ws_host = reinterpret_cast<char*>(a);
    c->ws_host_cap = (size_t)host_bytes - (size_t)(a - reinterpret_cast<uintptr_t>(host_workspace));
  }
  // (cudaGetDeviceProperties costs milliseconds; one attribute is all that is needed)
  int sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) {
    delete c;
    return fail(MB2_ECUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  }
  c->sm_count = sms;
  void* p = nullptr;
  int rc = MB2_OK;
  if (!rc) rc = dev_alloc(c, sizeof(double) * dim, &p), c->best_x = (double*)p;
  if (!rc) rc = dev_alloc(c, sizeof(DeviceState), &p), c->st = (DeviceState*)p;
  if (!rc) rc = dev_alloc(c, sizeof(double) * dim, &p), c->lo = (double*)p;
  if (!rc) rc = dev_alloc(c, sizeof(double) * dim, &p), c->hi = (double*)p;
  if (!rc) rc = dev_alloc(c, sizeof(double) * dim, &p), c->init_lo = (double*)p;
  if (!rc) rc = dev_alloc(c, sizeof(double) * dim, &p), c->init_hi = (double*)p;
  if (!rc) rc = host_alloc(c, sizeof(DeviceState), &p), c->h_st = (DeviceState*)p;
  if (!rc) {
    void* hp = nullptr;
    void* dp = nullptr;
    if (host_alloc(c, 64, &hp) == MB2_OK && cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
      c->h_hint = (long long*)hp;
      c->d_hint = (long long*)dp;
      *c->h_hint = -1;
    } else {
      cudaGetLastError();  // (pinned memory the device cannot address: `auto` follows the host's reads only)
    }
  }
  if (!rc) rc = host_alloc(c, sizeof(double) * dim, &p), c->h_best_x = (double*)p;
  if (!rc) rc = ensure_parts(c, 1);
  if (!rc) rc = ensure_exp_length_table(c, c->CR);
  if (!rc) {
    DeviceState init = {INFINITY, -1, 0, 0, -1};
    if (cudaMemcpy(c->st, &init, sizeof(init), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(c->best_x, 0, sizeof(double) * dim) != cudaSuccess)
      rc = fail(MB2_ECUDA, "context initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc) {
    mb2_destroy(c);
    return rc;
  }
  *out = c;
  return MB2_OK;
}

int mb2_create(int device, int64_t np_local, int32_t dim, int64_t np_global, int64_t global_offset, mb2_ctx** out) {
  return mb2_create_ws(device, np_local, dim, np_global, global_offset, nullptr, 0, nullptr, 0, out);
}

int mb2_destroy(mb2_ctx* c) {
  if (c == nullptr) return MB2_OK;
  DeviceGuard _device_guard(c->device);
  if (c->mailbox != nullptr) {
    int err = 1;
    if (c->xchg_connected && cudaDeviceSynchronize() == cudaSuccess &&
        cudaMemcpy(&err, c->xp.error, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess && err == 0) {
      ParkedChannel pc;
      pc.device = c->device;
      pc.xp = c->xp;
      pc.mailbox = c->mailbox;
      pc.bytes = c->mailbox_bytes;
      memcpy(pc.peer_maps, c->peer_maps, sizeof(pc.peer_maps));
      std::lock_guard<std::mutex> lock(g_park_mutex);
      g_parked.push_back(pc);
    } else {
      for (int r = 0; r < kMaxWorld; ++r)
        if (c->peer_maps[r] != nullptr) cudaIpcCloseMemHandle(c->peer_maps[r]);
      cudaFree(c->mailbox);
    }
  }
  if (c->run_isl != nullptr) cudaFree(c->run_isl);
  for (void* p : c->owned_dev) cudaFree(p);
  for (void* p : c->owned_host) cudaFreeHost(p);
  delete c;
  return MB2_OK;
}

int mb2_bind_buffers(mb2_ctx* c, double* pop_a, double* pop_b, double* e_a, double* e_b) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (!pop_a || !pop_b || !e_a || !e_b || pop_a == pop_b || e_a == e_b)
    return fail(MB2_EINVAL, "need four distinct device buffers");
  c->pop[0] = pop_a;
  c->pop[1] = pop_b;
  c->en[0] = e_a;
  c->en[1] = e_b;
  c->cur = 0;
  c->generation = -1;
  c->rows_in_range = 0;
  c->rows_checked = 0;
  return MB2_OK;
}

int mb2_current(mb2_ctx* c, double** pop, double** energy) {
  if (c == nullptr || c->pop[0] == nullptr) return fail(MB2_ESTATE, "no buffers bound");
  if (pop) *pop = c->pop[c->cur];
  if (energy) *energy = c->en[c->cur];
  return MB2_OK;
}

int mb2_set_strategy(mb2_ctx* c, int strategy, double scale, double cross) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  StrategyInfo s;
  if (!strategy_info(strategy, &s)) return fail(MB2_EINVAL, "unknown strategy id %d", strategy);
  if (c->np < s.k + 1)
    return fail(MB2_EINVAL, "strategy needs %d distinct donors but the shard has only %lld other members", s.k,
                (long long)(c->np - 1));
  if (int rc = ensure_exp_length_table(c, cross)) return rc;
  c->strategy = strategy;
  c->sinfo = s;
  c->F = scale;
  c->CR = cross;
  return MB2_OK;
}

int mb2_set_bounds(mb2_ctx* c, int mode, const double* lo, const double* hi) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (mode < BOUNDS_NONE || mode > BOUNDS_CLAMP) return fail(MB2_EINVAL, "unknown bounds mode %d", mode);
  ENTER_DEVICE(c->device);
  if (c->generation >= 0) CUDA_TRY(cudaDeviceSynchronize());  // kernels of a non-blocking stream may still read lo / hi
  if (mode != BOUNDS_NONE) {
    if (!lo || !hi) return fail(MB2_EINVAL, "bounds arrays are NULL");
    for (int i = 0; i < c->dim; ++i)
      if (lo[i] > hi[i]) return fail(MB2_EINVAL, "each min[i] must be <= the corresponding max[i]");
    CUDA_TRY(cudaMemcpy(c->lo, lo, sizeof(double) * c->dim, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->hi, hi, sizeof(double) * c->dim, cudaMemcpyHostToDevice));
    bool uni = true;
    for (int i = 1; i < c->dim; ++i) uni = uni && lo[i] == lo[0] && hi[i] == hi[0];
    c->bounds_uniform = uni ? 1 : 0;
    double am = 0.0;
    for (int i = 0; i < c->dim; ++i) am = fmax(am, fmax(fabs(lo[i]), fabs(hi[i])));
    c->bounds_absmax = am;
    c->lo_s = lo[0];
    c->hi_s = hi[0];
  }
  c->bounds_mode = mode;
  c->rows_in_range = 0;
  c->rows_checked = 0;
  return MB2_OK;
}

int mb2_set_cost(mb2_ctx* c, int cost, const double* params, int32_t nparams) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (cost < MB2_COST_ROSEN || cost > MB2_COST_DECAY) return fail(MB2_EINVAL, "unknown cost id %d", cost);
  ENTER_DEVICE(c->device);
  if (c->generation >= 0) CUDA_TRY(cudaDeviceSynchronize());  // kernels of a non-blocking stream may still read the cost tables
  c->cp.fit_kind = FIT_LORENTZIAN;
  if (cost == MB2_COST_DECAY) {  // a second elementwise forward model on the lorentzian instantiations
    if (c->dim != 5) return fail(MB2_EINVAL, "the dual exponential decay has 5 parameters (a1,a2,a3,a4,a5)");
    c->cp.fit_kind = FIT_DECAY;
    cost = MB2_COST_LORENTZIAN;
  }
  if (is_separable(cost)) c->cp.sep_kind = cost - MB2_COST_SPHERE;
  // closed-form functions one thread evaluates (fixed_cost, de_device.cuh) run on the zimmermann instantiations
  c->cp.fixed_kind = FIX_ZIMMERMANN;
  if (cost >= MB2_COST_BRANINS && cost <= MB2_COST_PAVIANI) {
    if (cost <= MB2_COST_SHEKEL && c->dim != 2) return fail(MB2_EINVAL, "cost %d is a function of two variables", cost);
    c->cp.fixed_kind = FIX_BRANINS + (cost - MB2_COST_BRANINS);
    const double pi = 3.141592653589793;  // math.pi
    c->cp.fixed_c[0] = 5.1 / (4 * pow(pi, 2));  // pohlheim.py:279, as Python evaluates it
    c->cp.fixed_c[1] = 5. / pi;
    c->cp.fixed_c[2] = 1. / (8 * pi);
    cost = MB2_COST_ZIMMERMANN;
  }
  if (cost == MB2_COST_LORENTZIAN && c->cp.fit_kind == FIT_LORENTZIAN && c->dim != 7) return fail(MB2_EINVAL, "lorentzian has 7 parameters (a1,a2,a3,A0,E0,G0,n)");
  if ((cost == MB2_COST_POLYFIT || cost == MB2_COST_LORENTZIAN) && c->dim > 32)
    return fail(MB2_EINVAL, "data-fit costs support at most 32 model parameters");
  if (cost == MB2_COST_POLYFIT_MMA && c->dim > kChebMaxDim)
    return fail(MB2_EINVAL, "polynomial-fit DMMA cost supports at most %d coefficients", kChebMaxDim);
  if (cost == MB2_COST_ZIMMERMANN && c->cp.fixed_kind == FIX_ZIMMERMANN && c->dim != 2)
    return fail(MB2_EINVAL, "zimmermann requires dim == 2");
  if (cost == MB2_COST_CHEBYSHEV_MMA && c->dim > kChebMaxDim)
    return fail(MB2_EINVAL, "chebyshev DMMA cost supports at most %d coefficients", kChebMaxDim);
  // evaluation points -> device; Vandermonde fragments of the DMMA variants (see de_cheb_mma.cuh)
  auto upload_points = [&](const std::vector<double>& xs, bool mma) -> int {
    const int M = (int)xs.size();
    if (c->cap_cheb_x < sizeof(double) * M) {
      void* p = nullptr;
      if (int rc = dev_alloc(c, sizeof(double) * M, &p)) return rc;
      c->cheb_x = (double*)p;
      c->cap_cheb_x = sizeof(double) * M;
    }
    CUDA_TRY(cudaMemcpy(c->cheb_x, xs.data(), sizeof(double) * M, cudaMemcpyHostToDevice));
    c->cp.cheb_M = M;
    c->cp.cheb_x = c->cheb_x;
    if (!mma) return MB2_OK;
    const int n = c->dim - 1;  // order of the trial polynomial
    const int KS = n <= 4 ? 1 : (n <= 8 ? 2 : 4);
    const int ntiles = (M + 7) / 8;
    std::vector<double> vf((size_t)ntiles * KS * 32, 0.0);
    for (int t = 0; t < ntiles; ++t)
      for (int s = 0; s < KS; ++s)
        for (int lane = 0; lane < 32; ++lane) {
          const int k = 4 * s + (lane & 3), i = 8 * t + (lane >> 2);
          double v = 0.0;
          if (k < n && i < M) {
            volatile double pw = 1.0;
            for (int e = 0; e < n - k; ++e) pw = pw * xs[i];
            v = pw;
          }
          vf[((size_t)t * KS + s) * 32 + lane] = v;
        }
    if (c->cap_vfrag < sizeof(double) * vf.size()) {
      void* p = nullptr;
      if (int rc = dev_alloc(c, sizeof(double) * vf.size(), &p)) return rc;
      c->cheb_vfrag = (double*)p;
      c->cap_vfrag = sizeof(double) * vf.size();
    }
    CUDA_TRY(cudaMemcpy(c->cheb_vfrag, vf.data(), sizeof(double) * vf.size(), cudaMemcpyHostToDevice));
    c->cp.cheb_vfrag = c->cheb_vfrag;
    c->cp.cheb_ntiles = ntiles;
    c->cheb_ks = KS;
    return MB2_OK;
  };
  if (cost == MB2_COST_POLYFIT || cost == MB2_COST_POLYFIT_MMA || cost == MB2_COST_LORENTZIAN) {
    if (!params || nparams < 4) return fail(MB2_EINVAL, "data-fit params = {M, sigma, evalpts[M], observations[M][, sigma[M]]}");
    const int M = (int)params[0];
    const double sigma = params[1];
    const bool per_point = (nparams == 2 + 3 * M);  // (x - observations) / sigma with a sigma per point (models/br8.py:47)
    if (M < 1 || (nparams != 2 + 2 * M && !per_point)) return fail(MB2_EINVAL, "bad data-fit params (M=%d, %d values)", M, nparams);
    if (!per_point && !(sigma != 0.0)) return fail(MB2_EINVAL, "sigma must not be zero");
    if (per_point && cost == MB2_COST_POLYFIT_MMA) return fail(MB2_EINVAL, "the DMMA polynomial fit takes a scalar sigma");
    c->cp.fit_sigma_v = nullptr;
    if (per_point) {
      for (int i = 0; i < M; ++i)
        if (!(params[2 + 2 * M + i] != 0.0)) return fail(MB2_EINVAL, "sigma[%d] must not be zero", i);
      if (c->cap_fit_sigma_v < sizeof(double) * M) {
        void* p = nullptr;
        if (int rc = dev_alloc(c, sizeof(double) * M, &p)) return rc;
        c->fit_sigma_v = (double*)p;
        c->cap_fit_sigma_v = sizeof(double) * M;
      }
      CUDA_TRY(cudaMemcpy(c->fit_sigma_v, params + 2 + 2 * M, sizeof(double) * M, cudaMemcpyHostToDevice));
      c->cp.fit_sigma_v = c->fit_sigma_v;
    }
    std::vector<double> xs(params + 2, params + 2 + M);
    if (int rc = upload_points(xs, cost == MB2_COST_POLYFIT_MMA)) return rc;
    if (c->cap_fit_obs < sizeof(double) * M) {
      void* p = nullptr;
      if (int rc = dev_alloc(c, sizeof(double) * M, &p)) return rc;
      c->fit_obs = (double*)p;
      c->cap_fit_obs = sizeof(double) * M;
    }
    CUDA_TRY(cudaMemcpy(c->fit_obs, params + 2 + M, sizeof(double) * M, cudaMemcpyHostToDevice));
    c->cp.fit_obs = c->fit_obs;
    c->cp.fit_sigma = per_point ? 1.0 : sigma;
  }
  if (cost == MB2_COST_CHEBYSHEV || cost == MB2_COST_CHEBYSHEV_MMA) {
    if (!params || nparams < 3) return fail(MB2_EINVAL, "chebyshev params = {M, order, target[order+1]}");
    const int M = (int)params[0], order = (int)params[1];
    if (M < 2 || order < 0 || nparams != order + 3) return fail(MB2_EINVAL, "bad chebyshev params (M=%d order=%d)", M, order);
    // abscissae exactly as the reference accumulates them (_model_helper.py:19-25)
    std::vector<double> xs(M);
    volatile double x = -1.0;
    const double dx = 2.0 / (double)(M - 1);
    for (int i = 0; i < M; ++i) {
      xs[i] = x;
      x = x + dx;
    }
    // polyeval(target, +-1.2), math/poly.py:37-39
    double tp, tm;
    for (int sgn = 0; sgn < 2; ++sgn) {
      const double xe = sgn ? -1.2 : 1.2;
      volatile double val = 0.0 * xe;
      for (int j = 0; j <= order; ++j) {
        volatile double prod = val * xe;
        val = params[2 + j] + prod;
      }
      if (sgn) tm = val; else tp = val;
    }
    if (int rc = upload_points(xs, cost == MB2_COST_CHEBYSHEV_MMA)) return rc;
    c->cp.cheb_tp = tp;
    c->cp.cheb_tm = tm;
  }
  if (cost == MB2_COST_GRIEWANGK && c->grw_tab == nullptr) {
    // {sqrt(i+1.0), 1/sqrt(i+1.0)}: glibc sqrt and the host divide are correctly rounded
    std::vector<double2> tab(c->dim);
    for (int i = 0; i < c->dim; ++i) {
      volatile double sq = std::sqrt((double)i + 1.0);
      volatile double rc = 1.0 / sq;
      tab[i].x = sq;
      tab[i].y = rc;
    }
    void* p = nullptr;
    if (int rc = dev_alloc(c, sizeof(double2) * c->dim, &p)) return rc;
    c->grw_tab = (double2*)p;
    CUDA_TRY(cudaMemcpy(c->grw_tab, tab.data(), sizeof(double2) * c->dim, cudaMemcpyHostToDevice));
    c->cp.grw_tab = c->grw_tab;
  }
  c->cost = cost;
  return MB2_OK;
}

int mb2_selftest_division(mb2_ctx* c, int64_t n, uint64_t seed, double span, int64_t* mismatches) {
  if (c == nullptr || mismatches == nullptr || n < 1) return fail(MB2_EINVAL, "bad argument");
  if (c->grw_tab == nullptr) return fail(MB2_ESTATE, "set the griewangk cost first");
  ENTER_DEVICE(c->device);
  unsigned long long* d = c->part_ninf;  // scratch: not in use between generations
  CUDA_TRY(cudaMemset(d, 0, sizeof(unsigned long long)));
  selftest_division_kernel<<<c->sm_count * 8, 256>>>(n, (uint32_t)seed, (uint32_t)(seed >> 32), span, c->grw_tab, c->dim, d);
  CUDA_TRY(cudaGetLastError());
  unsigned long long h = 0;
  CUDA_TRY(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
  ++c->launches;
  *mismatches = (int64_t)h;
  return MB2_OK;
}

int mb2_selftest_cosine(mb2_ctx* c, int64_t n, uint64_t seed, double span, double* max_abs_err) {
  if (c == nullptr || max_abs_err == nullptr || n < 1 || !(span <= kCosFastMax)) return fail(MB2_EINVAL, "bad argument");
  ENTER_DEVICE(c->device);
  if (int rc = ensure_parts(c, 1)) return rc;
  unsigned long long* d = c->part_ninf;  // scratch: not in use between generations
  CUDA_TRY(cudaMemset(d, 0, sizeof(unsigned long long)));
  selftest_cosine_kernel<<<c->sm_count * 8, 256>>>(n, (uint32_t)seed, (uint32_t)(seed >> 32), span, d);
  CUDA_TRY(cudaGetLastError());
  unsigned long long h = 0;
  CUDA_TRY(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
  ++c->launches;
  memcpy(max_abs_err, &h, sizeof(h));
  return MB2_OK;
}

int mb2_set_penalty_ex(mb2_ctx* c, int32_t nterms, const int32_t* ptype, const double* G, const double* h,
                       const double* kscale, const double* mult) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (nterms < 0) return fail(MB2_EINVAL, "nterms < 0");
  ENTER_DEVICE(c->device);
  if (c->generation >= 0) CUDA_TRY(cudaDeviceSynchronize());  // kernels of a non-blocking stream may still read the terms
  c->pp = {0, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (nterms == 0) return MB2_OK;
  if (!ptype || !G || !h || !kscale) return fail(MB2_EINVAL, "penalty arrays are NULL");
  for (int i = 0; i < nterms; ++i) {
    if (ptype[i] < MB2_PEN_QUADRATIC_EQUALITY || ptype[i] > MB2_PEN_LAGRANGE_EQUALITY)
      return fail(MB2_EINVAL, "unknown penalty type %d", ptype[i]);
    if (ptype[i] >= MB2_PEN_LAGRANGE_INEQUALITY && mult == nullptr)
      return fail(MB2_EINVAL, "lagrange penalties need their multipliers (mb2_set_penalty_ex)");
  }
  const size_t pen_bytes = ((sizeof(int) * nterms + 255) & ~(size_t)255) + sizeof(double) * nterms * (c->dim + 3);
  if (c->cap_pen < pen_bytes) {
    void* p = nullptr;
    if (int rc = dev_alloc(c, pen_bytes, &p)) return rc;
    c->pen_type = (int*)p;
    c->cap_pen = pen_bytes;
  }
  c->pen_G = reinterpret_cast<double*>(reinterpret_cast<char*>(c->pen_type) + ((sizeof(int) * nterms + 255) & ~(size_t)255));
  c->pen_h = c->pen_G + (size_t)nterms * c->dim;
  c->pen_k = c->pen_h + nterms;
  c->pen_mult = c->pen_k + nterms;
  std::vector<double> m((size_t)nterms, 0.0);
  if (mult != nullptr) m.assign(mult, mult + nterms);
  CUDA_TRY(cudaMemcpy(c->pen_type, ptype, sizeof(int) * nterms, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->pen_G, G, sizeof(double) * nterms * c->dim, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->pen_h, h, sizeof(double) * nterms, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->pen_k, kscale, sizeof(double) * nterms, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->pen_mult, m.data(), sizeof(double) * nterms, cudaMemcpyHostToDevice));
  c->pp = {nterms, c->pen_type, c->pen_G, c->pen_h, c->pen_k, c->pen_mult};
  return MB2_OK;
}

int mb2_set_penalty(mb2_ctx* c, int32_t nterms, const int32_t* ptype, const double* G, const double* h,
                    const double* kscale) {
  return mb2_set_penalty_ex(c, nterms, ptype, G, h, kscale, nullptr);
}

int mb2_set_rng_philox(mb2_ctx* c, uint64_t seed) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  c->seed = seed;
  c->replay = 0;
  return MB2_OK;
}

int mb2_set_rng_replay(mb2_ctx* c, const int32_t* donors, const int32_t* nstart, const double* uni) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (!donors || !nstart || !uni) return fail(MB2_EINVAL, "replay buffers are NULL");
  c->r_donors = donors;
  c->r_nstart = nstart;
  c->r_uni = uni;
  c->replay = 1;
  return MB2_OK;
}

int mb2_set_capture(mb2_ctx* c, double* trial, double* energy) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if ((trial == nullptr) != (energy == nullptr)) return fail(MB2_EINVAL, "capture needs both buffers or neither");
  c->cap_trial = trial;
  c->cap_energy = energy;
  return MB2_OK;
}

int mb2_init_population(mb2_ctx* c, uint64_t seed, const double* lo, const double* hi, void* stream) {
  if (c == nullptr || c->pop[0] == nullptr) return fail(MB2_ESTATE, "no buffers bound");
  if (!lo || !hi) return fail(MB2_EINVAL, "init bounds are NULL");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  double *dlo = c->init_lo, *dhi = c->init_hi;
  CUDA_TRY(cudaMemcpyAsync(dlo, lo, sizeof(double) * c->dim, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(dhi, hi, sizeof(double) * c->dim, cudaMemcpyHostToDevice, s));
  const int64_t total = c->np * ((c->dim + 1) / 2);
  int64_t grid = (total + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (grid > cap) grid = cap;
  de_init_population_kernel<<<(unsigned)grid, 256, 0, s>>>(c->pop[c->cur], c->en[c->cur], c->np, c->dim, c->offset, dlo,
                                                           dhi, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  CUDA_TRY(cudaStreamSynchronize(s));  // lo/hi are pageable host buffers
  c->generation = -1;
  c->rows_in_range = 0;
  c->rows_checked = 0;
  reset_hint(c);
  return MB2_OK;
}

int mb2_upload_population(mb2_ctx* c, const double* pop_host, void* stream) {
  if (c == nullptr || c->pop[0] == nullptr) return fail(MB2_ESTATE, "no buffers bound");
  if (!pop_host) return fail(MB2_EINVAL, "pop_host is NULL");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(c->pop[c->cur], pop_host, sizeof(double) * c->np * c->dim, cudaMemcpyHostToDevice, s));
  int64_t grid = (c->np + 255) / 256;
  if (grid > (int64_t)c->sm_count * 16) grid = (int64_t)c->sm_count * 16;
  fill_kernel<<<(unsigned)grid, 256, 0, s>>>(c->en[c->cur], c->np, INFINITY);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  CUDA_TRY(cudaStreamSynchronize(s));
  c->generation = -1;
  c->rows_in_range = 0;
  c->rows_checked = 0;
  reset_hint(c);
  return MB2_OK;
}

int mb2_restore_state(mb2_ctx* c, const double* energy_host, double best_energy, int64_t best_index,
                      const double* best_x_host, int64_t generation) {
  if (c == nullptr || c->pop[0] == nullptr) return fail(MB2_ESTATE, "no buffers bound");
  if (!energy_host || !best_x_host || generation < 0) return fail(MB2_EINVAL, "bad argument");
  if (c->n_islands > 1) return fail(MB2_EINVAL, "restart files of batched solvers are not supported");
  ENTER_DEVICE(c->device);
  CUDA_TRY(cudaMemcpy(c->en[c->cur], energy_host, sizeof(double) * c->np, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->best_x, best_x_host, sizeof(double) * c->dim, cudaMemcpyHostToDevice));
  DeviceState st = {best_energy, best_index, 0, 0, generation};
  CUDA_TRY(cudaMemcpy(c->st, &st, sizeof(st), cudaMemcpyHostToDevice));
  c->generation = generation;
  c->rows_in_range = 0;  // the caller installed rows this context has not checked
  c->rows_checked = 0;
  reset_hint(c);
  return MB2_OK;
}

int mb2_eval_initial(mb2_ctx* c, void* stream) {
  if (c == nullptr || c->pop[0] == nullptr) return fail(MB2_ESTATE, "no buffers bound");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  c->generation = -1;
  if (int rc = launch_step(c, MODE_GEN0, c->pop[c->cur], c->pop[c->cur ^ 1], c->en[c->cur], c->en[c->cur ^ 1], c->np, s))
    return rc;
  if (int rc = launch_finalize(c, 1, s)) return rc;
  c->cur ^= 1;
  c->generation = 0;
  c->rows_in_range = (c->bounds_mode != BOUNDS_NONE) ? 1 : 0;  // MODE_GEN0 clipped every row
  return MB2_OK;
}

int mb2_step(mb2_ctx* c, void* stream) {
  if (c == nullptr || c->pop[0] == nullptr) return fail(MB2_ESTATE, "no buffers bound");
  if (c->generation < 0) return fail(MB2_ESTATE, "mb2_eval_initial must run before mb2_step");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = launch_step(c, MODE_STEP, c->pop[c->cur], c->pop[c->cur ^ 1], c->en[c->cur], c->en[c->cur ^ 1], c->np, s))
    return rc;
  if (c->last_sparse) {
    if (int rc = launch_commit(c, s)) return rc;
  }
  if (int rc = launch_finalize(c, 0, s)) return rc;  // (reads the best member from the other buffer in both modes)
  if (!c->last_sparse) c->cur ^= 1;
  c->generation += 1;
  c->replay = 0;  // recorded draws are consumed by exactly one generation
  return MB2_OK;
}

int mb2_set_termination(mb2_ctx* c, const mb2_termination* t, const double* history_host, int64_t history_len) {
  if (c == nullptr || t == nullptr) return fail(MB2_EINVAL, "NULL argument");
  if (t->kind < TERM_NONE || t->kind > TERM_LIMITS) return fail(MB2_EINVAL, "unknown termination kind %d", t->kind);
  if (t->lookback < 0 || t->lookback >= kHistMax) return fail(MB2_EINVAL, "lookback must be in [0, %d)", kHistMax);
  if (history_len < 0 || (history_len > 0 && history_host == nullptr)) return fail(MB2_EINVAL, "bad history");
  ENTER_DEVICE(c->device);
  if (c->n_islands > 1 && (history_len != 0 || c->generation >= 0))
    return fail(MB2_EINVAL, "batched solvers take their condition before generation 0 (the device keeps every island's history)");
  if (c->n_islands > 1 && c->island_group > 1)
    return fail(MB2_EINVAL, "islands that share their best member are stepped with mb2_step");
  if (c->run == nullptr) {
    void* p = nullptr;
    if (int rc = dev_alloc(c, sizeof(RunState), &p)) return rc;
    c->run = (RunState*)p;
    if (int rc = host_alloc(c, sizeof(RunState), &p)) return rc;
    c->h_run = (RunState*)p;
  }
  RunState* r = c->h_run;
  memset(r, 0, sizeof(RunState));
  r->kind = t->kind;
  r->lookback = t->lookback;
  r->tol = t->tolerance;
  r->target = t->target;
  r->gtol = t->gtol;
  r->term_generations = t->term_generations;
  r->term_evaluations = t->term_evaluations;
  r->max_generations = t->max_generations;
  r->max_evaluations = t->max_evaluations;
  r->generations = t->generations;
  r->evaluations = t->evaluations;
  r->np_eval = c->np_global;
  r->hist_count = history_len;
  if (history_len > 0) {
    r->hist0 = history_host[0];
    const int64_t first = history_len > kHistMax ? history_len - kHistMax : 0;
    for (int64_t i = first; i < history_len; ++i) r->hist[i % kHistMax] = history_host[i];
  }
  CUDA_TRY(cudaMemcpy(c->run, r, sizeof(RunState), cudaMemcpyHostToDevice));
  if (c->n_islands > 1) {
    // one RunState per island, all starting from the same description (each island counts its own evaluations)
    const int n = c->n_islands;
    if (c->run_isl_cap < n) {
      if (c->run_isl != nullptr) cudaFree(c->run_isl);  // (not from the workspace: kilobytes per island)
      c->run_isl = nullptr;
      CUDA_TRY(cudaMalloc((void**)&c->run_isl, sizeof(RunState) * (size_t)n));
      c->run_isl_cap = n;
      void* p = nullptr;
      if (int rc = dev_alloc(c, sizeof(int) * 2 * (size_t)n, &p)) return rc;
      c->isl_status = (int*)p;
      if (int rc = host_alloc(c, sizeof(int) * 2 * (size_t)n, &p)) return rc;
      c->h_isl_status = (int*)p;
    }
    r->np_eval = c->np / n;
    for (int i = 0; i < n; ++i)
      CUDA_TRY(cudaMemcpy(c->run_isl + i, r, offsetof(RunState, hist), cudaMemcpyHostToDevice));
    c->run_isl_set = 1;
  }
  return MB2_OK;
}

int mb2_run(mb2_ctx* c, int32_t max_steps, double* log_dev, void* stream) {
  if (c == nullptr || c->pop[0] == nullptr) return fail(MB2_ESTATE, "no buffers bound");
  if (c->run == nullptr) return fail(MB2_ESTATE, "mb2_set_termination must run before mb2_run");
  if (max_steps < 1 || log_dev == nullptr) return fail(MB2_EINVAL, "bad argument");
  if (c->replay) return fail(MB2_EINVAL, "batch stepping needs the Philox generator (replayed draws are per generation)");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (c->n_islands > 1) {
    // batch of solvers: every island logs into its own [max_steps, dim+4] block of log_dev and stops on its own
    // condition; generation 0 is the first "step" of the first batch
    if (!c->run_isl_set) return fail(MB2_ESTATE, "mb2_set_termination must run (after mb2_set_islands) before mb2_run");
    const int n = c->n_islands;
    islands_arm_kernel<<<(n + 127) / 128, 128, 0, s>>>(c->run_isl, n, log_dev, (long long)max_steps, c->dim);
    CUDA_TRY(cudaGetLastError());
    ++c->launches;
    c->run_active = 1;
    c->run_enqueued = max_steps;
    int rc = MB2_OK;
    for (int k = 0; k < max_steps && rc == MB2_OK; ++k) {
      const int gen0 = c->generation < 0 ? 1 : 0;
      rc = launch_step(c, gen0 ? MODE_GEN0 : MODE_STEP, c->pop[c->cur], c->pop[c->cur ^ 1], c->en[c->cur], c->en[c->cur ^ 1], c->np, s);
      if (rc == MB2_OK) rc = launch_finalize(c, gen0, s);
      if (rc == MB2_OK) {
        c->cur ^= 1;
        c->generation += 1;
        if (gen0) c->rows_in_range = (c->bounds_mode != BOUNDS_NONE) ? 1 : 0;
      }
    }
    c->run_active = 0;
    return rc;
  }
  if (c->generation < 0) return fail(MB2_ESTATE, "mb2_eval_initial must run before mb2_run");
  // arm the log for this batch (stream ordered: the previous batch has been read back by mb2_run_result)
  struct { long long count, cap; double* log; } arm = {0, max_steps, log_dev};
  CUDA_TRY(cudaMemcpyAsync(&c->run->log_count, &arm, sizeof(arm), cudaMemcpyHostToDevice, s));
  c->run_active = 1;
  c->run_cur0 = c->cur;
  c->run_gen0 = c->generation;
  c->run_enqueued = max_steps;
  int rc = MB2_OK;
  c->run_sparse = 0;
  c->batch_in_place = -1;
  c->batch_in_place = want_in_place(c) ? 1 : 0;  // (the asynchronous hint may move while the batch is enqueued)
  for (int k = 0; k < max_steps && rc == MB2_OK; ++k) {
    rc = launch_step(c, MODE_STEP, c->pop[c->cur], c->pop[c->cur ^ 1], c->en[c->cur], c->en[c->cur ^ 1], c->np, s);
    if (rc == MB2_OK && k == 0) c->run_sparse = c->last_sparse;
    if (rc == MB2_OK && c->last_sparse != c->run_sparse) rc = fail(MB2_ESTATE, "commit mode changed inside a batch");
    if (rc == MB2_OK && c->last_sparse) rc = launch_commit(c, s);
    if (rc == MB2_OK) rc = launch_finalize(c, 0, s);
    if (rc == MB2_OK) {
      if (!c->last_sparse) c->cur ^= 1;
      c->generation += 1;
    }
  }
  c->batch_in_place = -1;
  c->run_active = 0;
  return rc;
}

int mb2_run_result(mb2_ctx* c, int32_t* executed, int32_t* stopped, void* stream) {
  if (c == nullptr || c->run == nullptr || executed == nullptr || stopped == nullptr) return fail(MB2_EINVAL, "bad argument");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (c->xchg_connected) CUDA_TRY(cudaMemcpyAsync(c->h_xerr, c->xp.error, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(c->h_run, c->run, offsetof(RunState, hist0), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(c->h_st, c->st, sizeof(DeviceState), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  const int done = (int)c->h_run->log_count;
  *executed = done;
  *stopped = (c->h_run->stop == 1) ? 1 : 0;
  if (done > 0) note_accepted(c, c->h_st->n_acc);
  // generations after the stop were no-ops: put the double buffer and the counter where the device is
  // (a batch committed in place never swapped)
  c->cur = c->run_cur0 ^ (c->run_sparse ? 0 : (done & 1));
  c->generation = c->run_gen0 + done;
  // a sharded batch whose best exchange timed out: the shards' bests (and stop decisions) may differ across ranks
  if ((c->xchg_connected && *c->h_xerr != 0) || c->h_run->stop == 3)
    return fail(MB2_ECUDA, "best exchange: a peer did not deliver its record within %.1f s (batch stopped after %d generations)",
                c->xp.timeout_ns * 1e-9, done);
  if (c->h_run->stop == 2) {  // log full without a termination: clear the flag for the next batch
    int zero = 0;
    CUDA_TRY(cudaMemcpy(&c->run->stop, &zero, sizeof(int), cudaMemcpyHostToDevice));
  }
  return MB2_OK;
}

int mb2_run_result_islands(mb2_ctx* c, int32_t* executed, int32_t* stopped, void* stream) {
  if (c == nullptr || executed == nullptr || stopped == nullptr) return fail(MB2_EINVAL, "bad argument");
  if (c->n_islands < 2 || !c->run_isl_set) return fail(MB2_ESTATE, "no batch of solvers with a device-side condition");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int n = c->n_islands;
  islands_status_kernel<<<(n + 127) / 128, 128, 0, s>>>(c->run_isl, n, c->isl_status);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  CUDA_TRY(cudaMemcpyAsync(c->h_isl_status, c->isl_status, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (int i = 0; i < n; ++i) {
    executed[i] = c->h_isl_status[2 * i];
    stopped[i] = (c->h_isl_status[2 * i + 1] == 1) ? 1 : 0;
  }
  return MB2_OK;
}

int mb2_pack_best(mb2_ctx* c, double* record, void* stream) {
  if (c == nullptr || record == nullptr) return fail(MB2_EINVAL, "NULL argument");
  ENTER_DEVICE(c->device);
  pack_best_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c->st, c->best_x, c->dim, record);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  return MB2_OK;
}

int mb2_merge_best(mb2_ctx* c, const double* records, int32_t world, void* stream) {
  if (c == nullptr || records == nullptr || world < 1) return fail(MB2_EINVAL, "bad argument");
  ENTER_DEVICE(c->device);
  merge_best_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(records, world, c->dim, c->best_x, c->st);
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  c->counts_global = world > 1 ? 1 : 0;
  return MB2_OK;
}

int mb2_set_commit_mode(mb2_ctx* c, int mode) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (mode < MB2_COMMIT_AUTO || mode > MB2_COMMIT_IN_PLACE) return fail(MB2_EINVAL, "unknown commit mode %d", mode);
  c->commit_mode = mode;
  return MB2_OK;
}

int mb2_last_commit_in_place(mb2_ctx* c) { return (c != nullptr && c->last_sparse) ? 1 : 0; }

int mb2_set_islands(mb2_ctx* c, int32_t n_islands) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (n_islands < 1 || c->np % n_islands != 0)
    return fail(MB2_EINVAL, "%lld rows do not split into %d equal populations", (long long)c->np, n_islands);
  if (c->xchg_connected) return fail(MB2_ESTATE, "batched solvers and sharded populations do not combine");
  if (c->generation >= 0) return fail(MB2_ESTATE, "mb2_set_islands must come before mb2_eval_initial");
  const int64_t inp = c->np / n_islands;
  if (n_islands > 1 && inp < c->sinfo.k + 1)
    return fail(MB2_EINVAL, "each population needs at least %d members for this strategy", c->sinfo.k + 1);
  ENTER_DEVICE(c->device);
  if (n_islands > 1 && (c->st_isl == nullptr || n_islands > c->n_islands)) {
    void* p = nullptr;
    if (int rc = dev_alloc(c, sizeof(DeviceState) * n_islands, &p)) return rc;
    c->st_isl = (DeviceState*)p;
    if (int rc = dev_alloc(c, sizeof(double) * (size_t)n_islands * c->dim, &p)) return rc;
    c->best_x_isl = (double*)p;
    if (int rc = host_alloc(c, sizeof(DeviceState) * n_islands, &p)) return rc;
    c->h_st_isl = (DeviceState*)p;
    if (int rc = host_alloc(c, sizeof(double) * (size_t)n_islands * c->dim, &p)) return rc;
    c->h_best_isl = (double*)p;
  }
  if (n_islands > 1) {
    std::vector<DeviceState> init((size_t)n_islands, DeviceState{INFINITY, -1, 0, 0, -1});
    CUDA_TRY(cudaMemcpy(c->st_isl, init.data(), sizeof(DeviceState) * n_islands, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(c->best_x_isl, 0, sizeof(double) * (size_t)n_islands * c->dim));
  }
  c->n_islands = n_islands;
  c->island_group = 1;
  c->run_isl_set = 0;
  return MB2_OK;
}

int mb2_set_island_groups(mb2_ctx* c, int32_t group_size) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (group_size < 1 || c->n_islands % group_size != 0)
    return fail(MB2_EINVAL, "%d islands do not split into groups of %d", c->n_islands, group_size);
  if (c->generation >= 0) return fail(MB2_ESTATE, "mb2_set_island_groups must come before mb2_eval_initial");
  c->island_group = group_size;
  return MB2_OK;
}

int mb2_read_islands(mb2_ctx* c, mb2_result* out, double* x_host, void* stream) {
  if (c == nullptr || out == nullptr) return fail(MB2_EINVAL, "NULL argument");
  if (c->n_islands < 2) return fail(MB2_ESTATE, "mb2_set_islands first");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)c->n_islands;
  CUDA_TRY(cudaMemcpyAsync(c->h_st_isl, c->st_isl, sizeof(DeviceState) * n, cudaMemcpyDeviceToHost, s));
  if (x_host) CUDA_TRY(cudaMemcpyAsync(c->h_best_isl, c->best_x_isl, sizeof(double) * n * c->dim, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  for (size_t i = 0; i < n; ++i) {
    out[i].best_energy = c->h_st_isl[i].best_e;
    out[i].best_index = c->h_st_isl[i].best_idx;
    out[i].n_inf = c->h_st_isl[i].n_inf;
    out[i].n_accepted = c->h_st_isl[i].n_acc;
    out[i].generation = c->h_st_isl[i].generation;
  }
  if (x_host) memcpy(x_host, c->h_best_isl, sizeof(double) * n * c->dim);
  return MB2_OK;
}

int mb2_exchange_open(mb2_ctx* c, int32_t world, int32_t rank, void* handle_out, int32_t handle_bytes) {
  if (c == nullptr || handle_out == nullptr) return fail(MB2_EINVAL, "NULL argument");
  if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
    return fail(MB2_EINVAL, "world %d / rank %d out of range (2..%d ranks)", world, rank, kMaxWorld);
  if (handle_bytes < (int32_t)sizeof(cudaIpcMemHandle_t)) return fail(MB2_EINVAL, "handle buffer too small");
  if (c->mailbox != nullptr) return fail(MB2_ESTATE, "exchange already opened");
  if (c->n_islands > 1) return fail(MB2_ESTATE, "batched solvers and sharded populations do not combine");
  ENTER_DEVICE(c->device);
  const int stride = (c->dim + 4 + 1) & ~1;  // 16-byte slots
  memset(handle_out, 0, (size_t)handle_bytes);
  {
    void* p = nullptr;
    if (int rc = host_alloc(c, sizeof(int), &p)) return rc;
    c->h_xerr = (int*)p;
    *c->h_xerr = 0;
  }
  if (!c->xchg_no_adopt) {
    // a parked channel of an earlier context of this process?
    std::lock_guard<std::mutex> lock(g_park_mutex);
    for (size_t k = 0; k < g_parked.size(); ++k) {
      const ParkedChannel& pc = g_parked[k];
      if (pc.device == c->device && pc.xp.world == world && pc.xp.rank == rank && pc.xp.stride == stride) {
        c->xp = pc.xp;  // the epoch counter carries on
        c->mailbox = pc.mailbox;
        c->mailbox_bytes = pc.bytes;
        memcpy(c->peer_maps, pc.peer_maps, sizeof(c->peer_maps));
        c->xchg_connected = 1;
        g_parked.erase(g_parked.begin() + (long)k);
        return MB2_OK;
      }
    }
  }
  const size_t bytes = sizeof(double) * 2 * (size_t)world * stride + sizeof(unsigned long long) * 2 * (size_t)world + sizeof(int) * 4;
  // its own allocation: an IPC handle names a whole cudaMalloc block
  CUDA_TRY(cudaMalloc((void**)&c->mailbox, bytes));
  CUDA_TRY(cudaMemset(c->mailbox, 0, bytes));  // flags 0: epochs start at 1
  c->mailbox_bytes = bytes;
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c->mailbox));
  memcpy(handle_out, &h, sizeof(h));
  c->xp = {};
  c->xp.world = world;
  c->xp.rank = rank;
  c->xp.stride = stride;
  c->xp.epoch = 0;
  {
    const int ms = env_int("MB2_EXCHANGE_TIMEOUT_MS", 5000);  // tests shorten it
    c->xp.timeout_ns = (unsigned long long)(ms > 0 ? ms : 5000) * 1000000ull;
  }
  c->xp.box[rank] = c->mailbox;
  c->xp.error = reinterpret_cast<int*>(reinterpret_cast<char*>(c->mailbox) + bytes - sizeof(int) * 4);
  CUDA_TRY(cudaDeviceSynchronize());
  return MB2_OK;
}

int mb2_exchange_connected(mb2_ctx* c) { return (c != nullptr && c->xchg_connected) ? 1 : 0; }

int mb2_exchange_reset(mb2_ctx* c) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  ENTER_DEVICE(c->device);
  if (c->mailbox != nullptr) {
    CUDA_TRY(cudaDeviceSynchronize());
    for (int r = 0; r < kMaxWorld; ++r) {
      if (c->peer_maps[r] != nullptr) cudaIpcCloseMemHandle(c->peer_maps[r]);
      c->peer_maps[r] = nullptr;
    }
    cudaFree(c->mailbox);
    c->mailbox = nullptr;
  }
  c->xp = {};
  c->xchg_connected = 0;
  c->xchg_no_adopt = 1;
  return MB2_OK;
}

int mb2_exchange_connect(mb2_ctx* c, const void* handles, int32_t handle_bytes) {
  if (c == nullptr || handles == nullptr) return fail(MB2_EINVAL, "NULL argument");
  if (c->mailbox == nullptr) return fail(MB2_ESTATE, "mb2_exchange_open first");
  if (c->xchg_connected) return MB2_OK;
  if (handle_bytes < (int32_t)sizeof(cudaIpcMemHandle_t)) return fail(MB2_EINVAL, "handle stride too small");
  ENTER_DEVICE(c->device);
  for (int r = 0; r < c->xp.world; ++r) {
    if (r == c->xp.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)r * handle_bytes, sizeof(h));
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_maps[r] = p;
    c->xp.box[r] = static_cast<double*>(p);
  }
  c->xchg_connected = 1;
  return MB2_OK;
}

int mb2_exchange_best(mb2_ctx* c, void* stream) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (!c->xchg_connected) return fail(MB2_ESTATE, "mb2_exchange_connect first");
  ENTER_DEVICE(c->device);
  exchange_best_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c->dim, c->best_x, c->st, next_exchange(c));
  CUDA_TRY(cudaGetLastError());
  ++c->launches;
  return MB2_OK;
}

int mb2_read_result(mb2_ctx* c, mb2_result* out, double* x_host, void* stream) {
  if (c == nullptr || out == nullptr) return fail(MB2_EINVAL, "NULL argument");
  ENTER_DEVICE(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (c->xchg_connected) CUDA_TRY(cudaMemcpyAsync(c->h_xerr, c->xp.error, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(c->h_st, c->st, sizeof(DeviceState), cudaMemcpyDeviceToHost, s));
  if (x_host) CUDA_TRY(cudaMemcpyAsync(c->h_best_x, c->best_x, sizeof(double) * c->dim, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (c->xchg_connected && *c->h_xerr != 0)
    return fail(MB2_ECUDA, "best exchange: a peer did not deliver its record within %.1f s", c->xp.timeout_ns * 1e-9);
  out->best_energy = c->h_st->best_e;
  out->best_index = c->h_st->best_idx;
  out->n_inf = c->h_st->n_inf;
  out->n_accepted = c->h_st->n_acc;
  out->generation = c->h_st->generation;
  if (c->h_st->generation > 0) note_accepted(c, c->h_st->n_acc);  // (generation 0 "accepts" every row)
  if (x_host) memcpy(x_host, c->h_best_x, sizeof(double) * c->dim);
  return MB2_OK;
}

int mb2_eval_cost(mb2_ctx* c, const double* x, int64_t n, double* out, void* stream) {
  if (c == nullptr || x == nullptr || out == nullptr || n < 1) return fail(MB2_EINVAL, "bad argument");
  ENTER_DEVICE(c->device);
  const int64_t saved_gen = c->generation;
  const int saved_grid = c->grid;
  double* cap_t = c->cap_trial;
  c->cap_trial = nullptr;
  int rc = launch_step(c, MODE_EVAL, x, nullptr, nullptr, out, n, (cudaStream_t)stream);
  c->cap_trial = cap_t;
  c->generation = saved_gen;
  c->grid = saved_grid;
  return rc;
}

int mb2_launch_info(mb2_ctx* c, int32_t* grid, int32_t* block, int32_t* smem) {
  if (c == nullptr) return fail(MB2_EINVAL, "ctx is NULL");
  if (grid) *grid = c->grid;
  if (block) *block = c->wpc * 32;
  if (smem) *smem = c->smem;
  return MB2_OK;
}

int64_t mb2_launch_count(mb2_ctx* c) { return c ? c->launches : 0; }

int mb2_probe_pool_geometry(int32_t dim, int32_t* out) {
  if (out == nullptr) return fail(MB2_EINVAL, "out is NULL");
  int p = 0, q = 0, g = 0, gw = 0, np_ = 0, smem = 0;
  for (int i = 0; i < 6; ++i) out[i] = 0;
  if (!pool_geometry((int)dim, &p, &q, &g, &gw, &np_, &smem)) return MB2_EINVAL;  // (no message: "K2a runs these rows" is an answer)
  out[0] = p;
  out[1] = q;
  out[2] = g;
  out[3] = gw;
  out[4] = np_;
  out[5] = smem;
  return MB2_OK;
}

int mb2_kernel_family(mb2_ctx* c) { return c ? c->family : MB2_KERNEL_NONE; }

}  // extern "C"
