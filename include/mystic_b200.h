/*
 * mystic_b200.h -- C ABI of the B200-native differential-evolution generation engine.
 *
 * Drop-in boundary for the hot path of mystic's DifferentialEvolutionSolver2
 * (reference: mystic/differential_evolution.py:531-625, `_Step`).  The reference has no
 * FFI of its own (it is pure Python); this header is what a binding for that path binds.
 * Each entry point cites the reference code it replaces (paths relative to the mystic
 * repository root).  The Python host mirror (mystic_b200/solver.py) calls exactly these
 * symbols through ctypes; see INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MB2_E* code on failure; the message is
 *     available from mb2_last_error() (thread local).  No exceptions cross the boundary.
 *   - `*_dev` pointers are device pointers owned by the caller (torch tensors in the Python
 *     host); `*_host` pointers are host pointers and are copied before the call returns.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls that
 *     take a stream are asynchronous on it unless stated otherwise.
 *   - a context is bound to one device and used from one host thread at a time.
 *   - all floating-point data is IEEE binary64; the population is [np_local, dim] row-major
 *     (the reference's list-of-lists `population[NP][D]`, mystic/abstract_solver.py:118-120).
 */
#ifndef MYSTIC_B200_H
#define MYSTIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: draw protocol 2 of the Philox mode (csrc/philox.cuh): one geometric draw for the exponential crossover
 * length, 16-bit fields for the binomial decisions.  Signatures are those of version 1. */
#define MB2_ABI_VERSION 2

/* error codes */
#define MB2_OK 0
#define MB2_EINVAL (-1)   /* bad argument / unsupported combination */
#define MB2_ECUDA (-2)    /* a CUDA runtime call failed */
#define MB2_ESTATE (-3)   /* call order violated (e.g. step before bind_buffers) */

/* mystic/strategy.py -- ids in file order.  Only Best1Bin is binomial in the reference; the
 * other *Bin names run the exponential loop (strategy.py:203-311). */
enum mb2_strategy {
  MB2_BEST1EXP = 0,       /* strategy.py:34-59   */
  MB2_BEST1BIN = 1,       /* strategy.py:61-88   */
  MB2_RAND1EXP = 2,       /* strategy.py:90-115  */
  MB2_RANDTOBEST1EXP = 3, /* strategy.py:119-145 */
  MB2_BEST2EXP = 4,       /* strategy.py:147-173 */
  MB2_RAND2EXP = 5,       /* strategy.py:175-201 */
  MB2_RAND1BIN = 6,       /* strategy.py:203-227 */
  MB2_RANDTOBEST1BIN = 7, /* strategy.py:229-255 */
  MB2_BEST2BIN = 8,       /* strategy.py:257-283 */
  MB2_RAND2BIN = 9        /* strategy.py:285-311 */
};

/* mystic/models -- batched device cost functions */
enum mb2_cost {
  MB2_COST_ROSEN = 0,      /* models/dejong.py:89-101 */
  MB2_COST_CORANA = 1,     /* models/storn.py:57-96   */
  MB2_COST_GRIEWANGK = 2,  /* models/storn.py:120-139 */
  MB2_COST_ZIMMERMANN = 3, /* models/storn.py:161-191 */
  MB2_COST_CHEBYSHEV = 4,  /* models/_model_helper.py:16-33, Horner (math/poly.py:37-39), bit-faithful px */
  MB2_COST_CHEBYSHEV_MMA = 5, /* same cost, [NP x D]*[D x M] contraction on FP64 tensor cores (DMMA) */
  /* data-fit residual costs of forward_model.py:204-260 (CostFactory.getCostFunction with the default
   * L2 metric): sum(((F(params)(x_i) - obs_i)/sigma)^2) over M evaluation points */
  MB2_COST_POLYFIT = 6,     /* F = models/poly.py:36-54 (numpy.poly1d), dim = number of coefficients <= 32 */
  MB2_COST_POLYFIT_MMA = 7, /* same cost, the coefficient x point contraction on DMMA (dim <= 17) */
  MB2_COST_LORENTZIAN = 8,  /* F = models/lorentzian.py:32-37, dim = 7 */
  /* separable test functions (any dim) */
  MB2_COST_SPHERE = 9,      /* models/dejong.py:54-66   */
  MB2_COST_RASTRIGIN = 10,  /* models/pohlheim.py:123-132 */
  MB2_COST_SCHWEFEL = 11,   /* models/pohlheim.py:60-71 */
  MB2_COST_ACKLEY = 12,     /* models/pohlheim.py:186-204 */
  MB2_COST_STEP = 13,       /* models/dejong.py:166-192 */
  MB2_COST_MICHAL = 14,     /* models/pohlheim.py:226-240 */
  MB2_COST_POWERS = 15,     /* models/pohlheim.py:153-163 */
  /* closed-form functions of two variables (dim must be 2) ... */
  MB2_COST_BRANINS = 16,    /* models/pohlheim.py:261-282 */
  MB2_COST_EASOM = 17,      /* models/pohlheim.py:303-313 */
  MB2_COST_GOLDSTEIN = 18,  /* models/pohlheim.py:336-355 */
  MB2_COST_PEAKS = 19,      /* models/nag.py:48-59 */
  MB2_COST_VENKAT91 = 20,   /* models/venkataraman.py:38-52 */
  MB2_COST_FOSC3D = 21,     /* models/wolfram.py:44-64 */
  MB2_COST_NMIN51 = 22,     /* models/wolfram.py:87-103 */
  MB2_COST_SHEKEL = 23,     /* models/dejong.py:252-278 */
  /* ... and of any number of variables, evaluated coordinate by coordinate in the reference's order */
  MB2_COST_ELLIPSOID = 24,  /* models/pohlheim.py:92-102 */
  MB2_COST_PAVIANI = 25,    /* models/schittkowski.py:46-71 */
  /* data-fit residual cost (like 6 and 8) with F = models/br8.py:27-32, dual exponential decay, dim = 5 */
  MB2_COST_DECAY = 26
};

/* SetStrictRanges modes (abstract_solver.py:334-448, tools.py:404-430) */
enum mb2_bounds {
  MB2_BOUNDS_NONE = 0,   /* no strict range */
  MB2_BOUNDS_REJECT = 1, /* default SetStrictRanges: out-of-range trial -> energy inf, cost not evaluated */
  MB2_BOUNDS_CLAMP = 2   /* tight=True: x[i]=max(lo,x[i]); x[i]=min(hi,x[i]) before the cost (symbolic.py:1139-1144) */
};

/* mystic/penalty.py decorators whose condition is linear, G.x - h */
enum mb2_penalty {
  MB2_PEN_QUADRATIC_EQUALITY = 0,   /* penalty.py:41-98   : k*pf^2        */
  MB2_PEN_LINEAR_EQUALITY = 1,      /* penalty.py:100-157 : k*|pf|        */
  MB2_PEN_QUADRATIC_INEQUALITY = 2, /* penalty.py:341-398 : 2k*max(0,pf)^2 */
  MB2_PEN_LINEAR_INEQUALITY = 3,    /* penalty.py:400-457 : 2k*|max(0,pf)| */
  MB2_PEN_UNIFORM_EQUALITY = 4,     /* penalty.py:159-216 : k if pf else 0   (k may be inf)          */
  MB2_PEN_UNIFORM_INEQUALITY = 5,   /* penalty.py:218-274 : k if pf > 0 else 0                       */
  MB2_PEN_BARRIER_INEQUALITY = 6,   /* penalty.py:276-339 : inf if pf > 0 else -.5/k*log(-pf)        */
  MB2_PEN_LAGRANGE_INEQUALITY = 7,  /* penalty.py:459-530 : k*mpf^2 + beta*mpf, mpf = max(-beta/(2k), pf) */
  MB2_PEN_LAGRANGE_EQUALITY = 8     /* penalty.py:532-601 : k*pf^2 + lam*pf                          */
};

typedef struct mb2_ctx mb2_ctx;

/* result of one generation, as the host part of `_Step` needs it
 * (differential_evolution.py:596-617: fcalls bookkeeping, bestEnergy, stepmon) */
typedef struct mb2_result {
  double best_energy;   /* solver.bestEnergy after the generation                               */
  int64_t best_index;   /* GLOBAL index of the member that last became best, -1 if none yet      */
  int64_t n_inf;        /* trial energies equal to +-inf in this generation (:601 `isinf().sum()`) */
  int64_t n_accepted;   /* trials that replaced their parent (:604)                               */
  int64_t generation;   /* 0 = initial population evaluated                                      */
} mb2_result;

int mb2_abi_version(void);
const char* mb2_last_error(void);

/* One context per device shard.  Replaces the per-solver state of AbstractSolver.__init__
 * (abstract_solver.py:92-150): nDim, nPop, popEnergy=[inf]*NP, bestEnergy, bestSolution.
 * np_local rows live on this device; they are rows [global_offset, global_offset+np_local)
 * of a population of np_global rows (single GPU: np_local == np_global, offset 0). */
int mb2_create(int device, int64_t np_local, int32_t dim, int64_t np_global, int64_t global_offset,
               mb2_ctx** out);
int mb2_destroy(mb2_ctx* ctx);
/* Same, but the context's small device buffers (state, best member, bounds, per-CTA partials, cost
 * tables) and its pinned host staging are carved out of caller-owned workspaces (e.g. tensors from
 * a caching allocator) instead of cudaMalloc / cudaMallocHost, which cost milliseconds each next to
 * gigabytes of live allocations.  mb2_workspace_bytes gives sufficient sizes for the common costs;
 * anything that does not fit falls back to cudaMalloc.  The workspaces must outlive the context. */
int64_t mb2_workspace_bytes(int32_t dim, int64_t* host_bytes);
int mb2_create_ws(int device, int64_t np_local, int32_t dim, int64_t np_global, int64_t global_offset,
                  void* dev_workspace, int64_t dev_bytes, void* host_workspace, int64_t host_bytes, mb2_ctx** out);

/* Double-buffered population and energies (the reference keeps `population` and
 * `trialSolution`, abstract_map_solver.py:113-114).  pop_a is the current generation. */
int mb2_bind_buffers(mb2_ctx* ctx, double* pop_a_dev, double* pop_b_dev, double* energy_a_dev,
                     double* energy_b_dev);
/* device pointers of the CURRENT generation (they swap every step) */
int mb2_current(mb2_ctx* ctx, double** pop_dev, double** energy_dev);

/* differential_evolution.py:460-462, 676-693: strategy, ScalingFactor, CrossProbability */
int mb2_set_strategy(mb2_ctx* ctx, int strategy, double scale, double cross);
/* abstract_solver.py:334-405 SetStrictRanges; lo/hi have `dim` entries (ignored for NONE) */
int mb2_set_bounds(mb2_ctx* ctx, int mode, const double* lo_host, const double* hi_host);
/* cost(x, *ExtraArgs): params for CHEBYSHEV* = {M, order, target coeffs[order+1]};
 * for POLYFIT* / LORENTZIAN / DECAY = {M, sigma, evalpts[M], observations[M]}, optionally followed by
 * sigma[M] (a sigma per point: forward_model.py:255 divides elementwise; the scalar is then ignored) */
int mb2_set_cost(mb2_ctx* ctx, int cost, const double* params_host, int32_t nparams);
/* stacked penalty decorators, term 0 innermost (penalty.py:80-98 `dec`); G is [nterms, dim],
 * kscale[i] = k*h**n (the decorator's current multiplier before the inequality factor 2) */
int mb2_set_penalty(mb2_ctx* ctx, int32_t nterms, const int32_t* ptype_host, const double* G_host,
                    const double* h_host, const double* kscale_host);
/* The same with per-term multipliers: the lagrange kinds accumulate beta / lambda over the values `store(x)`
 * recorded at earlier penalty iterations (penalty.py:511-515, 579-582 -- host state, a handful of scalars);
 * mult_host[i] is that sum for term i (ignored by the other kinds; NULL = all zero).  kscale[i] is the
 * multiplier k*h**n after the same loop. */
int mb2_set_penalty_ex(mb2_ctx* ctx, int32_t nterms, const int32_t* ptype_host, const double* G_host,
                       const double* h_host, const double* kscale_host, const double* mult_host);

/* Random draws.  Philox4x32-10 counter RNG keyed by `seed`, counter = (block, global candidate,
 * generation, stream) -- replaces random.sample / random.randrange / random.random of
 * strategy.py:27,42,49 in production mode. */
int mb2_set_rng_philox(mb2_ctx* ctx, uint64_t seed);
/* Validation mode: consume the reference's own recorded draws for the NEXT mb2_step call.
 * donors [np_local,5] int32 (unused = -1), nstart [np_local] int32, uniforms [np_local, dim+1]
 * float64 with NaN for draws the reference did not make. */
int mb2_set_rng_replay(mb2_ctx* ctx, const int32_t* donors_dev, const int32_t* nstart_dev,
                       const double* uniforms_dev);

/* abstract_solver.py:508-535 SetRandomInitialPoints, generated on the device (Philox stream 2)
 * into the current population buffer; resets energies to +inf. */
int mb2_init_population(mb2_ctx* ctx, uint64_t seed, const double* lo_host, const double* hi_host,
                        void* stream);
/* host population [np_local, dim] -> current buffer (validation / small problems); resets
 * energies to +inf.  Synchronous w.r.t. the host buffer. */
int mb2_upload_population(mb2_ctx* ctx, const double* pop_host, void* stream);
/* optional capture of every trial vector / trial energy of the next steps (NULL disables) */
int mb2_set_capture(mb2_ctx* ctx, double* trial_dev, double* trial_energy_dev);

/* Resume a saved solver (abstract_solver.py:975-1027 SaveSolver / solvers.py:90-125 LoadSolver):
 * after mb2_upload_population of the saved population, install its energies, the best member
 * and the generation counter without re-evaluating anything.  Synchronous. */
int mb2_restore_state(mb2_ctx* ctx, const double* energy_host, double best_energy, int64_t best_index,
                      const double* best_x_host, int64_t generation);

/* generation 0: differential_evolution.py:516-520 (clip into the strict range), :559-567,
 * :575-582 with strategy=None, :589, :603-613 */
int mb2_eval_initial(mb2_ctx* ctx, void* stream);
/* one generation g>=1: trial generation (:574-582 + strategy.py), bounds, cost (:589),
 * penalty, selection and generation best (:603-613) in ONE fused kernel + a 1-CTA argmin */
int mb2_step(mb2_ctx* ctx, void* stream);

/* Device-resident multi-generation stepping.  `Solve` in the reference returns to the host after
 * every generation to evaluate the termination condition (abstract_solver.py:1134-1148); the
 * conditions that only read the energy history and the counters are evaluated here by the argmin
 * epilogue instead, so a batch of generations runs without a host round trip and stops exactly
 * where the reference would.  kinds: 0 none, 1 VTR (termination.py:181-196), 2 ChangeOverGeneration
 * (:198-217), 3 NormalizedChangeOverGeneration (:219-240), 4 VTRChangeOverGeneration (:320-342),
 * 5 EvaluationLimits (:386-408); max_* are the solver's own limits (abstract_solver.py:703-711),
 * negative = none. */
typedef struct mb2_termination {
  int32_t kind;
  int32_t lookback;          /* the condition's `generations` argument */
  double tolerance;          /* VTR tolerance / COG, NCOG tolerance / VTRCOG ftol */
  double target;
  double gtol;               /* VTRCOG only */
  int64_t term_generations;  /* EvaluationLimits condition */
  int64_t term_evaluations;
  int64_t max_generations;   /* solver._maxiter */
  int64_t max_evaluations;   /* solver._maxfun */
  int64_t generations;       /* solver.generations so far */
  int64_t evaluations;       /* solver.evaluations so far */
} mb2_termination;
/* install the condition and seed the device-side energy history (solver.energy_history) */
int mb2_set_termination(mb2_ctx* ctx, const mb2_termination* t, const double* history_host, int64_t history_len);
/* enqueue up to max_steps generations; each appends [best_energy, best_index, n_inf, n_accepted,
 * bestSolution[dim]] to log_dev [max_steps, dim+4]; generations after the stop are no-ops */
int mb2_run(mb2_ctx* ctx, int32_t max_steps, double* log_dev, void* stream);
/* synchronise; number of generations actually executed and whether a condition stopped the batch */
int mb2_run_result(mb2_ctx* ctx, int32_t* executed, int32_t* stopped, void* stream);
/* The same loop for a batch of solvers (mb2_set_islands; the ensemble solvers of abstract_ensemble_solver.py:737-826
 * run N nested solvers, each to ITS OWN termination): mb2_set_termination -- after mb2_set_islands, before generation 0,
 * history_len 0 -- installs the condition for every island; mb2_run then takes log_dev [n_islands, max_steps, dim+4],
 * starts with generation 0 when it has not run yet, and every island logs and stops on its own (a stopped island keeps
 * its rows, energies and best member while the others go on).  mb2_run_result_islands synchronises and returns, per
 * island, the generations logged in the batch and whether the island has met its condition. */
int mb2_run_result_islands(mb2_ctx* ctx, int32_t* executed, int32_t* stopped, void* stream);

/* multi-GPU best exchange (Best and RandToBest strategies read the global best,
 * strategy.py:52,83,137,164,247,274): pack this shard's [best_energy, best_index(global),
 * n_inf, n_accepted, bestSolution[dim]] (dim+4 doubles) for an all-gather, then merge the
 * gathered [world, dim+4] records: lowest energy wins, ties to the lowest global index
 * (differential_evolution.py:611 strict '<' in ascending candidate order). */
int mb2_pack_best(mb2_ctx* ctx, double* record_dev, void* stream);
int mb2_merge_best(mb2_ctx* ctx, const double* records_dev, int32_t world, void* stream);

/* The same exchange WITHOUT a collective library call: fused into the argmin epilogue of
 * mb2_eval_initial / mb2_step as stores into peer memory over NVLink.
 *   mb2_exchange_open     allocates this rank's mailbox and returns its CUDA IPC handle
 *                         (handle_bytes >= 64);
 *   mb2_exchange_connect  takes the handles of all ranks ([world][handle_bytes], exchanged by the
 *                         caller, e.g. torch.distributed.all_gather_object) and maps the peers'
 *                         mailboxes; from then on every mb2_eval_initial / mb2_step of this context
 *                         ends with the exchange, and every rank must make the same sequence of calls;
 *   mb2_exchange_connected  1 when mb2_exchange_open adopted the mailbox and peer mappings an
 *                         earlier, destroyed context of this process left behind (same device, world,
 *                         rank and dim): the handle round trip is not needed IF EVERY rank adopted
 *                         (the caller checks that with one small reduction); otherwise
 *   mb2_exchange_reset    drops what was adopted, and mb2_exchange_open then starts afresh;
 *   mb2_exchange_best     runs the exchange alone (after mb2_restore_state).
 * A peer that does not deliver within 5 s makes the next mb2_read_result fail. One node, <= 16 ranks. */
int mb2_exchange_open(mb2_ctx* ctx, int32_t world, int32_t rank, void* handle_out, int32_t handle_bytes);
int mb2_exchange_connect(mb2_ctx* ctx, const void* handles, int32_t handle_bytes);
int mb2_exchange_connected(mb2_ctx* ctx);
int mb2_exchange_reset(mb2_ctx* ctx);
int mb2_exchange_best(mb2_ctx* ctx, void* stream);

/* Batch of independent solvers in one launch (the ensemble path: mystic/abstract_ensemble_solver.py
 * launches N nested solvers through `map`; here N small DE2 populations advance together).
 * The `np_local` rows of the context are n_islands populations of np_local / n_islands members laid
 * side by side; every island draws its donors among its own members (its rows key the Philox
 * counters), follows its own best member and reports its own result: island i behaves exactly like
 * shard i of a sharded population without the best exchange.  Rows of at most ~3200 dimensions (shared-memory bound);
 * mb2_eval_initial / mb2_step advance all islands, mb2_read_islands returns n_islands results and
 * best members ([n_islands, dim], may be NULL).  Call before mb2_eval_initial. */
int mb2_set_islands(mb2_ctx* ctx, int32_t n_islands);
int mb2_read_islands(mb2_ctx* ctx, mb2_result* out, double* x_host, void* stream);
/* Islands that share their best member within groups of `group_size` consecutive islands (after mb2_set_islands,
 * before mb2_eval_initial): every generation ends with the group's winner -- lowest energy, ties to the lowest
 * global index, the rule of mb2_merge_best -- installed in each island of the group.  A group is then exactly a
 * population sharded over group_size ranks (shard-local donors drawn per strategy.py:27 from the shard, global best
 * for the Best / RandToBest strategies), on one GPU; 1 = independent islands. */
int mb2_set_island_groups(mb2_ctx* ctx, int32_t group_size);

/* synchronises `stream`, then returns the generation result and bestSolution (dim doubles;
 * x_host may be NULL) */
int mb2_read_result(mb2_ctx* ctx, mb2_result* out, double* x_host, void* stream);

/* the decorated cost of differential_evolution.py:515-525 for n arbitrary rows:
 * (inf if out of the strict range else cost(x)) + penalty(x).  x_dev [n, dim], out_dev [n]. */
int mb2_eval_cost(mb2_ctx* ctx, const double* x_dev, int64_t n, double* out_dev, void* stream);

/* self test of the table-driven divide used by the griewangk cost (x / sqrt(i+1) without a divide
 * instruction): compares n random operands in [-span, span] bit for bit with the IEEE divide and
 * returns the number of mismatches (expected 0).  Requires mb2_set_cost(GRIEWANGK) first. */
int mb2_selftest_division(mb2_ctx* ctx, int64_t n, uint64_t seed, double span, int64_t* mismatches);

/* Self-test of the griewangk cosine (cos_bounded in csrc/de_device.cuh, which replaces math.cos of
 * mystic/models/storn.py:139 on the device): max |cos_bounded(a) - cos(a)| over n arguments drawn
 * uniformly from [-span, span], span <= 1e5.  Synchronises. */
int mb2_selftest_cosine(mb2_ctx* ctx, int64_t n, uint64_t seed, double span, double* max_abs_err);

/* FP64 peaks of `device`, measured now: a DFMA chain per thread (vector pipe) and mma.sync.m8n8k4.f64
 * (SASS DMMA.8x8x4), in TFLOP/s.  bench.py divides the Chebyshev / polynomial-fit contraction's
 * achieved rate (MB2_COST_CHEBYSHEV_MMA, the reference's _model_helper.py:16-33 evaluated for a whole
 * population) by the DMMA figure instead of a constant.  Synchronises `stream`; no context needed. */
int mb2_probe_fp64_peak(int device, int32_t iters, double* dfma_tflops, double* dmma_tflops, void* stream);

/* Geometry the host would give the pooled producer / consumer kernel K2p (de_tma_pool.cuh: Best1Bin generations of
 * `dim`-dimensional rows committed in place -- the reference's population + trialSolution layout,
 * mystic/differential_evolution.py:585-609) with the MB2_POOL_* environment overrides applied:
 * out[0..5] = {parent + second-donor pairs P, first-donor buffers Q, consumer groups G, warps per group, producer warps,
 * dynamic shared memory in bytes}.  MB2_EINVAL (out zeroed) when K2p does not take such rows.  A pure host function:
 * no device, no context -- the CPU tests check the invariants the kernel's mbarrier protocol relies on (Q >= G,
 * 2P + Q row buffers + bestSolution within a CTA's shared memory, the builds' thread bounds). */
int mb2_probe_pool_geometry(int32_t dim, int32_t* out);

/* launch geometry of the fused kernel (for bench.py / profiling): CTAs, threads, dynamic smem */
int mb2_launch_info(mb2_ctx* ctx, int32_t* grid, int32_t* block, int32_t* smem_bytes);
/* number of kernels this library has launched through the context (bench.py gpu_launches) */
int64_t mb2_launch_count(mb2_ctx* ctx);
/* which kernel family the last fused launch (generation 0, step or evaluation) used: the tests assert
 * that a case runs on the kernel it is meant to cover */
enum mb2_kernel_family_id {
  MB2_KERNEL_NONE = 0,
  MB2_KERNEL_WARP_PER_ROW = 1,   /* de_kernels.cuh: plain loads, one warp per row */
  MB2_KERNEL_TMA_BIN = 2,        /* de_tma.cuh: Best1Bin, whole rows staged by cp.async.bulk */
  MB2_KERNEL_TMA_EXP = 3,        /* de_tma_exp.cuh: exponential crossover, parent row + donor windows staged */
  MB2_KERNEL_ROWS = 4,           /* de_rows.cuh: 16..128-dimensional rows, tiles of rows per warp */
  MB2_KERNEL_ROWS_PER_WARP = 5,  /* de_subwarp.cuh */
  MB2_KERNEL_THREAD_PER_ROW = 6, /* de_small.cuh */
  MB2_KERNEL_DMMA = 7,           /* de_cheb_mma.cuh: FP64 tensor-core contraction */
  MB2_KERNEL_ISLANDS = 8,        /* batch of solvers, grid.y = island */
  MB2_KERNEL_ROWS_BULK = 9,      /* de_rows_bulk.cuh: the rows kernel's step with cp.async.bulk staging */
  MB2_KERNEL_TMA_POOL = 10       /* de_tma_pool.cuh: Best1Bin committed in place, row buffers pooled behind producer warps */
};
int mb2_kernel_family(mb2_ctx* ctx);

/* How a generation's rows leave the fused kernel.  The reference keeps ONE population and a trialSolution
 * list and copies the accepted trials over their parents after the whole generation has been evaluated
 * (mystic/differential_evolution.py:603-609).  MB2_COMMIT_IN_PLACE is that layout on the device: the kernels
 * of rows of 16 dimensions and more (MB2_KERNEL_TMA_BIN / TMA_EXP / ROWS_BULK / ROWS / WARP_PER_ROW) write only
 * ACCEPTED trials to the other buffer and de_commit_accepted_kernel moves them over their parents once every
 * row has read its donors; a rejected row is never written.  MB2_COMMIT_DOUBLE_BUFFER writes every selected row
 * to the other buffer and swaps (the only mode of the rows-per-warp, thread-per-row, DMMA and batched kernels).  MB2_COMMIT_AUTO (default) commits in place while the
 * accepted fraction of the last generation the host read back is below 0.30 (0.12 for rows of up to 128
 * dimensions) -- in place moves (1 + k + 3a) rows per candidate, the double buffer (2 + k).  Results are bit-identical in every mode.
 * The environment variable MB2_COMMIT_MODE (0, 1, 2) overrides the context's setting. */
enum mb2_commit_mode_id { MB2_COMMIT_AUTO = 0, MB2_COMMIT_DOUBLE_BUFFER = 1, MB2_COMMIT_IN_PLACE = 2 };
int mb2_set_commit_mode(mb2_ctx* ctx, int mode);
/* 1 when the last generation was committed in place, 0 when the buffers were swapped */
int mb2_last_commit_in_place(mb2_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MYSTIC_B200_H */
