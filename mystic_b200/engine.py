"""Python face of the C ABI: one `DeviceEngine` per population shard on one GPU.

PyTorch owns the big device buffers (population double buffer, energies, replay draws) and the
stream; every computation is a call into libmystic_b200.so (include/mystic_b200.h).  There is
no CPU path: constructing an engine without a CUDA device raises.
"""
import ctypes

import numpy as np
import torch

from . import _native
from ._native import check, lib

BOUNDS_NONE, BOUNDS_REJECT, BOUNDS_CLAMP = 0, 1, 2


def _dptr(arr):
    return arr.ctypes.data_as(_native.c_double_p)


def probe_fp64_peak(device=None, iters=20000):
    """-> (DFMA TFLOP/s, DMMA m8n8k4 TFLOP/s) of `device`, measured now (mb2_probe_fp64_peak)."""
    if not torch.cuda.is_available():
        raise RuntimeError('mystic_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    if device is None:
        device = torch.cuda.current_device()
    dfma, dmma = ctypes.c_double(0.0), ctypes.c_double(0.0)
    check(lib.mb2_probe_fp64_peak(int(device), int(iters), ctypes.byref(dfma), ctypes.byref(dmma),
                                  ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return dfma.value, dmma.value


class DeviceEngine(object):
    """Device-resident state of one shard: rows [offset, offset+np_local) of np_global."""

    def __init__(self, np_local, dim, np_global=None, offset=0, device=None, keep_trials=False):
        if not torch.cuda.is_available():
            raise RuntimeError('mystic_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        self.np_local = int(np_local)
        self.dim = int(dim)
        self.np_global = int(np_global if np_global is not None else np_local)
        self.offset = int(offset)
        f64 = dict(dtype=torch.float64, device=self.device)
        self._pop = [torch.empty((self.np_local, self.dim), **f64) for _ in range(2)]
        self._en = [torch.full((self.np_local,), float('inf'), **f64) for _ in range(2)]
        # the context's small buffers live in workspaces owned by torch's caching allocators
        hb = ctypes.c_int64(0)
        db = int(lib.mb2_workspace_bytes(self.dim, ctypes.byref(hb)))
        self._ws_dev = torch.empty((db,), dtype=torch.uint8, device=self.device)
        self._ws_host = torch.empty((hb.value,), dtype=torch.uint8, pin_memory=True)
        self._ctx = ctypes.c_void_p()
        check(lib.mb2_create_ws(self.device.index, self.np_local, self.dim, self.np_global, self.offset,
                                self._ws_dev.data_ptr(), db, self._ws_host.data_ptr(), hb.value,
                                ctypes.byref(self._ctx)))
        check(lib.mb2_bind_buffers(self._ctx, self._pop[0].data_ptr(), self._pop[1].data_ptr(),
                                   self._en[0].data_ptr(), self._en[1].data_ptr()))
        self._trial = self._trial_e = None
        if keep_trials:
            self._trial = torch.zeros((self.np_local, self.dim), **f64)
            self._trial_e = torch.full((self.np_local,), float('inf'), **f64)
            check(lib.mb2_set_capture(self._ctx, self._trial.data_ptr(), self._trial_e.data_ptr()))
        self._replay = None  # keeps the uploaded draw tensors alive until consumed
        self.n_islands = 1
        self.peer_exchange = False  # True: eval_initial()/step() end with the exchange of the shards' bests
        self._log = self._log_host = None  # batch-mode generation log
        self._record = torch.empty((self.dim + 4,), **f64)
        self._xbuf = np.empty(self.dim, dtype=np.float64)

    def close(self):
        if getattr(self, '_ctx', None) is not None and self._ctx.value:
            lib.mb2_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- configuration -------------------------------------------------------------------
    def set_strategy(self, strategy_id, scale, cross):
        check(lib.mb2_set_strategy(self._ctx, int(strategy_id), float(scale), float(cross)))

    def set_bounds(self, mode, lo=None, hi=None):
        if mode == BOUNDS_NONE:
            check(lib.mb2_set_bounds(self._ctx, BOUNDS_NONE, None, None))
            return
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        check(lib.mb2_set_bounds(self._ctx, int(mode), _dptr(lo), _dptr(hi)))

    def set_cost(self, cost_id, params=()):
        p = np.ascontiguousarray(params, dtype=np.float64)
        check(lib.mb2_set_cost(self._ctx, int(cost_id), _dptr(p) if p.size else None, int(p.size)))

    def set_penalty(self, ptypes=(), G=(), h=(), kscale=(), mult=None):
        """stacked penalty terms (innermost first); `mult`: the lagrange kinds' accumulated multipliers"""
        n = len(ptypes)
        if n == 0:
            check(lib.mb2_set_penalty(self._ctx, 0, None, None, None, None))
            return
        pt = np.ascontiguousarray(ptypes, dtype=np.int32)
        Gm = np.ascontiguousarray(G, dtype=np.float64).reshape(n, self.dim)
        hv = np.ascontiguousarray(h, dtype=np.float64)
        kv = np.ascontiguousarray(kscale, dtype=np.float64)
        mv = np.zeros(n, dtype=np.float64) if mult is None else np.ascontiguousarray(mult, dtype=np.float64)
        check(lib.mb2_set_penalty_ex(self._ctx, n, pt.ctypes.data_as(_native.c_int32_p), _dptr(Gm), _dptr(hv), _dptr(kv),
                                     _dptr(mv)))

    def set_rng_philox(self, seed):
        check(lib.mb2_set_rng_philox(self._ctx, int(seed) & 0xFFFFFFFFFFFFFFFF))

    def set_rng_replay(self, donors, nstart, uniforms):
        """recorded draws for the next step: donors [np,<=5], nstart [np], uniforms [np, dim+1] (NaN = not drawn)"""
        d = np.full((self.np_local, 5), -1, dtype=np.int32)
        donors = np.asarray(donors)
        d[:, :donors.shape[1]] = donors[:, :5]
        u = np.full((self.np_local, self.dim + 1), np.nan, dtype=np.float64)
        uniforms = np.asarray(uniforms, dtype=np.float64)
        u[:, :uniforms.shape[1]] = uniforms[:, :self.dim + 1]
        t = (torch.from_numpy(d).to(self.device), torch.from_numpy(np.ascontiguousarray(nstart, dtype=np.int32)).to(self.device),
             torch.from_numpy(u).to(self.device))
        self._replay = t
        check(lib.mb2_set_rng_replay(self._ctx, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr()))

    # ---- population ----------------------------------------------------------------------
    def init_population(self, seed, lo, hi):
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        check(lib.mb2_init_population(self._ctx, int(seed) & 0xFFFFFFFFFFFFFFFF, _dptr(lo), _dptr(hi), self._stream()))

    def upload_population(self, pop):
        pop = np.ascontiguousarray(pop, dtype=np.float64)
        if pop.shape != (self.np_local, self.dim):
            raise ValueError('population shape %r != %r' % (pop.shape, (self.np_local, self.dim)))
        check(lib.mb2_upload_population(self._ctx, pop.ctypes.data_as(ctypes.c_void_p), self._stream()))

    def restore_state(self, energies, best_energy, best_index, best_x, generation):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        x = np.ascontiguousarray(best_x, dtype=np.float64)
        check(lib.mb2_restore_state(self._ctx, _dptr(e), float(best_energy), int(best_index), _dptr(x), int(generation)))

    def _current(self):
        p, e = ctypes.c_void_p(), ctypes.c_void_p()
        check(lib.mb2_current(self._ctx, ctypes.byref(p), ctypes.byref(e)))
        i = 0 if p.value == self._pop[0].data_ptr() else 1
        return self._pop[i], self._en[i]

    def population_tensor(self):
        return self._current()[0]

    def energy_tensor(self):
        return self._current()[1]

    def population(self):
        return self._current()[0].cpu().numpy()

    def energies(self):
        return self._current()[1].cpu().numpy()

    def trials(self):
        if self._trial is None:
            raise RuntimeError('trial vectors are not kept; construct the solver with keep_trials=True')
        return self._trial.cpu().numpy(), self._trial_e.cpu().numpy()

    # ---- the hot path --------------------------------------------------------------------
    def eval_initial(self):
        check(lib.mb2_eval_initial(self._ctx, self._stream()))

    def step(self):
        check(lib.mb2_step(self._ctx, self._stream()))

    # ---- device-resident multi-generation stepping -----------------------------------------
    def run(self, max_steps, term, history, generations, evaluations, max_generations, max_evaluations):
        """Enqueue up to max_steps generations with the termination test on the device; returns
        (records [executed, dim+4], stopped).  `term` = dict(kind, lookback, tolerance, target, gtol,
        term_generations, term_evaluations) as produced by mystic_b200.termination.device_descriptor."""
        t = _native.Termination()
        t.kind, t.lookback = int(term['kind']), int(term.get('lookback', 0))
        t.tolerance, t.target, t.gtol = float(term.get('tolerance', 0.0)), float(term.get('target', 0.0)), float(term.get('gtol', 0.0))
        t.term_generations = int(term.get('term_generations', -1))
        t.term_evaluations = int(term.get('term_evaluations', -1))
        t.max_generations = -1 if max_generations is None else int(max_generations)
        t.max_evaluations = -1 if max_evaluations is None else int(max_evaluations)
        t.generations, t.evaluations = int(generations), int(evaluations)
        tail = np.ascontiguousarray(history[-1024:] if len(history) > 1024 else history, dtype=np.float64)
        if len(history) > 1024:
            # the device ring is indexed by absolute position: hand over the full length with the tail
            full = np.empty(len(history), dtype=np.float64)
            full[:] = history
            tail = full
        check(lib.mb2_set_termination(self._ctx, ctypes.byref(t), _dptr(tail) if tail.size else None, int(tail.size)))
        if self._log is None or self._log.shape[0] < max_steps:
            self._log = torch.empty((max_steps, self.dim + 4), dtype=torch.float64, device=self.device)
            self._log_host = torch.empty((max_steps, self.dim + 4), dtype=torch.float64, pin_memory=True)
        check(lib.mb2_run(self._ctx, int(max_steps), self._log.data_ptr(), self._stream()))
        done, stopped = ctypes.c_int32(0), ctypes.c_int32(0)
        check(lib.mb2_run_result(self._ctx, ctypes.byref(done), ctypes.byref(stopped), self._stream()))
        n = done.value
        self._log_host[:n].copy_(self._log[:n])
        return self._log_host[:n].numpy().copy(), bool(stopped.value)

    # ---- batch of independent solvers -------------------------------------------------------
    def set_islands(self, n_islands, share_best_within=1):
        """the rows become n_islands independent populations advanced by the same launches; with
        share_best_within = w > 1, groups of w consecutive islands share their best member every generation
        (a group = one population sharded over w ranks, on this GPU)"""
        check(lib.mb2_set_islands(self._ctx, int(n_islands)))
        self.n_islands = int(n_islands)
        if share_best_within != 1:
            check(lib.mb2_set_island_groups(self._ctx, int(share_best_within)))

    def set_islands_termination(self, term, max_generations, max_evaluations):
        """install a device-side termination condition for every island (after set_islands, before generation 0);
        `term` as for run()"""
        t = _native.Termination()
        t.kind, t.lookback = int(term['kind']), int(term.get('lookback', 0))
        t.tolerance, t.target, t.gtol = float(term.get('tolerance', 0.0)), float(term.get('target', 0.0)), float(term.get('gtol', 0.0))
        t.term_generations = int(term.get('term_generations', -1))
        t.term_evaluations = int(term.get('term_evaluations', -1))
        t.max_generations = -1 if max_generations is None else int(max_generations)
        t.max_evaluations = -1 if max_evaluations is None else int(max_evaluations)
        t.generations, t.evaluations = 0, 0
        check(lib.mb2_set_termination(self._ctx, ctypes.byref(t), None, 0))

    def run_islands(self, max_steps):
        """enqueue up to max_steps generations of every island (generation 0 first when it has not run), each island
        logging and stopping on its own condition -> (records [n_islands, max_steps, dim+4], executed[n], stopped[n])"""
        n = self.n_islands
        shape = (n, int(max_steps), self.dim + 4)
        if self._log is None or tuple(self._log.shape) != shape:
            self._log = torch.empty(shape, dtype=torch.float64, device=self.device)
            self._log_host = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        check(lib.mb2_run(self._ctx, int(max_steps), self._log.data_ptr(), self._stream()))
        done = np.zeros(n, dtype=np.int32)
        stopped = np.zeros(n, dtype=np.int32)
        check(lib.mb2_run_result_islands(self._ctx, done.ctypes.data_as(_native.c_int32_p), stopped.ctypes.data_as(_native.c_int32_p),
                                         self._stream()))
        self._log_host.copy_(self._log)
        self._replay = None
        return self._log_host.numpy().copy(), done, stopped.astype(bool)

    def read_islands(self):
        """-> (best_energy[n], best_index[n], n_inf[n], n_accepted[n], generation, bestSolution[n, dim]); synchronises"""
        n = self.n_islands
        res = (_native.Result * n)()
        x = np.empty((n, self.dim), dtype=np.float64)
        check(lib.mb2_read_islands(self._ctx, ctypes.cast(res, ctypes.c_void_p), _dptr(x), self._stream()))
        self._replay = None
        return (np.array([r.best_energy for r in res]), np.array([r.best_index for r in res], dtype=np.int64),
                np.array([r.n_inf for r in res], dtype=np.int64), np.array([r.n_accepted for r in res], dtype=np.int64),
                int(res[0].generation), x)

    # ---- sharded populations: best exchange ------------------------------------------------
    def connect_peers(self, world, rank, all_gather_object, all_agree):
        """fuse the exchange of the shards' bests into the argmin epilogue (stores into peer memory
        over NVLink, include/mystic_b200.h mb2_exchange_*).  torch.distributed plumbing comes from the
        caller: `all_gather_object(obj) -> list` moves the 64-byte CUDA IPC handles between the
        ranks, `all_agree(flag) -> bool` is a logical AND over the ranks."""
        buf = ctypes.create_string_buffer(64)
        check(lib.mb2_exchange_open(self._ctx, int(world), int(rank), buf, 64))
        # mailbox and peer mappings of an earlier engine of this process adopted -- on every rank?
        if all_agree(bool(lib.mb2_exchange_connected(self._ctx))):
            self.peer_exchange = True
            return
        if lib.mb2_exchange_connected(self._ctx):
            check(lib.mb2_exchange_reset(self._ctx))
            check(lib.mb2_exchange_open(self._ctx, int(world), int(rank), buf, 64))
        handles = all_gather_object(buf.raw)
        if len(handles) != world:
            raise RuntimeError('expected %d handles, got %d' % (world, len(handles)))
        check(lib.mb2_exchange_connect(self._ctx, b''.join(handles), 64))
        self.peer_exchange = True

    def exchange_best(self):
        check(lib.mb2_exchange_best(self._ctx, self._stream()))

    def pack_best(self):
        check(lib.mb2_pack_best(self._ctx, self._record.data_ptr(), self._stream()))
        return self._record

    def merge_best(self, records):
        world = records.shape[0]
        check(lib.mb2_merge_best(self._ctx, records.data_ptr(), int(world), self._stream()))

    def read_result(self):
        """-> (best_energy, best_index, n_inf, n_accepted, generation, bestSolution ndarray); synchronises"""
        r = _native.Result()
        check(lib.mb2_read_result(self._ctx, ctypes.byref(r), _dptr(self._xbuf), self._stream()))
        self._replay = None
        return r.best_energy, r.best_index, r.n_inf, r.n_accepted, r.generation, self._xbuf.copy()

    def eval_cost(self, x):
        x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(self.device)
        out = torch.empty((x.shape[0],), dtype=torch.float64, device=self.device)
        check(lib.mb2_eval_cost(self._ctx, x.data_ptr(), int(x.shape[0]), out.data_ptr(), self._stream()))
        return out.cpu().numpy()

    def selftest_division(self, n, seed=1, span=400.0):
        bad = ctypes.c_int64(-1)
        check(lib.mb2_selftest_division(self._ctx, int(n), int(seed), float(span), ctypes.byref(bad)))
        return bad.value

    def selftest_cosine(self, n, seed=1, span=400.0):
        worst = ctypes.c_double(-1.0)
        check(lib.mb2_selftest_cosine(self._ctx, int(n), int(seed), float(span), ctypes.byref(worst)))
        return worst.value

    def launch_info(self):
        g, b, s = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        check(lib.mb2_launch_info(self._ctx, ctypes.byref(g), ctypes.byref(b), ctypes.byref(s)))
        return g.value, b.value, s.value

    KERNEL_FAMILIES = ('none', 'warp_per_row', 'tma_bin', 'tma_exp', 'rows', 'rows_per_warp', 'thread_per_row', 'dmma', 'islands', 'rows_bulk', 'tma_pool')

    def kernel_family(self):
        """kernel family of the last fused launch (mb2_kernel_family)"""
        return self.KERNEL_FAMILIES[int(lib.mb2_kernel_family(self._ctx))]

    COMMIT_MODES = {'auto': 0, 'double_buffer': 1, 'in_place': 2}

    def set_commit_mode(self, mode):
        """'auto' | 'double_buffer' | 'in_place' (mb2_set_commit_mode): whether a generation writes every selected row
        to the other buffer and swaps, or only the accepted trials, which are then moved over their parents in place --
        the reference's population / trialSolution layout (differential_evolution.py:603-609).  Same results."""
        check(lib.mb2_set_commit_mode(self._ctx, self.COMMIT_MODES[mode] if isinstance(mode, str) else int(mode)))

    def last_commit_in_place(self):
        return bool(lib.mb2_last_commit_in_place(self._ctx))

    def launch_count(self):
        return int(lib.mb2_launch_count(self._ctx))


def evaluate_rows(cost, X, extra_args=(), bounds=None, penalty=None, tensor_cores=False):
    """cost(x) for every row of X [n, D], on the device (mb2_eval_cost)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    eng = DeviceEngine(X.shape[0], X.shape[1])
    try:
        eng.set_cost(cost.device_id(tensor_cores, X.shape[1]), cost.params(tuple(extra_args)))
        if bounds is not None:
            eng.set_bounds(*bounds)
        if penalty is not None:
            eng.set_penalty(*penalty.device_terms(X.shape[1]))
        return eng.eval_cost(X)
    finally:
        eng.close()
