"""ctypes face of oracle/de_oracle.c (TEST INFRASTRUCTURE ONLY: checker and CPU baseline)."""
import ctypes
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libde_oracle.so')

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_int64)


class Config(ctypes.Structure):
    _fields_ = [('dim', ctypes.c_int32), ('strategy', ctypes.c_int32), ('np', ctypes.c_int64),
                ('F', ctypes.c_double), ('CR', ctypes.c_double), ('bounds_mode', ctypes.c_int32),
                ('cost', ctypes.c_int32), ('lo', _dp), ('hi', _dp), ('cheb_M', ctypes.c_int32),
                ('cheb_order', ctypes.c_int32), ('cheb_target', _dp), ('n_pen', ctypes.c_int32),
                ('pad_', ctypes.c_int32), ('pen_type', _ip), ('pen_G', _dp), ('pen_h', _dp), ('pen_k', _dp),
                ('fit_M', ctypes.c_int32), ('pad2_', ctypes.c_int32), ('fit_pts', _dp), ('fit_obs', _dp),
                ('fit_sigma', ctypes.c_double), ('fit_sigma_v', _dp), ('pen_mult', _dp)]


_lib = None


def build():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, 'de_oracle.c')):
        subprocess.run(['make', '-s', '-C', HERE], check=True)
    return LIB


def available():
    try:
        load()
        return True
    except Exception:
        return False


def load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


class COracle(object):
    """Same state machine as oracle.de_oracle.DEState + step0/step, in C with OpenMP."""

    COSTS = {'rosen': 0, 'corana': 1, 'griewangk': 2, 'zimmermann': 3, 'sphere': 9, 'rastrigin': 10, 'schwefel': 11, 'ackley': 12, 'step': 13, 'michal': 14, 'powers': 15,
             'branins': 16, 'easom': 17, 'goldstein': 18, 'peaks': 19, 'venkat91': 20, 'fosc3d': 21, 'nmin51': 22, 'shekel': 23,
             'ellipsoid': 24, 'paviani': 25}

    def __init__(self, population, strategy, CR, F, cost, extra=(), lo=None, hi=None, bounds_mode=0, penalty=(), trials=True):
        from oracle import de_oracle as orc
        self.lib = load()
        self.pop = [np.ascontiguousarray(population, dtype=np.float64), None]
        self.NP, self.D = self.pop[0].shape
        self.pop[1] = np.empty_like(self.pop[0])
        self.en = [np.full(self.NP, np.inf), np.full(self.NP, np.inf)]
        self.cur = 0
        self.best_x = np.zeros(self.D)
        self.best_e = ctypes.c_double(np.inf)
        self.best_idx = ctypes.c_int64(-1)
        self.trial = np.zeros((self.NP, self.D)) if trials else None   # full-size runs skip the NP x D capture
        self.trial_e = np.zeros(self.NP)
        self.n_inf = ctypes.c_int64(0)
        self.n_acc = ctypes.c_int64(0)
        self.generation = -1
        self.evaluations = 0
        self._keep = []
        cfg = Config()
        cfg.dim, cfg.np = self.D, self.NP
        cfg.strategy = orc.STRATEGIES[strategy][0]
        cfg.F, cfg.CR = float(F), float(CR)
        cfg.bounds_mode = int(bounds_mode)
        if bounds_mode:
            self._lo = np.ascontiguousarray(lo, dtype=np.float64)
            self._hi = np.ascontiguousarray(hi, dtype=np.float64)
            cfg.lo, cfg.hi = _d(self._lo), _d(self._hi)
        cfg.fit_sigma = 1.0
        if cost in self.COSTS:
            cfg.cost = self.COSTS[cost]
        elif cost in ('polyfit', 'lorentzian', 'decay'):   # extra = (pts, observations[, sigma]), sigma a scalar or one per point
            cfg.cost = {'polyfit': 6, 'lorentzian': 8, 'decay': 26}[cost]
            self._fit_pts = np.ascontiguousarray(extra[0], dtype=np.float64)
            self._fit_obs = np.ascontiguousarray(extra[1], dtype=np.float64)
            cfg.fit_M = int(self._fit_pts.size)
            cfg.fit_pts, cfg.fit_obs = _d(self._fit_pts), _d(self._fit_obs)
            sg = extra[2] if len(extra) > 2 else 1.0
            if np.ndim(sg) > 0:
                self._fit_sigma_v = np.ascontiguousarray(sg, dtype=np.float64)
                cfg.fit_sigma, cfg.fit_sigma_v = 1.0, _d(self._fit_sigma_v)
            else:
                cfg.fit_sigma = float(sg)
        else:
            order = int(cost[len('chebyshev'):-len('cost')])
            cfg.cost = 4
            cfg.cheb_M = int(extra[0]) if extra else 61
            cfg.cheb_order = order
            self._target = np.ascontiguousarray(orc.CHEBYSHEV_TARGETS[order], dtype=np.float64)
            cfg.cheb_target = _d(self._target)
        if penalty:
            self._pt = np.ascontiguousarray([orc.PTYPES[t['ptype']] for t in penalty], dtype=np.int32)
            self._G = np.ascontiguousarray([t['G'] for t in penalty], dtype=np.float64)
            self._h = np.ascontiguousarray([t['h'] for t in penalty], dtype=np.float64)
            ks, ms = [], []
            for t in penalty:
                if 'kscale' in t:
                    ks.append(t['kscale'])
                    ms.append(t.get('mult', 0.0))
                elif t['ptype'] in ('lagrange_inequality', 'lagrange_equality'):
                    k_, m_ = orc.lagrange_state(t['ptype'], t.get('k', 20), t.get('hpow', 5), t.get('n', 0), t.get('stored', ()))
                    ks.append(k_)
                    ms.append(m_)
                else:
                    ks.append(t.get('k', 100) * pow(t.get('hpow', 5), t.get('n', 0)))
                    ms.append(0.0)
            self._k = np.ascontiguousarray(ks, dtype=np.float64)
            self._m = np.ascontiguousarray(ms, dtype=np.float64)
            cfg.n_pen = len(penalty)
            cfg.pen_type = self._pt.ctypes.data_as(_ip)
            cfg.pen_G, cfg.pen_h, cfg.pen_k, cfg.pen_mult = _d(self._G), _d(self._h), _d(self._k), _d(self._m)
        self.cfg = cfg

    @property
    def population(self):
        return self.pop[self.cur]

    @property
    def popEnergy(self):
        return self.en[self.cur]

    @property
    def bestEnergy(self):
        return self.best_e.value

    def _after(self):
        self.cur ^= 1
        self.generation += 1
        self.evaluations += self.NP - self.n_inf.value

    def step0(self):
        c, n = self.cur, self.cur ^ 1
        self.lib.orc_step0(ctypes.byref(self.cfg), _d(self.pop[c]), _d(self.pop[n]), _d(self.en[n]), _d(self.best_x),
                           ctypes.byref(self.best_e), ctypes.byref(self.best_idx), _d(self.trial_e),
                           ctypes.byref(self.n_inf), ctypes.byref(self.n_acc))
        if self.trial is not None:
            self.trial[:] = self.pop[n]
        self._after()

    def step(self, replay=None, seed=0, offset=0, keep_trials=True):
        c, n = self.cur, self.cur ^ 1
        if replay is not None:
            d = np.full((self.NP, 5), -1, dtype=np.int32)
            dd = np.asarray(replay['donors'])
            d[:, :dd.shape[1]] = dd[:, :5]
            ns = np.ascontiguousarray(replay['nstart'], dtype=np.int32)
            u = np.full((self.NP, self.D + 1), np.nan)
            uu = np.asarray(replay['uni'], dtype=np.float64)
            u[:, :uu.shape[1]] = uu[:, :self.D + 1]
            mode, dp, nsp, up = 1, d.ctypes.data_as(_ip), ns.ctypes.data_as(_ip), _d(u)
        else:
            mode, dp, nsp, up = 0, None, None, None
        self.lib.orc_step(ctypes.byref(self.cfg), _d(self.pop[c]), _d(self.pop[n]), _d(self.en[c]), _d(self.en[n]),
                          _d(self.best_x), ctypes.byref(self.best_e), ctypes.byref(self.best_idx),
                          ctypes.c_int(mode), ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                          ctypes.c_uint32(self.generation + 1), ctypes.c_int64(offset), dp, nsp, up,
                          _d(self.trial) if (keep_trials and self.trial is not None) else None, _d(self.trial_e),
                          ctypes.byref(self.n_inf), ctypes.byref(self.n_acc))
        self._after()


def init_population(seed, NP, lo, hi, offset=0):
    lib = load()
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    pop = np.empty((NP, lo.shape[0]))
    lib.orc_init_population(_d(pop), ctypes.c_int64(NP), ctypes.c_int(lo.shape[0]), ctypes.c_int64(offset), _d(lo),
                            _d(hi), ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF))
    return pop


def time_workload(w, NP, steps, warmup, seed=0x5EED, threads=None):
    """bench.py cpu_baseline / --impl reference: -> (cand-evals/s, s/step, kind, cores, sample).
    `threads`: OpenMP threads to use (torchrun exports OMP_NUM_THREADS=1; the bench names the count itself)."""
    if threads:
        load().orc_set_num_threads(int(threads))
    D = w['dim']
    lo, hi = np.full(D, w['lo']), np.full(D, w['hi'])
    mode = {None: 0, 'reject': 1, 'clamp': 2}[w['bounds']]
    pen = ()
    if w['penalty']:
        p = w['penalty']
        pen = [dict(ptype='quadratic_inequality', G=[p['G']] * D, h=p['h'], k=p['k'], hpow=p['hpow'])]
    o = COracle(init_population(seed, NP, lo, hi), w['strategy'], w['CR'], w['F'], w['cost'], w['extra'], lo, hi, mode, pen)
    o.step0()
    for _ in range(warmup):
        o.step(seed=seed, keep_trials=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(seed=seed, keep_trials=False)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    cores = load().orc_num_threads()
    return NP / dt, dt, 'port', cores, ('C/OpenMP oracle port (oracle/de_oracle.c, %d threads), NP=%d of %d rows, '
                                        '%d generations' % (cores, NP, w['np'], steps))
