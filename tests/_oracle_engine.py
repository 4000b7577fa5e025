"""TEST-ONLY engine: the CPU oracle (oracle/de_oracle.py) behind the DeviceEngine interface.

It lets the `-m "not gpu"` suite exercise the HOST logic of mystic_b200.solver (Step / Solve /
termination / keyword handling / sharding and the best exchange over gloo) where no GPU exists.
It lives under tests/ and is only ever injected explicitly (`engine=OracleEngine`); the product
never selects it -- without the CUDA library and a GPU, mystic_b200 raises.
"""
import numpy as np
import torch

from oracle import de_oracle as orc

_STRATEGY_BY_ID = dict((v[0], k) for k, v in orc.STRATEGIES.items())
_COST_NAMES = {0: 'rosen', 1: 'corana', 2: 'griewangk', 3: 'zimmermann', 9: 'sphere', 10: 'rastrigin', 11: 'schwefel',
               12: 'ackley', 13: 'step', 14: 'michal', 15: 'powers', 16: 'branins', 17: 'easom', 18: 'goldstein', 19: 'peaks',
               20: 'venkat91', 21: 'fosc3d', 22: 'nmin51', 23: 'shekel', 24: 'ellipsoid', 25: 'paviani'}   # 4/5: chebyshev (Horner / DMMA)
_PTYPE_BY_ID = dict((v, k) for k, v in orc.PTYPES.items())


class OracleEngine(object):
    def __init__(self, np_local, dim, np_global=None, offset=0, device=None, keep_trials=False):
        self.np_local, self.dim = int(np_local), int(dim)
        self.np_global = int(np_global if np_global is not None else np_local)
        self.offset = int(offset)
        self._pop = np.zeros((self.np_local, self.dim))
        self.state = None
        self.strategy, self.F, self.CR = 'Best1Bin', 0.8, 0.9
        self.bounds = (orc.BOUNDS_NONE, None, None)
        self.cost = None
        self.pen = ()
        self.seed = 0
        self.replay = None
        self.n_inf = self.n_acc = 0
        self.best_idx = -1
        self.launches = 0

    def close(self):
        pass

    def set_strategy(self, sid, scale, cross):
        self.strategy, self.F, self.CR = _STRATEGY_BY_ID[sid], float(scale), float(cross)

    def set_bounds(self, mode, lo=None, hi=None):
        self.bounds = (mode, None if lo is None else np.array(lo, float), None if hi is None else np.array(hi, float))
        if self.state is not None:
            self.state.bounds_mode, self.state.lo, self.state.hi = self.bounds

    def set_cost(self, cost_id, params=()):
        if cost_id in _COST_NAMES:
            self.cost = orc.make_cost(_COST_NAMES[cost_id])
        elif cost_id in (6, 7, 8, 26):   # data-fit costs: params = [M, sigma, pts[M], obs[M][, sigma[M]]]
            M, sigma = int(params[0]), float(params[1])
            pts, obs = np.asarray(params[2:2 + M], float), np.asarray(params[2 + M:2 + 2 * M], float)
            if len(params) == 2 + 3 * M:
                sigma = np.asarray(params[2 + 2 * M:], float)
            self.cost = orc.make_cost({6: 'polyfit', 7: 'polyfit', 8: 'lorentzian', 26: 'decay'}[cost_id], (pts, obs, sigma))
        else:
            M, order = int(params[0]), int(params[1])
            self.cost = orc.make_cost('chebyshev%dcost' % order, (M,))

    def set_penalty(self, ptypes=(), G=(), h=(), kscale=(), mult=None):
        mult = [0.0] * len(ptypes) if mult is None else mult
        self.pen = [dict(ptype=_PTYPE_BY_ID[int(p)], G=list(g), h=float(hh), kscale=float(k), mult=float(m))
                    for p, g, hh, k, m in zip(ptypes, G, h, kscale, mult)]

    def set_rng_philox(self, seed):
        self.seed = int(seed)

    def set_rng_replay(self, donors, nstart, uniforms):
        self.replay = dict(donors=np.asarray(donors), nstart=np.asarray(nstart), uni=np.asarray(uniforms))

    def init_population(self, seed, lo, hi):
        self._pop = orc.philox_init_population(int(seed), self.np_local, lo, hi, offset=self.offset)
        self.state = None

    def upload_population(self, pop):
        self._pop = np.array(pop, dtype=np.float64)
        self.state = None

    def restore_state(self, energies, best_energy, best_index, best_x, generation):
        mode, lo, hi = self.bounds
        self.state = orc.DEState(self._pop, lo, hi, mode)
        self.state.popEnergy = np.array(energies, dtype=np.float64)
        self.state.bestEnergy = float(best_energy)
        self.state.bestSolution = np.array(best_x, dtype=np.float64)
        self.state.generation = int(generation)
        self.best_idx = int(best_index)

    # ---- batch of independent solvers: island i is shard i of the rows, without a best exchange
    def set_islands(self, n, share_best_within=1):
        n = int(n)
        if n < 1 or self.np_local % n:
            raise ValueError('%d rows do not split into %d equal populations' % (self.np_local, n))
        if share_best_within < 1 or n % share_best_within:
            raise ValueError('%d islands do not split into groups of %d' % (n, share_best_within))
        self.n_islands = n
        self.island_group = int(share_best_within)
        self._isl = None

    def _share(self):
        """islands_share_best_kernel (mystic_b200/csrc/de_misc.cuh): the group's winner goes to every island"""
        w = getattr(self, 'island_group', 1)
        if w <= 1:
            return
        isl = self._islands()
        for g0 in range(0, len(isl), w):
            grp = [e for e in isl[g0:g0 + w] if e.best_idx >= 0]
            if not grp:
                continue
            win = min(grp, key=lambda e: (e.state.bestEnergy, e.best_idx))
            for e in isl[g0:g0 + w]:
                if e is not win:
                    e.state.bestEnergy = win.state.bestEnergy
                    e.state.bestSolution = win.state.bestSolution.copy()
                    e.best_idx = win.best_idx

    def _islands(self):
        if getattr(self, '_isl', None) is None:
            inp = self.np_local // self.n_islands
            self._isl = []
            for i in range(self.n_islands):
                e = OracleEngine(inp, self.dim, self.np_global, self.offset + i * inp)
                e.cost, e.pen, e.bounds, e.seed = self.cost, self.pen, self.bounds, self.seed
                e._pop = self._pop[i * inp:(i + 1) * inp].copy()
                self._isl.append(e)
        return self._isl

    def read_islands(self):
        rs = [e.read_result() for e in self._islands()]
        return (np.array([r[0] for r in rs]), np.array([r[1] for r in rs], dtype=np.int64),
                np.array([r[2] for r in rs], dtype=np.int64), np.array([r[3] for r in rs], dtype=np.int64),
                int(rs[0][4]), np.array([r[5] for r in rs]))

    def population_tensor(self):
        arr = self._pop if self.state is None else self.state.population
        return torch.from_numpy(arr)

    def population(self):
        if getattr(self, 'n_islands', 1) > 1 and getattr(self, '_isl', None):
            return np.concatenate([e.population() for e in self._isl])
        return (self._pop if self.state is None else self.state.population).copy()

    def energies(self):
        if getattr(self, 'n_islands', 1) > 1 and getattr(self, '_isl', None):
            return np.concatenate([e.energies() for e in self._isl])
        return np.full(self.np_local, np.inf) if self.state is None else self.state.popEnergy.copy()

    def trials(self):
        return self.state.trialSolution.copy(), self.state.trialEnergy.copy()

    def _after(self):
        st = self.state
        self.n_inf = int(np.isinf(st.trialEnergy).sum())
        self.n_acc = int(st.accepted.sum())
        if st.bestIndex >= 0:   # select() found a new best in this generation
            self.best_idx = self.offset + st.bestIndex
        self.launches += 2

    def eval_initial(self):
        if getattr(self, 'n_islands', 1) > 1:
            for e in self._islands():
                e.eval_initial()
            self._share()
            self.launches += 2
            return
        mode, lo, hi = self.bounds
        self.state = orc.DEState(self._pop, lo, hi, mode)
        self.best_idx = -1
        orc.step0(self.state, self.cost, self.pen)
        self._after()

    def step(self):
        if getattr(self, 'n_islands', 1) > 1:
            for e in self._islands():
                e.strategy, e.F, e.CR = self.strategy, self.F, self.CR
                e.step()
            self._share()
            self.launches += 2
            return
        st = self.state
        st.bestIndex = -1
        if self.replay is not None:
            orc.step(st, self.strategy, self.CR, self.F, self.cost, self.pen, replay=self.replay)
            self.replay = None
        else:
            orc.step(st, self.strategy, self.CR, self.F, self.cost, self.pen,
                     philox=dict(seed=self.seed, offset=self.offset))
        self._after()

    def run(self, max_steps, term, history, generations, evaluations, max_generations, max_evaluations):
        """CPU twin of DeviceEngine.run: same termination rules as de_finalize_kernel"""
        hist = list(history)
        gens, evals = int(generations), int(evaluations)
        recs = []
        stopped = False
        g = int(term.get('lookback', 0))
        for _ in range(max_steps):
            self.step()
            st = self.state
            recs.append(np.concatenate([[st.bestEnergy, float(self.best_idx), float(self.n_inf), float(self.n_acc)],
                                        st.bestSolution]))
            evals += self.np_global - self.n_inf
            gens += 1
            hist.append(st.bestEnergy)
            lg, h1 = len(hist), hist[-1]
            hg = (hist[-g] if g else hist[0]) if lg > g else h1
            kind = term['kind']
            met = False
            if kind == 1:
                met = abs(h1 - term['target']) <= term['tolerance']
            elif kind == 2:
                met = lg > g and ((hg - h1) <= term['tolerance'] or hg == h1)
            elif kind == 3:
                met = lg > g and (hg == h1 or 2.0 * (hg - h1) <= term['tolerance'] * (abs(hg) + abs(h1)) + 1e-20)
            elif kind == 4:
                met = (lg > g and ((hg - h1) <= term['gtol'] or hg == h1)) or abs(h1 - term['target']) <= term['tolerance']
            elif kind == 5:
                met = (term['term_evaluations'] >= 0 and evals >= term['term_evaluations']) or \
                      (term['term_generations'] >= 0 and gens >= term['term_generations'])
            if max_evaluations is not None and evals >= max_evaluations:
                met = True
            if max_generations is not None and gens >= max_generations:
                met = True
            if met:
                stopped = True
                break
        return np.array(recs).reshape(len(recs), self.dim + 4), stopped

    @staticmethod
    def _met(term, hist, gens, evals, max_generations, max_evaluations):
        """the termination rules of de_finalize_kernel (mystic_b200/csrc/de_misc.cuh)"""
        g = int(term.get('lookback', 0))
        lg, h1 = len(hist), hist[-1]
        hg = (hist[-g] if g else hist[0]) if lg > g else h1
        kind = term['kind']
        met = False
        if kind == 1:
            met = abs(h1 - term['target']) <= term['tolerance']
        elif kind == 2:
            met = lg > g and ((hg - h1) <= term['tolerance'] or hg == h1)
        elif kind == 3:
            met = lg > g and (hg == h1 or 2.0 * (hg - h1) <= term['tolerance'] * (abs(hg) + abs(h1)) + 1e-20)
        elif kind == 4:
            met = (lg > g and ((hg - h1) <= term['gtol'] or hg == h1)) or abs(h1 - term['target']) <= term['tolerance']
        elif kind == 5:
            met = (term['term_evaluations'] >= 0 and evals >= term['term_evaluations']) or \
                  (term['term_generations'] >= 0 and gens >= term['term_generations'])
        if max_evaluations is not None and evals >= max_evaluations:
            met = True
        if max_generations is not None and gens >= max_generations:
            met = True
        return met

    def set_islands_termination(self, term, max_generations, max_evaluations):
        self._isl_term = (dict(term), max_generations, max_evaluations)
        self._isl_run = [dict(hist=[], gens=0, evals=0, stopped=False) for _ in range(self.n_islands)]

    def run_islands(self, max_steps):
        """CPU twin of DeviceEngine.run_islands: every island logs and stops on its own condition"""
        term, maxgen, maxfun = self._isl_term
        isl = self._islands()
        n = len(isl)
        recs = np.zeros((n, max_steps, self.dim + 4))
        done = np.zeros(n, dtype=np.int32)
        for _ in range(max_steps):
            for i, e in enumerate(isl):
                r = self._isl_run[i]
                if r['stopped']:
                    continue
                if e.state is None:
                    e.eval_initial()
                else:
                    e.strategy, e.F, e.CR = self.strategy, self.F, self.CR
                    e.step()
                    r['gens'] += 1
                st = e.state
                recs[i, done[i]] = np.concatenate([[st.bestEnergy, float(e.best_idx), float(e.n_inf), float(e.n_acc)], st.bestSolution])
                done[i] += 1
                r['evals'] += e.np_local - e.n_inf
                r['hist'].append(st.bestEnergy)
                r['stopped'] = self._met(term, r['hist'], r['gens'], r['evals'], maxgen, maxfun)
            self.launches += 2
        return recs, done, np.array([r['stopped'] for r in self._isl_run])

    def pack_best(self):
        st = self.state
        rec = np.concatenate([[st.bestEnergy, float(self.best_idx), float(self.n_inf), float(self.n_acc)],
                              st.bestSolution])
        return torch.from_numpy(rec)

    def merge_best(self, records):
        """same rule as merge_best_kernel (mystic_b200/csrc/de_kernels.cuh)"""
        recs = records.numpy()
        D = self.dim
        w, be, bi = -1, np.inf, 0.0
        ninf = nacc = 0.0
        for r in range(recs.shape[0]):
            rec = recs[r]
            ninf += rec[2]
            nacc += rec[3]
            if rec[1] < 0:
                continue
            if w < 0 or rec[0] < be or (rec[0] == be and rec[1] < bi):
                w, be, bi = r, rec[0], rec[1]
        st = self.state
        if w >= 0:
            st.bestEnergy = float(be)
            st.bestSolution = recs[w][4:4 + D].copy()
            self.best_idx = int(bi)
        self.n_inf, self.n_acc = int(ninf), int(nacc)

    def read_result(self):
        st = self.state
        return float(st.bestEnergy), int(self.best_idx), self.n_inf, self.n_acc, st.generation, st.bestSolution.copy()

    def eval_cost(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        mode, lo, hi = self.bounds
        tmp = orc.DEState(x, lo, hi, mode)
        return orc.evaluate(x, self.cost, tmp, self.pen)

    def launch_info(self):
        return (0, 0, 0)

    def launch_count(self):
        return self.launches
