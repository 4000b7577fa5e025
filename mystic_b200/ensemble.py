"""Ensemble solvers whose nested solvers are differential-evolution runs advanced TOGETHER on the
device: BuckshotSolver / LatticeSolver (mystic/ensemble.py:44-121 on top of
mystic/abstract_ensemble_solver.py).

The reference launches N nested solvers through `map` (abstract_ensemble_solver.py:737-826: every
nested solver gets `SetInitialPoints(x0_k)`, the strict range, and runs `Solve()` to ITS OWN
termination; the ensemble then reports the nested solver with the lowest energy).  Here the N
nested DifferentialEvolutionSolver2 populations live side by side in one device buffer and every
generation of all of them is ONE launch (include/mystic_b200.h mb2_set_islands); a terminated nested
solver keeps the state it stopped with while the others go on.  Conditions the device can evaluate
(VTR, ChangeOverGeneration, NormalizedChangeOverGeneration, VTRChangeOverGeneration, EvaluationLimits and the
nested solver's own limits) are tested per nested solver inside the argmin epilogue, 64 generations per host
round trip (mb2_run / mb2_run_result_islands); anything else is evaluated on the host every generation.

    solver = BuckshotSolver(dim, npts=64)
    solver.SetNestedSolver(DifferentialEvolutionSolver2(dim, 40))      # instance or class
    solver.SetStrictRanges(lo, hi)
    solver.Solve(models.rosen, termination.VTR(1e-6))
    solver.bestSolution, solver.bestEnergy, solver._all_bestEnergy, solver._total_evals
"""
import random

import numpy as np

from . import models as _models
from . import penalty as _penalty
from . import strategy as _strategy
from . import termination as _termination
from .solver import BOUNDS_CLAMP, BOUNDS_NONE, BOUNDS_REJECT, DifferentialEvolutionSolver2, _no_penalty


class _Member(object):
    """what a termination condition reads from a solver (mystic/termination.py), for one nested solver"""

    def __init__(self, owner, index):
        self._owner, self.id = owner, index
        self.energy_history = []
        self.solution_history = []
        self.bestEnergy = float('inf')
        self.bestSolution = None
        self._fcalls = [0]
        self.message = None          # why it stopped
        self.population = self.popEnergy = self.trialSolution = None

    generations = property(lambda self: max(0, len(self.energy_history) - 1))
    evaluations = property(lambda self: self._fcalls[0])
    nDim = property(lambda self: self._owner.nDim)
    nPop = property(lambda self: self._owner._np_each)
    _maxiter = property(lambda self: self._owner._maxiter)
    _maxfun = property(lambda self: self._owner._maxfun)
    _EARLYEXIT = False

    @property
    def _stepmon(self):
        return self   # len(_stepmon) is the number of recorded generations

    def __len__(self):
        return len(self.energy_history)


class _EnsembleOfDE(object):
    _batch_generations = 64   # generations per host round trip when the device can test the condition (0: every generation)

    def __init__(self, dim, npts, **kwds):
        self.nDim = int(dim)
        self._npts = int(npts)
        self._engine_factory = kwds.pop('engine', None)
        self._device = kwds.pop('device', None)
        self._seed = kwds.pop('seed', None)
        self._nested = DifferentialEvolutionSolver2(self.nDim, 4)
        self._strictMin, self._strictMax, self._tight = [], [], None
        self._defaultMin, self._defaultMax = [-1e3] * self.nDim, [1e3] * self.nDim
        self._termination = _termination.NormalizedChangeOverGeneration(1e-4)   # ensemble.py:60-62
        self._penalty = _no_penalty
        self._maxiter = self._maxfun = None
        self._members = []
        self._best = None
        self._engine = None
        self._init_solution = None

    # ---- configuration (abstract_ensemble_solver.py:170-248, abstract_solver.py) -----------
    def SetNestedSolver(self, solver, **kwds):
        """a mystic_b200 DifferentialEvolutionSolver2 class or configured instance: population size,
        strategy, CrossProbability, ScalingFactor, termination and limits of every nested solver"""
        configured = not isinstance(solver, type)
        if not configured:
            solver = solver(self.nDim, kwds.pop('NP', 4), **kwds)
        if not isinstance(solver, DifferentialEvolutionSolver2):
            raise TypeError('the nested solver must be a mystic_b200 DifferentialEvolutionSolver2 (class or instance)')
        if solver.nDim != self.nDim:
            raise ValueError('nested solver has dimension %d, ensemble %d' % (solver.nDim, self.nDim))
        self._nested = solver
        if configured:
            # a configured instance runs as configured (abstract_ensemble_solver.py:208-215); a class
            # takes the ensemble's termination and limits (:232-233)
            self._termination = solver._termination
            if solver._maxiter is not None:
                self._maxiter = solver._maxiter
            if solver._maxfun is not None:
                self._maxfun = solver._maxfun

    def SetStrictRanges(self, min=None, max=None, tight=None, **kwds):
        self._nested.SetStrictRanges(min, max, tight=tight, **kwds)    # validates like abstract_solver.py:334-405
        self._strictMin, self._strictMax, self._tight = list(min), list(max), tight

    def SetEvaluationLimits(self, generations=None, evaluations=None, **kwds):
        """per nested solver (abstract_solver.py:626-648)"""
        self._maxiter = kwds.get('maxiter', generations)
        self._maxfun = kwds.get('maxfun', evaluations)

    def SetTermination(self, termination):
        self._termination = termination

    def SetPenalty(self, penalty):
        self._penalty = _penalty.resolve(penalty)

    def SetInitialPoints(self, x0):
        """explicit start points, one per nested solver ([npts, dim])"""
        x0 = np.asarray(x0, dtype=np.float64)
        if x0.shape != (self._npts, self.nDim):
            raise ValueError('expected %d start points of dimension %d' % (self._npts, self.nDim))
        self._init_solution = x0.tolist()

    def _bounds(self):
        lower = list(self._strictMin) if len(self._strictMin) else list(self._defaultMin)
        upper = list(self._strictMax) if len(self._strictMax) else list(self._defaultMax)
        return lower, upper

    def _InitialPoints(self):
        raise NotImplementedError

    # ---- the run -------------------------------------------------------------------------
    def _make_engine(self, rows):
        if self._engine_factory is not None:
            return self._engine_factory(rows, self.nDim, device=self._device)
        from .engine import DeviceEngine   # raises without the CUDA library / a GPU: no fallback
        return DeviceEngine(rows, self.nDim, device=self._device)

    def Solve(self, cost=None, termination=None, ExtraArgs=None, strategy=None, CrossProbability=None, ScalingFactor=None,
              disp=False, **kwds):
        """every nested solver starts around its point (SetInitialPoints(x0, radius=0.05) of
        abstract_solver.py:471-506: members in x0*(1 +- 0.05), member 0 = x0), then all advance in
        lockstep until each has met the termination condition or its own limits"""
        nested = self._nested
        dev = _models.resolve(cost)
        if termination is not None:
            self._termination = termination
        strat = strategy if strategy is not None else nested.strategy
        (sid, k, _), _name = _strategy.strategy_id(strat)
        CR = nested.probability if CrossProbability is None else CrossProbability
        F = nested.scale if ScalingFactor is None else ScalingFactor
        NP, D, S = nested.nPop, self.nDim, self._npts
        if NP < k + 1:
            raise ValueError('Sample larger than population or is negative')
        self._np_each = NP
        if self._maxiter is None:
            self._maxiter = D * NP * 10          # abstract_solver.py:650-674
        if self._maxfun is None:
            self._maxfun = D * NP * 1000
        if self._init_solution is None:
            self._init_solution = self._InitialPoints()
        starts = np.asarray(self._init_solution, dtype=np.float64)
        radius = 0.05
        pop = np.empty((S * NP, D), dtype=np.float64)
        uniform = random.uniform
        for s in range(S):
            x0 = starts[s]
            rows = pop[s * NP:(s + 1) * NP]
            # abstract_solver.py:499-502: a zero coordinate gets [-radius, +radius] (x0*(1 +- radius) would pin
            # every member to 0 there and no difference vector could ever leave it), then :532-534
            # random.uniform(min[j], max[j]) member by member
            lo_ = x0 * (1 - radius)
            hi_ = x0 * (1 + radius)
            lo_[lo_ == 0] = -radius
            hi_[hi_ == 0] = radius
            for i in range(NP):
                for j in range(D):
                    rows[i, j] = uniform(lo_[j], hi_[j])
            rows[0] = x0
        eng = self._engine = self._make_engine(S * NP)
        mode = BOUNDS_NONE
        if len(self._strictMin):
            mode = BOUNDS_CLAMP if self._tight else BOUNDS_REJECT
            eng.set_bounds(mode, np.asarray(self._strictMin, float), np.asarray(self._strictMax, float))
        else:
            eng.set_bounds(BOUNDS_NONE)
        args = () if ExtraArgs is None else tuple(ExtraArgs)
        eng.set_cost(dev.device_id(False, D), dev.params(args))
        if self._penalty is _no_penalty:
            eng.set_penalty()
        else:
            eng.set_penalty(*self._penalty.device_terms(D))
        eng.set_strategy(sid, F, CR)
        if self._seed is None:
            self._seed = random.getrandbits(64)
        eng.set_rng_philox(self._seed)
        eng.upload_population(pop)
        if S > 1:
            eng.set_islands(S)     # S == 1 (e.g. LatticeSolver(dim >= 4) with the default nbins): one plain population

        self._members = [_Member(self, s) for s in range(S)]
        term = self._termination

        def absorb(m, be, ninf, bx):
            """one generation of nested solver m -> its stop message or None (differential_evolution.py:601, 617;
            abstract_solver.py:703-711)"""
            m._fcalls[0] += NP - int(ninf)
            m.bestEnergy = float(be)
            m.bestSolution = np.array(bx, dtype=np.float64)
            m.energy_history.append(m.bestEnergy)
            m.solution_history.append(m.bestSolution)
            if term(m):
                return str(term(m, info=True))
            if m.evaluations >= self._maxfun or m.generations >= self._maxiter:
                return 'EvaluationLimits with %s' % {'evaluations': self._maxfun, 'generations': self._maxiter}
            return None

        desc = _termination.device_descriptor(term) if self._batch_generations else None
        if desc is not None and S > 1 and hasattr(eng, 'run_islands'):
            # every nested solver's condition is tested on the device; the host replays the batch log
            eng.set_islands_termination(desc, self._maxiter, self._maxfun)
            active = set(range(S))
            while active:
                recs, done, stopped = eng.run_islands(self._batch_generations)
                progressed = False
                for s in sorted(active):
                    m = self._members[s]
                    for k in range(int(done[s])):
                        progressed = True
                        rec = recs[s, k]
                        msg = absorb(m, rec[0], rec[2], rec[4:4 + D])
                        if msg:
                            m.message = msg
                            break
                    if m.message:
                        active.discard(s)
                    elif stopped[s]:
                        raise RuntimeError('nested solver %d stopped on the device but not on the host' % s)
                if not progressed and active:
                    raise RuntimeError('batched nested solvers made no progress')
        else:
            active = list(range(S))
            first = True
            while active:
                if first:
                    eng.eval_initial()
                else:
                    eng.step()
                if S > 1:
                    be, bi, ninf, nacc, gen, bx = eng.read_islands()
                else:
                    r = eng.read_result()
                    be, bi, ninf, nacc, gen, bx = [r[0]], [r[1]], [r[2]], [r[3]], [r[4]], np.asarray(r[5]).reshape(1, D)
                still = []
                for s in active:
                    m = self._members[s]
                    msg = absorb(m, be[s], ninf[s], bx[s])
                    if msg:
                        m.message = msg
                    else:
                        still.append(s)
                active = still
                first = False
        # the nested solver with the lowest energy (abstract_ensemble_solver.py:465-478: `<=`, the later one wins ties)
        best = self._members[0]
        for m in self._members:
            if m.bestEnergy <= best.bestEnergy:
                best = m
        self._best = best
        if disp:
            print(self.Terminated(info=True))
        return

    # ---- results (abstract_ensemble_solver.py:146-168) ---------------------------------------
    bestSolution = property(lambda self: None if self._best is None else self._best.bestSolution)
    bestEnergy = property(lambda self: float('inf') if self._best is None else self._best.bestEnergy)
    generations = property(lambda self: 0 if self._best is None else self._best.generations)
    evaluations = property(lambda self: 0 if self._best is None else self._best.evaluations)
    energy_history = property(lambda self: [] if self._best is None else self._best.energy_history)
    solution_history = property(lambda self: [] if self._best is None else self._best.solution_history)
    _all_bestEnergy = property(lambda self: [m.bestEnergy for m in self._members])
    _all_bestSolution = property(lambda self: [m.bestSolution for m in self._members])
    _all_iters = property(lambda self: [m.generations for m in self._members])
    _all_evals = property(lambda self: [m.evaluations for m in self._members])
    _total_iters = property(lambda self: sum(m.generations for m in self._members))
    _total_evals = property(lambda self: sum(m.evaluations for m in self._members))

    def Solution(self):
        return self.bestSolution

    def Terminated(self, disp=False, info=False, termination=None, all=None):
        """the stop message of the best nested solver, or of all of them (all=True)"""
        if not self._members:
            return [] if all else ('' if info else False)
        if all:
            return [m.message if info else bool(m.message) for m in self._members]
        msg = self._best.message if self._best is not None else None
        if disp and msg:
            print(msg)
        return (msg or '') if info else bool(msg)


class BuckshotSolver(_EnsembleOfDE):
    """N nested solvers started from N uniform random points of the strict range (ensemble.py:84-121)"""

    def __init__(self, dim, npts=8, **kwds):
        super(BuckshotSolver, self).__init__(dim, npts, **kwds)

    def _InitialPoints(self):
        lower, upper = self._bounds()
        return [[random.uniform(lower[j], upper[j]) for j in range(self.nDim)] for _ in range(self._npts)]


class LatticeSolver(_EnsembleOfDE):
    """nested solvers started from the centres of a grid over the strict range (ensemble.py:44-82,
    mystic/math/grid.py:12-40); nbins is a tuple of bins per dimension (an int is spread over the
    dimensions so that the number of grid points does not exceed it)"""

    def __init__(self, dim, nbins=8, **kwds):
        if isinstance(nbins, int):
            per = max(1, int(round(nbins ** (1.0 / dim))))
            while per ** dim > nbins and per > 1:
                per -= 1
            nbins = (per,) * dim
        self._nbins = tuple(int(b) for b in nbins)
        if len(self._nbins) != dim:
            raise ValueError('nbins needs one entry per dimension')
        super(LatticeSolver, self).__init__(dim, int(np.prod(self._nbins)), **kwds)

    def _InitialPoints(self):
        lower, upper = self._bounds()
        axes = []
        for j, nb in enumerate(self._nbins):
            step = 1. * abs(upper[j] - lower[j]) / nb
            axes.append([lower[j] + (i + 0.5) * step for i in range(nb)])
        grid = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1).reshape(-1, self.nDim)
        return grid.tolist()
