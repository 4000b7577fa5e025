"""Termination conditions for solvers without a mystic installation.

Unmodified `mystic.termination` objects work on the B200 solver as they are: they only read
solver attributes (`energy_history`, `generations`, `_fcalls`, `population`, `popEnergy`,
`bestSolution`; mystic/termination.py:181-408).  This module provides the same factories
(same names, arguments and `__doc__` convention "<Name> with {kwds}") so the engine also runs
where mystic is absent, e.g. on the GPU box.  Each condition is `cond(solver, info=False)`,
returning bool, or the describing string / '' when info is true.
"""
import time as _time

import numpy as np

inf = float('inf')
EARLYEXIT = 0
_type = type


class _Condition(object):
    """A termination condition: `cond(solver, info=False)`.  A class (not a closure) so that solvers
    holding one can be pickled by SaveSolver without dill."""

    def __init__(self, name, kwds):
        self.name = name
        self.kwds = dict(kwds)
        self.__doc__ = '%s with %s' % (name, kwds)
        self.__name__ = '_' + name

    def __call__(self, inst, info=False):
        met = _TESTS[self.name](inst, **self.kwds)
        if info:
            return self.__doc__ if met else ''
        return bool(met)


def _lookback(generations):
    return 0 if generations is None else int(generations)


def _vtr(inst, tolerance, target):
    hist = inst.energy_history
    return bool(len(hist)) and abs(hist[-1] - target) <= tolerance


def _cog(inst, tolerance, generations):
    hist = inst.energy_history
    g = _lookback(generations)
    if not len(hist) or len(hist) <= g:
        return False
    return (hist[-g] - hist[-1]) <= tolerance or hist[-g] == hist[-1]


def _ncog(inst, tolerance, generations):
    hist = inst.energy_history
    g = _lookback(generations)
    if not len(hist) or len(hist) <= g:
        return False
    if hist[-g] == hist[-1]:
        return True
    return 2.0 * (hist[-g] - hist[-1]) <= tolerance * (abs(hist[-g]) + abs(hist[-1])) + 1e-20


def _vtrcog(inst, ftol, gtol, generations, target):
    hist = inst.energy_history
    if not len(hist):
        return False
    g = _lookback(generations)
    if len(hist) > g and ((hist[-g] - hist[-1]) <= gtol or hist[-g] == hist[-1]):
        return True
    return abs(hist[-1] - target) <= ftol


def _nct(inst, fval, tolerance, generations):
    hist = inst.energy_history
    if not len(hist):
        return False
    g = _lookback(generations)
    if g and fval is None:
        return len(hist) > g and ((hist[-g] - hist[-1]) <= 0 or hist[-g] == hist[-1])
    if not g and fval is None:
        return True
    return abs(hist[-1] - fval) <= abs(tolerance * fval)


def _crt(inst, xtol, ftol):
    sim = np.array(inst.population)
    fsim = np.array(inst.popEnergy)
    if len(fsim) < 2:
        return False
    with np.errstate(invalid='ignore'):
        return bool(np.max(np.abs(sim[1:] - sim[0])) <= xtol and np.max(np.abs(fsim[0] - fsim[1:])) <= ftol)


def _spread(inst, tolerance):
    sim = np.array(inst.population)
    return bool(np.all(np.abs(sim - sim[0]) <= np.abs(tolerance * sim[0])))


def _improvement(inst, tolerance):
    """termination.py:275-285: the largest sum over the parameters of |bestSolution - trial| over the trial
    population (needs the trial vectors: construct the solver with keep_trials=True)"""
    best = np.array(inst.bestSolution)
    trial = np.array(inst.trialSolution)
    answer = np.add.reduce(abs(best - trial).T)
    if isinstance(answer, np.ndarray):
        answer = max(answer)
    return bool(answer <= tolerance)


_EPSILON = float(np.sqrt(np.finfo(float).eps))   # mystic/_scipy060optimize.py:47


def _gradnorm(inst, tolerance, norm):
    """termination.py:369-382: forward-difference gradient of the raw cost at bestSolution
    (_scipy060optimize.py:611-619: D + 1 cost evaluations per test), Lnorm of math/distance.py:24-39"""
    grad = getattr(inst, 'gradient', [None])[-1]
    if grad is None:
        soln = np.asarray(inst.bestSolution, dtype=np.float64)
        cost = inst._cost[1]
        pts = np.vstack([soln] + [soln + _EPSILON * np.eye(len(soln))[k] for k in range(len(soln))])
        try:
            f = np.asarray(cost(pts), dtype=np.float64)          # device costs take a batch of rows
            if f.shape != (len(soln) + 1,):
                raise TypeError
        except TypeError:
            f = np.array([cost(x) for x in pts], dtype=np.float64)
        grad = (f[1:] - f[0]) / _EPSILON
    w = np.asarray(grad, dtype=np.float64)
    if not norm:
        gnorm = np.sum(w != 0.0, dtype=float)
    elif norm == inf:
        gnorm = np.max(np.abs(w))
    elif norm == -inf:
        gnorm = np.min(np.abs(w))
    else:
        with np.errstate(over='raise', invalid='raise'):
            try:
                gnorm = np.sum(np.abs(w ** norm)) ** (1. / norm)
            except FloatingPointError:
                gnorm = np.max(np.abs(w))
    return bool(gnorm <= tolerance)


def _limits(inst, generations, evaluations):
    maxiter = inf if generations is None else generations
    maxfun = inf if evaluations is None else evaluations
    return inst._fcalls[0] >= maxfun or inst.generations >= maxiter


def _interrupt(inst):
    return bool(inst._EARLYEXIT)


_TESTS = {'VTR': _vtr, 'ChangeOverGeneration': _cog, 'NormalizedChangeOverGeneration': _ncog,
          'VTRChangeOverGeneration': _vtrcog, 'NormalizedCostTarget': _nct, 'CandidateRelativeTolerance': _crt,
          'PopulationSpread': _spread, 'EvaluationLimits': _limits, 'SolverInterrupt': _interrupt,
          'SolutionImprovement': _improvement, 'GradientNormTolerance': _gradnorm}


def VTR(tolerance=0.005, target=0.0):
    """abs(cost[-1] - target) <= tolerance   (mystic/termination.py:181-196)"""
    return _Condition('VTR', {'tolerance': tolerance, 'target': target})


def ChangeOverGeneration(tolerance=1e-6, generations=30):
    """cost[-g] - cost[-1] <= tolerance   (mystic/termination.py:198-217)"""
    return _Condition('ChangeOverGeneration', {'tolerance': tolerance, 'generations': generations})


def NormalizedChangeOverGeneration(tolerance=1e-4, generations=10):
    """2*(cost[-g]-cost[-1]) <= tolerance*(|cost[-g]|+|cost[-1]|) + 1e-20   (termination.py:219-240)"""
    return _Condition('NormalizedChangeOverGeneration', {'tolerance': tolerance, 'generations': generations})


def VTRChangeOverGeneration(ftol=0.005, gtol=1e-6, generations=30, target=0.0):
    """cost[-g] - cost[-1] <= gtol or abs(cost[-1] - target) <= ftol   (termination.py:320-342)"""
    return _Condition('VTRChangeOverGeneration',
                      {'ftol': ftol, 'gtol': gtol, 'generations': generations, 'target': target})


def NormalizedCostTarget(fval=None, tolerance=1e-6, generations=30):
    """abs(cost[-1]-fval) <= abs(tolerance*fval), or no improvement over g iterations (termination.py:290-318)"""
    return _Condition('NormalizedCostTarget', {'fval': fval, 'tolerance': tolerance, 'generations': generations})


def CandidateRelativeTolerance(xtol=1e-4, ftol=1e-4):
    """abs(xi-x0) <= xtol & abs(fi-f0) <= ftol; reads the whole population (termination.py:242-267)"""
    return _Condition('CandidateRelativeTolerance', {'xtol': xtol, 'ftol': ftol})


def PopulationSpread(tolerance=1e-6):
    """abs(params - params[0]) <= abs(tolerance*params[0]); reads the whole population (termination.py:344-361)"""
    return _Condition('PopulationSpread', {'tolerance': tolerance})


def SolutionImprovement(tolerance=1e-5):
    """sum(abs(last_params - current_params)) <= tolerance; reads the trial population (termination.py:269-288)"""
    return _Condition('SolutionImprovement', {'tolerance': tolerance})


def GradientNormTolerance(tolerance=1e-5, norm=inf):
    """sum(abs(gradient)**norm)**(1.0/norm) <= tolerance, forward-difference gradient (termination.py:363-384)"""
    return _Condition('GradientNormTolerance', {'tolerance': tolerance, 'norm': norm})


def EvaluationLimits(generations=None, evaluations=None):
    """iterations >= generations or fcalls >= evaluations   (termination.py:386-408)"""
    return _Condition('EvaluationLimits', {'generations': generations, 'evaluations': evaluations})


class _TimeLimits(_Condition):
    def __init__(self, seconds, system):
        _Condition.__init__(self, 'TimeLimits', {'seconds': seconds, 'system': system})
        self._delta = inf if seconds is None else getattr(seconds, 'total_seconds', lambda: abs(seconds))()
        self.reset()

    def _timer(self):
        system = self.kwds['system']
        return _time.time() if system is None else (_time.perf_counter() if system else _time.process_time())

    def reset(self):
        self._start = self._timer()

    def __call__(self, inst, info=False):
        met = (self._timer() - self._start) >= self._delta
        if info:
            return self.__doc__ if met else ''
        return bool(met)


def TimeLimits(seconds=86400, system=None):
    """elapsed time >= seconds   (termination.py:410-437)"""
    return _TimeLimits(seconds, system)


def SolverInterrupt():
    """_EARLYEXIT == True   (termination.py:439-451)"""
    return _Condition('SolverInterrupt', {})


class When(tuple):
    """terminate when the given condition is satisfied (termination.py:63-108)"""

    def __new__(cls, arg):
        if isinstance(arg, tuple) and len(arg) == 1:
            arg = arg[0]
        if not getattr(arg, '__len__', None):
            arg = [arg]
        return tuple.__new__(cls, arg)

    def _results(self, solver, info):
        return dict((f, f(solver, info)) for f in self)

    def _met(self, values):
        return all(values)

    def _satisfied(self, stop, met):
        return stop if met else {}

    def __call__(self, solver, info=False):
        if info == 'not':
            return tuple(set(f for f in self if f not in self(solver, 'self')))
        stop = self._results(solver, info)
        met = self._met(stop.values())
        if not info:
            return met
        satisfied = self._satisfied(stop, met)
        if info == 'self':
            return tuple(set(satisfied.keys()))
        if not satisfied:
            return ''
        return '; '.join(sorted(set('; '.join(str(v) for v in satisfied.values()).split('; '))))

    def __repr__(self):
        return 'When(%s)' % str(self[0])


class And(When):
    """terminate when all given conditions are satisfied (termination.py:110-131)"""

    def __new__(cls, *args):
        if len(args) == 1:
            args = args[0]
        if not getattr(args, '__len__', None):
            args = [args]
        return tuple.__new__(cls, args)

    def __repr__(self):
        return 'And%s' % str(tuple(f for f in self))


class Or(When):
    """terminate when any of the given conditions is satisfied (termination.py:134-176)"""

    def __new__(cls, *args):
        if len(args) == 1:
            args = args[0]
        if not getattr(args, '__len__', None):
            args = [args]
        return tuple.__new__(cls, args)

    def _met(self, values):
        return any(values)

    def _satisfied(self, stop, met):
        return dict((k, v) for k, v in stop.items() if v)

    def __repr__(self):
        return 'Or%s' % str(tuple(f for f in self))


def type(condition):
    """get object that generated the given termination instance (termination.py:29-39)"""
    if isinstance(condition, _Condition):
        return globals()[condition.name]
    return _type(condition)


def state(condition):
    """{doc: kwds} of every leaf condition (termination.py:30-47)"""
    out = {}
    for term in (condition if isinstance(condition, tuple) else (condition,)):
        doc = term.__doc__
        if doc is None:
            continue
        if isinstance(term, tuple):
            out.update(state(term))
        elif ' with ' in doc:
            out[doc] = eval(doc.split(' with ', 1)[1], {'inf': inf, 'nan': float('nan')})
    return out


_DEVICE_KINDS = {'VTR': 1, 'ChangeOverGeneration': 2, 'NormalizedChangeOverGeneration': 3,
                 'VTRChangeOverGeneration': 4, 'EvaluationLimits': 5}


def device_descriptor(condition):
    """Describe a condition for the device-side check of the batched generation loop
    (mb2_set_termination), or None when it cannot run there (compound conditions, conditions that
    read the population, time limits, user callables).  Works for these factories and for the
    unmodified mystic.termination ones alike: both carry "<Name> with {kwds}" in __doc__."""
    if isinstance(condition, tuple):
        return None
    doc = getattr(condition, '__doc__', None)
    if not doc or ' with ' not in doc:
        return None
    name, kw = doc.split(' with ', 1)
    if name not in _DEVICE_KINDS:
        return None
    try:
        kw = eval(kw, {'inf': inf, 'nan': float('nan')})
    except Exception:
        return None
    d = dict(kind=_DEVICE_KINDS[name], lookback=0, tolerance=0.0, target=0.0, gtol=0.0,
             term_generations=-1, term_evaluations=-1)
    if name == 'VTR':
        d.update(tolerance=kw['tolerance'], target=kw['target'])
    elif name in ('ChangeOverGeneration', 'NormalizedChangeOverGeneration'):
        d.update(tolerance=kw['tolerance'], lookback=_lookback(kw['generations']))
    elif name == 'VTRChangeOverGeneration':
        d.update(tolerance=kw['ftol'], gtol=kw['gtol'], target=kw['target'], lookback=_lookback(kw['generations']))
    else:
        g, e = kw['generations'], kw['evaluations']
        d.update(term_generations=-1 if g is None or g == inf else int(g),
                 term_evaluations=-1 if e is None or e == inf else int(e))
    if not (0 <= d['lookback'] < 1024):
        return None
    return d
