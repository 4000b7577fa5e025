#!/usr/bin/env python
"""Record golden vectors from the UNMODIFIED reference (mystic @ /root/reference).

TEST INFRASTRUCTURE ONLY.  Runs in the build container (the reference cannot travel
to the GPU box); writes small .npz fixtures into tests/golden/.  Usage:

    python oracle/record_reference.py            # all cases
    python oracle/record_reference.py c1_rosen3  # one case

How the recording works, without touching reference code:
  * mystic/strategy.py:19 does `import random` and calls random.sample /
    random.randrange / random.random (strategy.py:27,42,49,...).  We set
    `mystic.strategy.random = Recorder()`, an object with those three methods that
    delegates to the stdlib generator and logs every value returned.  Per candidate
    the draw order is sample(k) -> randrange(D) -> uniforms (Best1Bin: randrange
    comes after the trial copy, but still before the uniforms).
  * trial vectors / trial energies are captured by a recording `map` installed with
    SetMapper (the only place the reference evaluates the cost,
    mystic/differential_evolution.py:589).
  * per-generation solver state is snapshotted from the `callback` hook
    (differential_evolution.py:622).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, '_shim'))
sys.path.insert(1, '/root/reference')
sys.path.insert(2, REPO)

import random as _stdlib_random  # noqa: E402
import numpy as np  # noqa: E402

KMAX = 5


class Recorder(object):
    """Stand-in for the `random` module seen by mystic.strategy."""

    def __init__(self):
        self.cands = []  # one dict per strategy call

    def sample(self, population, k):
        out = _stdlib_random.sample(population, k)
        self.cands.append({'donors': list(out), 'n': None, 'u': []})
        return out

    def randrange(self, *args):
        out = _stdlib_random.randrange(*args)
        self.cands[-1]['n'] = out
        return out

    def random(self):
        out = _stdlib_random.random()
        self.cands[-1]['u'].append(out)
        return out


class RecordingMap(object):
    """indexable-result map (differential_evolution.py:592 needs trialEnergy[0])."""

    def __init__(self):
        self.trials = []
        self.energies = []

    def __call__(self, f, seq, **kwds):
        self.trials.append(np.array([list(map(float, row)) for row in seq], dtype=np.float64))
        out = [f(row) for row in seq]
        self.energies.append(np.array(out, dtype=np.float64))
        return out


def _fit_points(spec):
    """('arange', lo, hi, step) | ('linspace', lo, hi, M) | explicit list -> float64 array"""
    if isinstance(spec, (list, tuple)) and spec and spec[0] == 'arange':
        return np.arange(spec[1], spec[2], spec[3], dtype=np.float64)
    if isinstance(spec, (list, tuple)) and spec and spec[0] == 'linspace':
        return np.linspace(spec[1], spec[2], int(spec[3]))
    return np.asarray(spec, dtype=np.float64)


def _fit_cost(case):
    """data-fit residual costs built by the REFERENCE's own factories:
    mystic.models.poly.CostFactory(target, pts) / mystic.models.lorentzian.CostFactory(target, pts)
    (abstract_model.py:146-153 -> forward_model.py:204-260).  Returns (cost, pts, observations)."""
    import mystic.models as mm
    if case['cost'] == 'decay':
        # mystic.models.br8: Bevington's dual exponential decay; `data` = the book's counts, `cost` the
        # module's own CostFactory2(data[:,0], data[:,1], 5) with sigma = sqrt(counts) (br8.py:51-56, 92)
        import importlib
        br8 = importlib.import_module('mystic.models.br8')
        pts = np.asarray(br8.data[:, 0], dtype=np.float64)
        obs = np.asarray(br8.data[:, 1], dtype=np.float64)
        return br8.decay.CostFactory2(br8.data[:, 0], br8.data[:, 1], 5), pts, obs, np.sqrt(br8.data[:, 1])
    model = mm.poly if case['cost'] == 'polyfit' else mm.lorentzian
    pts = _fit_points(case['fit']['pts'])
    target = [float(v) for v in case['fit']['target']]
    cost = model.CostFactory(target, pts)
    obs = np.asarray(model.evaluate(target, pts), dtype=np.float64)
    return cost, pts, obs


def _cost(name):
    import importlib
    import mystic.models as mm
    if hasattr(mm, name):
        return getattr(mm, name)
    # `mystic.models.poly` the attribute is a Polynomial() instance (models/__init__.py),
    # so fetch the module itself from sys.modules
    importlib.import_module('mystic.models.poly')
    return getattr(sys.modules['mystic.models.poly'], name)


def _termination(spec):
    import mystic.termination as mt
    if spec is None:
        return None
    kind = spec[0]
    return getattr(mt, kind)(*spec[1:])


def record(case):
    import mystic.strategy as mstrategy
    from mystic.solvers import DifferentialEvolutionSolver2
    from mystic.tools import random_seed
    from oracle.host_callables import build_reference_penalty

    dim, NP = case['dim'], case['NP']
    random_seed(case['seed'])
    # NOTE (reference quirk, differential_evolution.py:676-692 + abstract_solver.py:1171):
    # CrossProbability/ScalingFactor given to Solve() are overwritten by the stale
    # 'cross'/'scale' entries of `settings` on the first Step, so they only take effect when
    # given to the constructor (or to Step directly).  via='solve' records that quirk
    # (BASELINE config 1 as literally specified); via='ctor' really runs with CR/F.
    via = case.get('via', 'ctor')
    if via == 'ctor':
        solver = DifferentialEvolutionSolver2(dim, NP, CrossProbability=case['CR'], ScalingFactor=case['F'])
    else:
        solver = DifferentialEvolutionSolver2(dim, NP)
    NP = solver.nPop
    init = case['init']
    if init[0] == 'random':
        solver.SetRandomInitialPoints(list(init[1]), list(init[2]))
    else:
        solver.SetInitialPoints(list(init[1]), init[2])
    bounds = case.get('bounds')
    if bounds is not None:
        lo, hi, tight = bounds
        if tight:
            solver.SetStrictRanges(list(lo), list(hi), tight=True)
        else:
            solver.SetStrictRanges(list(lo), list(hi))
    pen = case.get('penalty')
    if pen:
        solver.SetPenalty(build_reference_penalty(pen))
    solver.SetEvaluationLimits(generations=case['maxgen'])
    rmap = RecordingMap()
    solver.SetMapper(rmap)

    pop0 = np.array([list(map(float, r)) for r in solver.population], dtype=np.float64)

    rec = Recorder()
    snaps = []

    def callback(best):
        snaps.append({
            'pop': np.array([list(map(float, r)) for r in solver.population], dtype=np.float64),
            'popE': np.array(solver.popEnergy, dtype=np.float64),
            'bestE': float(solver.bestEnergy),
            'bestX': np.array(solver.bestSolution, dtype=np.float64).copy(),
            'evals': int(solver.evaluations),
        })

    # Mersenne-Twister state right before Solve(): building the symbolic bounds constraint in
    # SetStrictRanges(tight=True) consumes stdlib draws inside mystic.symbolic.simplify
    # (symbolic.py:575-812), so seed + setter order alone does not determine it.
    mt_state = np.array(_stdlib_random.getstate()[1], dtype=np.uint32)
    saved = mstrategy.random
    mstrategy.random = rec
    try:
        strategy = getattr(mstrategy, case['strategy'])
        term = _termination(case.get('termination'))
        if term is None:
            import mystic.termination as mt
            term = mt.VTR(-1.0)  # never satisfied for the non-negative costs used here
        extra = tuple(case['ExtraArgs']) if case.get('ExtraArgs') else None
        fit = None
        if case['cost'] in ('polyfit', 'lorentzian', 'decay'):
            fit = _fit_cost(case)
        the_cost = fit[0] if fit else _cost(case['cost'])
        if via == 'ctor':
            solver.Solve(the_cost, term, ExtraArgs=extra, strategy=strategy,
                         callback=callback)
        else:
            solver.Solve(the_cost, term, ExtraArgs=extra, strategy=strategy,
                         CrossProbability=case['CR'], ScalingFactor=case['F'],
                         callback=callback)
    finally:
        mstrategy.random = saved

    G = len(snaps) - 1  # generations after generation 0
    assert len(rmap.trials) == G + 1, (len(rmap.trials), G)
    assert len(rec.cands) == G * NP, (len(rec.cands), G, NP)
    assert solver.generations == G

    donors = -np.ones((G, NP, KMAX), dtype=np.int32)
    nstart = np.zeros((G, NP), dtype=np.int32)
    nuni = np.zeros((G, NP), dtype=np.int32)
    uni = np.full((G, NP, dim + 1), np.nan, dtype=np.float64)
    for g in range(G):
        for c in range(NP):
            e = rec.cands[g * NP + c]
            donors[g, c, :len(e['donors'])] = e['donors']
            nstart[g, c] = e['n']
            nuni[g, c] = len(e['u'])
            uni[g, c, :len(e['u'])] = e['u']

    out = {
        'meta': json.dumps(dict(case, NP=NP, generations=G, evaluations=int(solver.evaluations),
                                via=via, CR_eff=float(solver.probability), F_eff=float(solver.scale),
                                stop=str(solver.Terminated(info=True)),
                                python=sys.version.split()[0], numpy=np.__version__)),
        'pop0': pop0,
        'trial': np.stack(rmap.trials),          # [G+1, NP, D]  (index 0 = generation 0)
        'trialE': np.stack(rmap.energies),       # [G+1, NP]
        'pop': np.stack([s['pop'] for s in snaps]),    # after selection, per generation
        'popE': np.stack([s['popE'] for s in snaps]),
        'bestE': np.array([s['bestE'] for s in snaps]),
        'bestX': np.stack([s['bestX'] for s in snaps]),
        'evals': np.array([s['evals'] for s in snaps], dtype=np.int64),
        'energy_history': np.array(solver.energy_history, dtype=np.float64),
        'donors': donors, 'nstart': nstart, 'nuni': nuni, 'uni': uni, 'mt_state': mt_state,
    }
    if fit:
        out['fit_pts'], out['fit_obs'] = fit[1], fit[2]
        if len(fit) > 3:
            out['fit_sigma'] = fit[3]
    return out


ALL_STRATEGIES = ['Best1Exp', 'Best1Bin', 'Rand1Exp', 'RandToBest1Exp', 'Best2Exp', 'Rand2Exp',
                  'Rand1Bin', 'RandToBest1Bin', 'Best2Bin', 'Rand2Bin']


def cases():
    cs = {}
    # BASELINE config 1 as literally specified (SURVEY.md 8c golden vector)
    cs['c1_rosen3_best1exp'] = dict(dim=3, NP=40, strategy='Best1Exp', cost='rosen', CR=0.5, F=0.6,
                                    seed=123, init=('random', [0.] * 3, [2.] * 3), maxgen=99999,
                                    termination=('VTR', 1e-4), via='solve')
    # same problem with CR/F really in effect (given to the constructor)
    cs['c1b_rosen3_best1exp_ctor'] = dict(dim=3, NP=40, strategy='Best1Exp', cost='rosen', CR=0.5, F=0.6,
                                          seed=123, init=('random', [0.] * 3, [2.] * 3), maxgen=99999,
                                          termination=('VTR', 1e-4), via='ctor')
    # every strategy name, unconstrained rosen
    for i, s in enumerate(ALL_STRATEGIES):
        cs['strat_%s_rosen8' % s] = dict(dim=8, NP=24, strategy=s, cost='rosen', CR=0.9, F=0.8,
                                         seed=1000 + i, init=('random', [-2.] * 8, [2.] * 8), maxgen=10)
    # bounds mode A: default SetStrictRanges -> out-of-bounds trials get inf (tools.py:420-426)
    cs['boundsA_best1bin_rosen16'] = dict(dim=16, NP=32, strategy='Best1Bin', cost='rosen', CR=0.9, F=0.8,
                                          seed=7, init=('random', [-1.5] * 16, [1.5] * 16),
                                          bounds=([-1.5] * 16, [1.5] * 16, False), maxgen=12)
    # bounds mode B: tight=True -> symbolic clamp (symbolic.py:1139-1144)
    cs['boundsB_best1bin_rosen16'] = dict(dim=16, NP=32, strategy='Best1Bin', cost='rosen', CR=0.9, F=0.8,
                                          seed=7, init=('random', [-1.5] * 16, [1.5] * 16),
                                          bounds=([-1.5] * 16, [1.5] * 16, True), maxgen=12)
    cs['boundsB_rand1exp_rosen16'] = dict(dim=16, NP=32, strategy='Rand1Exp', cost='rosen', CR=0.9, F=0.8,
                                          seed=8, init=('random', [-1.5] * 16, [1.5] * 16),
                                          bounds=([-1.5] * 16, [1.5] * 16, True), maxgen=12)
    # population initialised outside the strict range: generation-0 clamp
    # (differential_evolution.py:516-520)
    cs['boundsA_init_outside_rosen4'] = dict(dim=4, NP=16, strategy='Rand1Bin', cost='rosen', CR=0.9, F=0.8,
                                             seed=11, init=('random', [-4.] * 4, [4.] * 4),
                                             bounds=([-2.] * 4, [2.] * 4, False), maxgen=8)
    # corana: reference uses only x[0..3] (storn.py:87-88); clamp bounds + penalty (BASELINE config 4 shape)
    cs['corana8_rand1exp_pen'] = dict(dim=8, NP=32, strategy='Rand1Exp', cost='corana', CR=0.9, F=0.8,
                                      seed=21, init=('random', [-1000.] * 8, [1000.] * 8),
                                      bounds=([-1000.] * 8, [1000.] * 8, True),
                                      penalty=[dict(ptype='quadratic_inequality', G=[1.] * 8, h=100., k=100, hpow=5)],
                                      maxgen=12)
    cs['corana4_rand1bin'] = dict(dim=4, NP=40, strategy='Rand1Bin', cost='corana', CR=0.9, F=0.8,
                                  seed=123, init=('random', [-1000.] * 4, [1000.] * 4),
                                  bounds=([-1000.] * 4, [1000.] * 4, False), maxgen=25)
    cs['corana2_best1exp'] = dict(dim=2, NP=12, strategy='Best1Exp', cost='corana', CR=0.7, F=0.8,
                                  seed=22, init=('random', [-10.] * 2, [10.] * 2), maxgen=10)
    # griewangk
    cs['griewangk10_best1bin'] = dict(dim=10, NP=40, strategy='Best1Bin', cost='griewangk', CR=0.9, F=0.8,
                                      seed=31, init=('random', [-400.] * 10, [400.] * 10),
                                      bounds=([-400.] * 10, [400.] * 10, True), maxgen=15)
    cs['griewangk10_rand1exp'] = dict(dim=10, NP=50, strategy='Rand1Exp', cost='griewangk', CR=0.3, F=1.0,
                                      seed=32, init=('random', [-400.] * 10, [400.] * 10),
                                      bounds=([-400.] * 10, [400.] * 10, False), maxgen=15)
    # zimmermann (2-D only)
    cs['zimmermann_rand1bin'] = dict(dim=2, NP=40, strategy='Rand1Bin', cost='zimmermann', CR=0.9, F=0.8,
                                     seed=321, init=('random', [0.] * 2, [5.] * 2),
                                     bounds=([0.] * 2, [5.] * 2, False), maxgen=20)
    cs['zimmermann_best2bin'] = dict(dim=2, NP=20, strategy='Best2Bin', cost='zimmermann', CR=0.9, F=0.8,
                                     seed=322, init=('random', [-5.] * 2, [10.] * 2), maxgen=15)
    # chebyshev polynomial fit (examples/test_ffit.py:72-81: Best1Exp CR=1.0 F=0.9, init +-100)
    cs['cheb8_best1exp'] = dict(dim=9, NP=36, strategy='Best1Exp', cost='chebyshev8cost', CR=1.0, F=0.9,
                                seed=41, init=('random', [-100.] * 9, [100.] * 9), maxgen=12)
    cs['cheb8_M128_best1bin'] = dict(dim=9, NP=36, strategy='Best1Bin', cost='chebyshev8cost', CR=0.9, F=0.8,
                                     seed=42, init=('random', [-100.] * 9, [100.] * 9), maxgen=8,
                                     ExtraArgs=[128])
    cs['cheb4_rand2exp'] = dict(dim=5, NP=20, strategy='Rand2Exp', cost='chebyshev4cost', CR=0.8, F=0.7,
                                seed=43, init=('random', [-10.] * 5, [10.] * 5), maxgen=8)
    # stacked penalties of all four device-expressible kinds
    cs['rosen6_stacked_pen'] = dict(dim=6, NP=24, strategy='RandToBest1Exp', cost='rosen', CR=0.9, F=0.8,
                                    seed=51, init=('random', [-2.] * 6, [2.] * 6),
                                    penalty=[dict(ptype='quadratic_inequality', G=[1., 1., 1., 1., 1., 1.], h=2., k=100, hpow=5),
                                             dict(ptype='linear_equality', G=[1., -1., 0., 0., 0., 0.], h=0., k=10, hpow=5),
                                             dict(ptype='quadratic_equality', G=[0., 0., 1., 0., 0., -2.], h=0.5, k=50, hpow=5),
                                             dict(ptype='linear_inequality', G=[0., 0., 0., 1., 1., 0.], h=1., k=100, hpow=5)],
                                    maxgen=10)
    # the other penalty kinds over linear conditions (penalty.py:159-339, 459-601); the lagrange kinds after two
    # penalty iterations with stored condition values
    cs['pen_uniform_rosen4'] = dict(dim=4, NP=20, strategy='Best1Exp', cost='rosen', CR=0.9, F=0.8, seed=501,
                                    init=('random', [-2.] * 4, [2.] * 4),
                                    penalty=[dict(ptype='uniform_inequality', G=[1., 1., 1., 1.], h=3., k=1e3, hpow=5),
                                             dict(ptype='uniform_equality', G=[1., -1., 0., 0.], h=0., k=0.5, hpow=5)],
                                    maxgen=10)
    cs['pen_barrier_rosen4'] = dict(dim=4, NP=20, strategy='Rand1Exp', cost='rosen', CR=0.9, F=0.8, seed=502,
                                    init=('random', [-1.] * 4, [1.] * 4),
                                    penalty=[dict(ptype='barrier_inequality', G=[1., 1., 1., 1.], h=3., k=100, hpow=5)],
                                    maxgen=10)
    cs['pen_lagrange_rosen5'] = dict(dim=5, NP=25, strategy='RandToBest1Exp', cost='rosen', CR=0.9, F=0.8, seed=503,
                                     init=('random', [-2.] * 5, [2.] * 5),
                                     penalty=[dict(ptype='lagrange_inequality', G=[1., 1., 1., 1., 1.], h=2., k=20, hpow=5, n=2,
                                                   store_points=[[1., 1., 1., 0., 0.], [0.5, 0.5, 0.5, 0.5, 0.25]]),
                                              dict(ptype='lagrange_equality', G=[1., -2., 0., 0., 1.], h=0.25, k=20, hpow=5, n=2,
                                                   store_points=[[1., 0., 0., 0., 0.], [0.5, 0.2, 0., 0., 0.1]])],
                                     maxgen=10)
    cs['pen_lagrange_corana64_rand1exp'] = dict(dim=64, NP=64, strategy='Rand1Exp', cost='corana', CR=0.9, F=0.8, seed=504,
                                                init=('random', [-1000.] * 64, [1000.] * 64),
                                                bounds=([-1000.] * 64, [1000.] * 64, True),
                                                penalty=[dict(ptype='lagrange_inequality', G=[1.] * 64, h=100., k=20, hpow=5, n=1,
                                                              store_points=[[10.] * 64])],
                                                maxgen=4)
    # edge cases: CR extremes, SetInitialPoints, NP raised to max(NP, dim, 4)
    cs['edge_cr0_best1bin'] = dict(dim=5, NP=12, strategy='Best1Bin', cost='rosen', CR=0.0, F=0.8,
                                   seed=61, init=('random', [-2.] * 5, [2.] * 5), maxgen=6)
    cs['edge_cr0_rand1exp'] = dict(dim=5, NP=12, strategy='Rand1Exp', cost='rosen', CR=0.0, F=0.8,
                                   seed=62, init=('random', [-2.] * 5, [2.] * 5), maxgen=6)
    cs['edge_cr1_best2exp'] = dict(dim=5, NP=12, strategy='Best2Exp', cost='rosen', CR=1.0, F=0.5,
                                   seed=63, init=('random', [-2.] * 5, [2.] * 5), maxgen=6)
    cs['edge_x0_rosen2'] = dict(dim=2, NP=40, strategy='Rand1Bin', cost='rosen', CR=0.9, F=0.8,
                                seed=123, init=('x0', [2., 3.], 0.05),
                                bounds=([-5.] * 2, [5.] * 2, False), maxgen=30)
    cs['edge_np_raised'] = dict(dim=7, NP=2, strategy='Rand2Bin', cost='rosen', CR=0.9, F=0.8,
                                seed=64, init=('random', [-2.] * 7, [2.] * 7), maxgen=6)
    cs['edge_dim1'] = dict(dim=1, NP=8, strategy='Best1Bin', cost='griewangk', CR=0.9, F=0.8,
                           seed=65, init=('random', [-10.], [10.]), maxgen=6)
    # wide rows (exercise multi-chunk row handling in the kernels): D not a multiple of 32/64
    cs['wide_rosen100_best1bin'] = dict(dim=100, NP=100, strategy='Best1Bin', cost='rosen', CR=0.9, F=0.8,
                                        seed=71, init=('random', [-5.] * 100, [5.] * 100),
                                        bounds=([-5.] * 100, [5.] * 100, True), maxgen=4)
    cs['wide_griewangk70_randtobest'] = dict(dim=70, NP=72, strategy='RandToBest1Bin', cost='griewangk', CR=0.9,
                                             F=0.8, seed=72, init=('random', [-400.] * 70, [400.] * 70),
                                             bounds=([-400.] * 70, [400.] * 70, False), maxgen=4)
    cs['wide_corana64_rand1exp_pen'] = dict(dim=64, NP=64, strategy='Rand1Exp', cost='corana', CR=0.9, F=0.8,
                                            seed=73, init=('random', [-1000.] * 64, [1000.] * 64),
                                            bounds=([-1000.] * 64, [1000.] * 64, True),
                                            penalty=[dict(ptype='quadratic_inequality', G=[1.] * 64, h=100., k=100, hpow=5)],
                                            maxgen=4)
    # separable test functions of mystic.models
    cs['sep_sphere_best1exp'] = dict(dim=12, NP=30, strategy='Best1Exp', cost='sphere', CR=0.9, F=0.8, seed=401,
                                     init=('random', [-5.] * 12, [5.] * 12), maxgen=10)
    cs['sep_rastrigin_rand1bin'] = dict(dim=10, NP=40, strategy='Rand1Bin', cost='rastrigin', CR=0.9, F=0.8, seed=402,
                                        init=('random', [-5.12] * 10, [5.12] * 10),
                                        bounds=([-5.12] * 10, [5.12] * 10, True), maxgen=10)
    cs['sep_schwefel_randtobest'] = dict(dim=8, NP=32, strategy='RandToBest1Exp', cost='schwefel', CR=0.9, F=0.8, seed=403,
                                         init=('random', [-500.] * 8, [500.] * 8),
                                         bounds=([-500.] * 8, [500.] * 8, False), maxgen=10)
    cs['sep_ackley_best2exp'] = dict(dim=16, NP=36, strategy='Best2Exp', cost='ackley', CR=0.9, F=0.6, seed=404,
                                     init=('random', [-10.] * 16, [10.] * 16), maxgen=8)
    cs['sep_step_best1bin'] = dict(dim=9, NP=30, strategy='Best1Bin', cost='step', CR=0.9, F=0.8, seed=405,
                                   init=('random', [-7.] * 9, [7.] * 9), maxgen=8)
    cs['sep_michal_rand2exp'] = dict(dim=6, NP=30, strategy='Rand2Exp', cost='michal', CR=0.9, F=0.7, seed=406,
                                     init=('random', [0.] * 6, [3.14] * 6), bounds=([0.] * 6, [3.14] * 6, True), maxgen=8)
    cs['sep_powers_best1exp'] = dict(dim=7, NP=28, strategy='Best1Exp', cost='powers', CR=0.9, F=0.8, seed=407,
                                     init=('random', [-1.5] * 7, [1.5] * 7), maxgen=8)
    # closed-form functions of mystic.models evaluated by one thread on the device
    cs['fix_branins_best1exp'] = dict(dim=2, NP=24, strategy='Best1Exp', cost='branins', CR=0.9, F=0.8, seed=501,
                                      init=('random', [-5., 0.], [10., 15.]), bounds=([-5., 0.], [10., 15.], True), maxgen=10)
    cs['fix_easom_rand1bin'] = dict(dim=2, NP=24, strategy='Rand1Bin', cost='easom', CR=0.9, F=0.8, seed=502,
                                    init=('random', [0.] * 2, [6.] * 2), maxgen=10)
    cs['fix_goldstein_best1bin'] = dict(dim=2, NP=20, strategy='Best1Bin', cost='goldstein', CR=0.9, F=0.8, seed=503,
                                        init=('random', [-2.] * 2, [2.] * 2), bounds=([-2.] * 2, [2.] * 2, False), maxgen=10)
    cs['fix_peaks_rand2exp'] = dict(dim=2, NP=24, strategy='Rand2Exp', cost='peaks', CR=0.9, F=0.7, seed=504,
                                    init=('random', [-3.] * 2, [3.] * 2), maxgen=10)
    cs['fix_venkat91_randtobest'] = dict(dim=2, NP=24, strategy='RandToBest1Exp', cost='venkat91', CR=0.9, F=0.8, seed=505,
                                         init=('random', [-10.] * 2, [10.] * 2), bounds=([-10.] * 2, [10.] * 2, True), maxgen=10)
    cs['fix_fosc3d_best2exp'] = dict(dim=2, NP=24, strategy='Best2Exp', cost='fosc3d', CR=0.9, F=0.6, seed=506,
                                     init=('random', [-5.] * 2, [5.] * 2), maxgen=10)
    cs['fix_nmin51_best1exp'] = dict(dim=2, NP=24, strategy='Best1Exp', cost='nmin51', CR=0.9, F=0.8, seed=507,
                                     init=('random', [-1.] * 2, [1.] * 2), bounds=([-1.] * 2, [1.] * 2, True), maxgen=10)
    cs['fix_shekel_rand1exp'] = dict(dim=2, NP=24, strategy='Rand1Exp', cost='shekel', CR=0.9, F=0.8, seed=508,
                                     init=('random', [-50.] * 2, [50.] * 2), bounds=([-50.] * 2, [50.] * 2, False), maxgen=10)
    cs['fix_ellipsoid_best1bin'] = dict(dim=12, NP=30, strategy='Best1Bin', cost='ellipsoid', CR=0.9, F=0.8, seed=509,
                                        init=('random', [-65.] * 12, [65.] * 12), maxgen=8)
    cs['fix_paviani_rand1exp'] = dict(dim=10, NP=30, strategy='Rand1Exp', cost='paviani', CR=0.9, F=0.8, seed=510,
                                      init=('random', [2.001] * 10, [9.999] * 10), bounds=([2.001] * 10, [9.999] * 10, True),
                                      maxgen=8)
    # data-fit residual costs: the reference's own CostFactory closures
    cs['fit_poly3_rand1exp'] = dict(dim=3, NP=20, strategy='Rand1Exp', cost='polyfit', CR=0.5, F=0.5, seed=301,
                                    init=('random', [-100.] * 3, [100.] * 3), bounds=([-100.] * 3, [100.] * 3, False),
                                    fit=dict(target=[1., 2., 1.], pts=('arange', -50., 51., 1.)), maxgen=12)
    cs['fit_poly9_best1bin'] = dict(dim=9, NP=36, strategy='Best1Bin', cost='polyfit', CR=0.9, F=0.8, seed=302,
                                    init=('random', [-200.] * 9, [200.] * 9),
                                    fit=dict(target=[128., 0., -256., 0., 160., 0., -32., 0., 1.], pts=('linspace', -1., 1., 200)),
                                    maxgen=8)
    cs['fit_lorentzian_rand1exp'] = dict(dim=7, NP=28, strategy='Rand1Exp', cost='lorentzian', CR=0.9, F=0.8, seed=303,
                                         init=('random', [0.5, 30., -15., 10., 0., 0.05, 100.], [2., 60., -5., 50., 2., 1., 200.]),
                                         bounds=([0.5, 30., -15., 10., 0., 0.05, 100.], [2., 60., -5., 50., 2., 1., 200.], True),
                                         fit=dict(target=[1., 45., -10., 20., 1., 0.1, 120.], pts=('arange', 0.05, 3.0, 0.1)),
                                         maxgen=10)
    cs['fit_decay_best1exp'] = dict(dim=5, NP=30, strategy='Best1Exp', cost='decay', CR=0.9, F=0.8, seed=304,
                                    init=('random', [0., 100., 100., 10., 100.], [20., 1500., 300., 60., 400.]),
                                    bounds=([0., 100., 100., 10., 100.], [20., 1500., 300., 60., 400.], True), maxgen=10)
    return cs


def main(argv):
    outdir = os.path.join(REPO, 'tests', 'golden')
    os.makedirs(outdir, exist_ok=True)
    cs = cases()
    names = argv[1:] or sorted(cs)
    for name in names:
        data = record(cs[name])
        path = os.path.join(outdir, name + '.npz')
        np.savez_compressed(path, **data)
        meta = json.loads(str(data['meta']))
        print('%-32s G=%3d evals=%6d bestE=%.17g  %6.1f KiB' % (
            name, meta['generations'], meta['evaluations'], data['bestE'][-1], os.path.getsize(path) / 1024.))


if __name__ == '__main__':
    main(sys.argv)
