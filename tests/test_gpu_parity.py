"""Parity of the CUDA path (through the C ABI) with the oracle and the recorded reference runs.

Bars (north_star): trial vectors, selections (populations) and best-member indices bit-exact in
validation (replay) mode; energies within 1e-12 relative.  For griewangk the tolerance is
|dE| <= 1e-12 * max(|f|, term1 + |term2| + 1) because term1 - term2 + 1 cancels near the optimum
and CUDA's cos differs from glibc's by <= 2 ulp (hard parts)."""
import random

import numpy as np
import pytest

from oracle import de_oracle as orc
from tests import _golden as G

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def energy_close(dev, ref, x=None, cost=None, what=''):
    dev = np.asarray(dev, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert dev.shape == ref.shape, what
    both_inf = np.isinf(dev) & np.isinf(ref) & (np.sign(dev) == np.sign(ref))
    assert (np.isinf(dev) == np.isinf(ref)).all(), '%s: inf pattern differs' % what
    scale = np.abs(ref)
    if cost == 'griewangk' and x is not None:
        xx = np.atleast_2d(x)
        scale = np.maximum(scale, (xx * xx).sum(axis=1) / 4000. + 2.0)
    elif cost in ('rastrigin', 'schwefel', 'ackley', 'michal') and x is not None:
        # sums of transcendental terms that cancel (CUDA's sin/cos/exp differ from glibc's by <= 2 ulp):
        # the tolerance is relative to the magnitude of the terms, not of the (possibly small) total
        xx = np.atleast_2d(x)
        mag = {'rastrigin': (xx * xx).sum(axis=1) + 20.0 * xx.shape[1], 'schwefel': np.abs(xx).sum(axis=1),
               'ackley': np.full(xx.shape[0], 45.0), 'michal': np.full(xx.shape[0], float(xx.shape[1]))}[cost]
        scale = np.maximum(scale, mag)
    elif cost in orc.FIXED_COST_SCALE:
        scale = np.maximum(scale, orc.FIXED_COST_SCALE[cost])
    ok = both_inf | (np.abs(dev - ref) <= RTOL * scale)
    if not ok.all():
        i = int(np.argmax(~ok))
        raise AssertionError('%s: energy mismatch at %d: %r vs %r' % (what, i, dev[i], ref[i]))


def configure(eng, meta, z=None):
    from mystic_b200 import models, penalty
    lo, hi, mode = G.bounds_of(meta)
    eng.set_bounds(mode, lo, hi)
    cost = G.device_cost_of(meta, z)
    eng.set_cost(cost.cost_id, cost.params(tuple(meta.get('ExtraArgs') or ())))
    if meta.get('penalty'):
        eng.set_penalty(*G.device_penalty_of(meta).device_terms(meta['dim']))
    sid = orc.STRATEGIES[meta['strategy']][0]
    eng.set_strategy(sid, meta['F_eff'], meta['CR_eff'])


@pytest.fixture(params=['rows_tiled', 'rows_per_warp', 'thread_per_row', 'warp_per_row'])
def kernel_path(request, monkeypatch):
    """rows of <= 128 dimensions normally run several to a warp: de_rows.cuh (MB2_SMALL_KERNEL=3, the
    default) where it applies (rosen / corana / griewangk, 16 <= D <= 128, D % 4 == 0), de_subwarp.cuh
    (1) otherwise; 2 selects the thread-per-candidate kernel (de_small.cuh) and 0 sends them through
    the warp/group-per-row kernels (de_kernels.cuh, de_tma.cuh), so the recordings exercise every
    kernel family"""
    monkeypatch.setenv('MB2_SMALL_KERNEL', {'rows_tiled': '3', 'rows_per_warp': '1', 'thread_per_row': '2',
                                            'warp_per_row': '0'}[request.param])
    return request.param


@pytest.mark.parametrize('name', G.names())
def test_engine_replay_matches_reference_recording(name, kernel_path):
    """free-running: recorded initial population + recorded draws in; every generation's trial
    vectors / populations / best solution bit-exact, energies <= 1e-12, evaluation counts exact"""
    from mystic_b200.engine import DeviceEngine
    meta, z = G.load(name)
    NP, D = z['pop0'].shape
    eng = DeviceEngine(NP, D, keep_trials=True)
    configure(eng, meta, z)
    eng.upload_population(z['pop0'])
    evals = 0
    for g in range(meta['generations'] + 1):
        if g == 0:
            eng.eval_initial()
        else:
            eng.set_rng_replay(z['donors'][g - 1], z['nstart'][g - 1], z['uni'][g - 1])
            eng.step()
        best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
        assert gen == g
        trial, trial_e = eng.trials()
        tag = '%s gen %d' % (name, g)
        G.assert_bits_equal(trial, z['trial'][g], tag + ' trial')
        energy_close(trial_e, z['trialE'][g], trial, meta['cost'], tag + ' trialE')
        G.assert_bits_equal(eng.population(), z['pop'][g], tag + ' population')
        energy_close(eng.energies(), z['popE'][g], eng.population(), meta['cost'], tag + ' popE')
        energy_close([best_e], [z['bestE'][g]], best_x, meta['cost'], tag + ' bestE')
        G.assert_bits_equal(best_x, z['bestX'][g], tag + ' bestX')
        evals += NP - n_inf
        assert evals == int(z['evals'][g]), tag
        prev = z['popE'][g - 1] if g else np.full(NP, np.inf)
        assert n_acc == int((z['trialE'][g] < prev).sum()), tag
        if g == 0 or z['bestE'][g] < z['bestE'][g - 1]:
            # the member that became best is where the kernel says it is
            assert best_idx >= 0 and np.array_equal(eng.population()[best_idx], best_x), tag
    eng.close()


def _solver_case(meta, z, rng):
    from tests.test_host_solver import make_solver
    return make_solver(meta, z, rng, engine=None)


@pytest.mark.parametrize('name', G.names())
def test_solver_reference_rng_reproduces_seeded_reference_run(name):
    """the drop-in solver on the GPU, rng='reference', same seed and Mersenne-Twister state as the
    reference run: same generations, evaluations, stop message, bestSolution (bit-exact) and
    energy history"""
    from tests.test_host_solver import solve
    meta, z = G.load(name)
    random.seed(meta['seed'])
    s = _solver_case(meta, z, 'reference')
    if meta['init'][0] == 'random':
        s.SetRandomInitialPoints(list(meta['init'][1]), list(meta['init'][2]))
    else:
        s.SetInitialPoints(list(meta['init'][1]), meta['init'][2])
    G.restore_mt_state(z)
    solve(s, meta, z)
    assert s.generations == meta['generations']
    assert s.evaluations == meta['evaluations']
    assert str(s.Terminated(info=True)) == meta['stop']
    energy_close(np.array(s.energy_history), z['energy_history'], np.array(s.solution_history), meta['cost'],
                 'energy_history')
    G.assert_bits_equal(np.array(s.bestSolution), z['bestX'][-1], 'bestSolution')
    G.assert_bits_equal(s.population, z['pop'][-1], 'population')


def test_baseline_config1_on_device():
    from mystic_b200 import DifferentialEvolutionSolver2, models, strategy, termination
    random.seed(123)
    s = DifferentialEvolutionSolver2(3, 40, rng='reference')
    s.SetRandomInitialPoints([0.] * 3, [2.] * 3)
    s.SetEvaluationLimits(generations=99999)
    s.Solve(models.rosen, termination.VTR(1e-4), strategy=strategy.Best1Exp, CrossProbability=0.5, ScalingFactor=0.6)
    assert s.generations == 39 and s.evaluations == 1600
    assert abs(s.bestEnergy - 7.710567500336223e-05) <= RTOL * 7.710567500336223e-05
    assert list(s.bestSolution) == [1.0012335280875173, 1.0029551101232417, 1.0052618377557658]


PHILOX_CASES = [
    # (dim, NP, strategy, cost, extra, bounds(lo,hi,mode), penalty terms, CR, F, gens)
    (16, 64, 'Best1Bin', 'rosen', (), (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 12),
    (16, 64, 'Best1Bin', 'rosen', (), (-2., 2., orc.BOUNDS_REJECT), None, 0.9, 0.8, 12),
    (7, 33, 'Rand1Exp', 'rosen', (), None, None, 0.9, 0.8, 12),
    (130, 40, 'Best1Exp', 'rosen', (), None, None, 0.99, 0.8, 6),
    (300, 36, 'RandToBest1Exp', 'griewangk', (), (-400., 400., orc.BOUNDS_CLAMP), None, 1.0, 0.8, 4),
    (64, 128, 'Rand1Exp', 'corana', (), (-1000., 1000., orc.BOUNDS_CLAMP),
     [dict(ptype='quadratic_inequality', G=[1.] * 64, h=100., k=100, hpow=5)], 0.9, 0.8, 8),
    (9, 48, 'Best2Exp', 'chebyshev8cost', (200,), None, None, 1.0, 0.9, 6),
    (2, 24, 'Rand2Bin', 'zimmermann', (), (0., 5., orc.BOUNDS_REJECT), None, 0.9, 0.8, 10),
    (5, 20, 'Best2Bin', 'rosen', (), None, None, 0.0, 0.8, 5),
    (1024, 96, 'Best1Bin', 'rosen', (), (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 3),
    (256, 80, 'Best1Bin', 'griewangk', (), (-400., 400., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 3),
    # TMA group widths 2 and 8, a single group per CTA, rows that are not whole trips
    (512, 40, 'Best1Bin', 'rosen', (), (-5., 5., orc.BOUNDS_REJECT), None, 0.5, 0.8, 3),
    (2048, 20, 'Best1Bin', 'griewangk', (), (-400., 400., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 3),
    (4096, 12, 'Best1Bin', 'rosen', (), None, None, 0.9, 0.8, 3),
    (1000, 30, 'Best1Bin', 'rosen', (), (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 3),
    (129, 30, 'Best1Bin', 'rosen', (), (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 3),
    (6001, 8, 'Rand2Exp', 'griewangk', (), None, None, 0.999, 0.8, 2),
    # short rows, several to a warp (de_rows.cuh): both group widths, every donor count, windows that
    # wrap / cover the whole row / are empty, NP that leaves idle groups in the last warp
    (32, 50, 'Best1Bin', 'rosen', (), (-5., 5., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6),
    (128, 41, 'Rand2Exp', 'griewangk', (), (-400., 400., orc.BOUNDS_REJECT), None, 0.97, 0.8, 5),
    (100, 30, 'Best2Exp', 'rosen', (), None, None, 0.98, 0.7, 5),
    (64, 37, 'RandToBest1Exp', 'griewangk', (), (-400., 400., orc.BOUNDS_CLAMP), None, 1.0, 0.8, 4),
    (48, 33, 'Best1Exp', 'corana', (), (-1000., 1000., orc.BOUNDS_REJECT),
     [dict(ptype='quadratic_inequality', G=[1.] * 48, h=100., k=100, hpow=5),
      dict(ptype='linear_equality', G=[0.5] * 48, h=3., k=10, hpow=2)], 0.9, 0.8, 6),
    (16, 19, 'Rand1Exp', 'rosen', (), (-2., 2., orc.BOUNDS_CLAMP), None, 0.0, 0.8, 3),
    (128, 64, 'Best1Bin', 'corana', (), (-1000., 1000., orc.BOUNDS_CLAMP), None, 0.3, 0.9, 4),
    # separable test functions (one kernel body, function picked at run time)
    (24, 40, 'Best1Bin', 'rastrigin', (), (-5.12, 5.12, orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6),
    (200, 36, 'Best1Bin', 'ackley', (), (-30., 30., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 4),
    (10, 33, 'Rand2Exp', 'schwefel', (), (-500., 500., orc.BOUNDS_REJECT), None, 0.9, 0.8, 6),
    (300, 30, 'Best2Exp', 'sphere', (), None, None, 0.95, 0.8, 4),
    (40, 36, 'Rand1Exp', 'step', (), (-7., 7., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 5),
    (12, 30, 'Best1Bin', 'michal', (), (0., 3.14, orc.BOUNDS_CLAMP), None, 0.9, 0.8, 5),
    (9, 28, 'RandToBest1Exp', 'powers', (), (-1.5, 1.5, orc.BOUNDS_REJECT), None, 0.9, 0.8, 5),
    # closed-form functions one thread evaluates (fixed_cost)
    (2, 40, 'Best1Bin', 'branins', (), (-5., 15., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6),
    (2, 33, 'Rand1Exp', 'peaks', (), None, None, 0.9, 0.8, 6),
    (2, 36, 'Best2Exp', 'shekel', (), (-50., 50., orc.BOUNDS_REJECT), None, 0.9, 0.8, 5),
    (20, 40, 'RandToBest1Exp', 'ellipsoid', (), (-65., 65., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 5),
    (10, 36, 'Rand1Exp', 'paviani', (), (2.5, 9.5, orc.BOUNDS_CLAMP), None, 0.9, 0.8, 5),
    (130, 30, 'Best1Bin', 'ellipsoid', (), None, None, 0.9, 0.8, 3),
    # data-fit residual costs: extra = (evaluation points, observations)
    (5, 40, 'Best1Exp', 'polyfit', (np.linspace(-2., 2., 77), np.cos(np.linspace(-2., 2., 77))), (-3., 3., orc.BOUNDS_CLAMP),
     None, 0.9, 0.8, 6),
    (9, 30, 'Best1Bin', 'polyfit', (np.linspace(-1., 1., 300), np.linspace(-1., 1., 300) ** 8), None, None, 0.9, 0.7, 5),
    (7, 33, 'Rand1Exp', 'lorentzian', (np.arange(0.05, 3.0, 0.1), 1.0 / (1.0 + np.arange(0.05, 3.0, 0.1) ** 2)),
     (0.1, 3., orc.BOUNDS_CLAMP), None, 0.9, 0.8, 6),
]


def _device_cost(cost, extra, dim):
    from mystic_b200 import models
    if cost == 'polyfit':
        return models.poly.CostFactory2(extra[0], extra[1], dim), ()
    if cost == 'lorentzian':
        return models.lorentzian.CostFactory2(extra[0], extra[1], dim), ()
    return getattr(models, cost), extra


@pytest.mark.parametrize('case', PHILOX_CASES, ids=lambda c: '%s-%s-D%d' % (c[2], c[3], c[0]))
def test_philox_mode_matches_oracle_philox(case, kernel_path):
    """production RNG: device-side population init, donors and crossover masks follow the same
    Philox protocol as the oracle, so whole trajectories are comparable bit for bit"""
    from mystic_b200 import models, penalty
    from mystic_b200.engine import DeviceEngine
    dim, NP, strat, cost, extra, bounds, pen, CR, F, gens = case
    seed = 0x5EED0000 + dim * 131 + NP
    lo = hi = None
    mode = orc.BOUNDS_NONE
    if bounds:
        lo, hi, mode = np.full(dim, bounds[0]), np.full(dim, bounds[1]), bounds[2]
    ilo = np.full(dim, -3.0) if bounds is None else lo * 1.25   # start partly outside: gen-0 clip
    ihi = np.full(dim, 3.0) if bounds is None else hi * 1.25

    eng = DeviceEngine(NP, dim, keep_trials=True)
    c, dev_extra = _device_cost(cost, extra, dim)
    eng.set_cost(c.cost_id, c.params(dev_extra))
    eng.set_bounds(mode, lo, hi)
    terms = pen or ()
    if pen:
        def p(x):
            return 0.0
        for t in pen:
            p = getattr(penalty, t['ptype'])(penalty.linear(t['G'], t['h']), k=t['k'], h=t['hpow'])(p)
        eng.set_penalty(*p.device_terms(dim))
    eng.set_strategy(orc.STRATEGIES[strat][0], F, CR)
    eng.set_rng_philox(seed)
    eng.init_population(seed, ilo, ihi)

    pop0 = orc.philox_init_population(seed, NP, ilo, ihi)
    G.assert_bits_equal(eng.population(), pop0, 'philox init population')
    st = orc.DEState(pop0, lo, hi, mode)
    ocost = orc.make_cost(cost, extra)
    for g in range(gens + 1):
        if g == 0:
            eng.eval_initial()
            orc.step0(st, ocost, terms)
        else:
            eng.step()
            orc.step(st, strat, CR, F, ocost, terms, philox=dict(seed=seed, offset=0))
        best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
        tag = 'gen %d' % g
        trial, trial_e = eng.trials()
        G.assert_bits_equal(trial, st.trialSolution, tag + ' trial')
        energy_close(trial_e, st.trialEnergy, trial, cost, tag + ' trialE')
        G.assert_bits_equal(eng.population(), st.population, tag + ' population')
        G.assert_bits_equal(best_x, st.bestSolution, tag + ' bestX')
        energy_close([best_e], [st.bestEnergy], best_x, cost, tag + ' bestE')
        assert n_inf == int(np.isinf(st.trialEnergy).sum())
        assert n_acc == int(st.accepted.sum())
    eng.close()


def test_device_cost_callable_known_answers():
    """SURVEY 8(c) known answers through mb2_eval_cost (DeviceCost.__call__)"""
    from mystic_b200 import models
    def close(a, b):
        assert abs(a - b) <= RTOL * max(abs(b), 1e-300) or a == b, (a, b)
    close(models.rosen([0.8, 1.2, 0.7]), 86.19999999999997)
    assert models.rosen([1., 1., 1.]) == 0.0
    close(models.rosen(np.linspace(-2, 2, 7)), 3608.567901234568)
    close(models.corana([0.3, -0.7, 1.234, 5.0]), 859.6112499999999)
    close(models.corana([0.3, -0.7, 1.234, 5.0] + [9.0] * 60), 859.6112499999999)
    assert models.corana([0.04, -0.04, 0.01, 0.0]) == 0.0
    close(models.griewangk([1., 2., 3.]), 1.0170279701835734)
    assert abs(models.griewangk([0.] * 10)) <= 1e-15
    close(models.griewangk(np.linspace(-400, 400, 10)), 164.05614975873598)
    assert models.zimmermann([7., 2.]) == 0.0
    assert models.zimmermann([1., 1.]) == 7.0
    assert models.zimmermann([-1., 6.]) == 1600.0
    # minima quoted by the reference's docstrings (pohlheim.py:276, 311, 350; dejong.py:264)
    assert abs(models.branins([np.pi, 2.275]) - 0.397887) < 1e-6
    assert abs(models.easom([np.pi, np.pi]) + 1.0) < 1e-15
    assert models.goldstein([0., -1.]) == 3.0
    assert models.ellipsoid([1., 2., 3.]) == 1. + 9. + 36.
    assert models.shekel([-32., -32.]) == 0.0
    assert models.paviani([1., 5.]) == np.inf
    assert models.chebyshev8cost(models.CHEBYSHEV_TARGETS[8]) == 0.0
    close(models.chebyshev8cost([1.] * 9), 7823.744334789609)
    close(models.chebyshev8cost([1.] * 9, 4096), 22615.651073208195)
    close(models.chebyshev8cost([100, -50, 3, 7, -20, 1, 0.5, -2, 1], 4096), 2014464.4595239898)


def test_costs_random_batches_match_oracle(kernel_path):
    """every device cost on random batches, several row lengths (vector and scalar load paths)"""
    from mystic_b200 import models
    rng = np.random.default_rng(7)
    for name, dims, span, extra in (('rosen', (2, 3, 8, 16, 31, 64, 100, 128, 257, 1024), 3., ()),
                                    ('corana', (1, 2, 3, 4, 9, 16, 64, 128), 1000., ()),
                                    ('griewangk', (1, 5, 10, 20, 63, 64, 128, 256), 400., ()),
                                    ('zimmermann', (2,), 8., ()),
                                    ('sphere', (1, 7, 64, 300), 5., ()),
                                    ('rastrigin', (2, 10, 64, 257), 5.12, ()),
                                    ('schwefel', (3, 20, 128), 500., ()),
                                    ('ackley', (2, 16, 100, 1024), 30., ()),
                                    ('step', (1, 9, 64, 200), 8., ()),
                                    ('michal', (2, 10, 64), 3.14, ()),
                                    ('powers', (1, 7, 20), 1.5, ()),
                                    ('branins', (2,), 10., ()), ('easom', (2,), 6., ()), ('goldstein', (2,), 2., ()),
                                    ('peaks', (2,), 3., ()), ('venkat91', (2,), 10., ()), ('fosc3d', (2,), 3., ()),
                                    ('nmin51', (2,), 1., ()), ('shekel', (2,), 50., ()), ('ellipsoid', (1, 5, 33, 200), 65., ()),
                                    ('paviani', (1, 10, 40), None, ()),
                                    ('chebyshev8cost', (9,), 100., (61,)),
                                    ('chebyshev8cost', (9,), 3., (4096,)),
                                    ('chebyshev16cost', (17,), 2., (333,))):
        for D in dims:
            X = rng.uniform(-span, span, size=(97, D)) if span else rng.uniform(1.5, 10.5, size=(97, D))
            got = getattr(models, name)(X, *extra)
            want = orc.make_cost(name, extra)(X)
            energy_close(got, want, X, name, '%s D=%d' % (name, D))


def test_datafit_costs_random_batches_match_oracle(kernel_path):
    """forward_model.CostFactory costs (polynomial least squares, lorentzian) on random parameter
    batches: Horner / elementwise paths <= 1e-12 of the oracle (numpy in the reference's order), with
    and without sigma; the DMMA contraction of the polynomial fit on the same batches"""
    from mystic_b200 import models
    rng = np.random.default_rng(11)
    pts = np.linspace(-1., 1., 333)
    for n, span in ((3, 100.), (9, 200.), (17, 5.)):
        target = rng.uniform(-span, span, size=n)
        cost = models.poly.CostFactory(target, pts)
        X = rng.uniform(-span, span, size=(97, n))
        want = orc.cost_polyfit(X, pts, cost.observations)
        energy_close(cost(X), want, what='polyfit n=%d' % n)
        energy_close(cost(X, tensor_cores=True), want, what='polyfit DMMA n=%d' % n)
        scaled = models.DataFitCost('poly.cost', models.COST_POLYFIT, pts, cost.observations, n, sigma=2.5)
        energy_close(scaled(X), orc.cost_polyfit(X, pts, cost.observations, 2.5), what='polyfit sigma n=%d' % n)
        energy_close(scaled(X, tensor_cores=True), orc.cost_polyfit(X, pts, cost.observations, 2.5), what='polyfit DMMA sigma')
    assert models.poly.CostFactory([1., 2., 1.], np.arange(-50., 51.))([1., 2., 1.]) == 0.0
    lo = np.array([0.5, 30., -15., 10., 0., 0.05, 100.])
    hi = np.array([2., 60., -5., 50., 2., 1., 200.])
    bins = np.arange(0.05, 3.0, 0.1)
    lcost = models.lorentzian.CostFactory([1., 45., -10., 20., 1., 0.1, 120.], bins)
    X = lo + (hi - lo) * rng.uniform(size=(97, 7))
    energy_close(lcost(X), orc.cost_lorentzian(X, bins, lcost.observations), what='lorentzian')
    assert lcost([1., 45., -10., 20., 1., 0.1, 120.]) == 0.0
    # dual exponential decay (models/br8.py), every residual weighed by its own sigma = sqrt(observation);
    # a per-point sigma on the polynomial fit too
    t = np.arange(15., 900., 15.)
    dcost = models.decay.CostFactory([10., 900., 80., 27., 225.], t)
    assert isinstance(dcost.sigma, np.ndarray) and dcost([10., 900., 80., 27., 225.]) < 1e-18   # CUDA's exp vs numpy's: <= 1 ulp
    X = np.array([0., 100., 100., 10., 100.]) + np.array([20., 1400., 200., 50., 300.]) * rng.uniform(size=(97, 5))
    energy_close(dcost(X), orc.cost_decay(X, t, dcost.observations, dcost.sigma), what='decay')
    counts = np.round(dcost.observations) + 1.0
    energy_close(models.decay.CostFactory2(t, counts, 5)(X), orc.cost_decay(X, t, counts, np.sqrt(counts)), what='decay data')
    sig = rng.uniform(0.5, 3.0, size=pts.size)
    wcost = models.DataFitCost('poly.cost', models.COST_POLYFIT, pts, np.cos(pts), 4, sigma=sig)
    X = rng.uniform(-3., 3., size=(97, 4))
    energy_close(wcost(X), orc.cost_polyfit(X, pts, np.cos(pts), sig), what='polyfit per-point sigma')
    energy_close(wcost(X, tensor_cores=True), orc.cost_polyfit(X, pts, np.cos(pts), sig), what='polyfit per-point sigma (Horner: no DMMA variant)')


def test_polyfit_dmma_trajectory_matches_oracle():
    """the tensor-core polynomial fit inside the fused generation kernel: trial vectors / selections
    bit-exact against the oracle on a random population, energies <= 1e-12"""
    from mystic_b200 import models
    from mystic_b200.engine import DeviceEngine
    D, NP, M = 9, 200, 513
    pts = np.linspace(-1., 1., M)
    obs = np.polyval([128., 0., -256., 0., 160., 0., -32., 0., 1.], pts) + 0.25 * np.sin(7 * pts)
    cost = models.poly.CostFactory2(pts, obs, D)
    seed = 0xF17
    eng = DeviceEngine(NP, D, keep_trials=True)
    eng.set_cost(cost.device_id(True, D), cost.params())
    eng.set_bounds(orc.BOUNDS_NONE)
    eng.set_strategy(orc.STRATEGIES['Best1Exp'][0], 0.9, 1.0)
    eng.set_rng_philox(seed)
    ilo, ihi = np.full(D, -100.), np.full(D, 100.)
    eng.init_population(seed, ilo, ihi)
    st = orc.DEState(orc.philox_init_population(seed, NP, ilo, ihi), None, None, orc.BOUNDS_NONE)
    ocost = orc.make_cost('polyfit', (pts, obs))
    for g in range(5):
        if g == 0:
            eng.eval_initial()
            orc.step0(st, ocost, ())
        else:
            eng.step()
            orc.step(st, 'Best1Exp', 1.0, 0.9, ocost, (), philox=dict(seed=seed, offset=0))
        best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
        trial, trial_e = eng.trials()
        G.assert_bits_equal(trial, st.trialSolution, 'gen %d trial' % g)
        energy_close(trial_e, st.trialEnergy, what='gen %d trialE' % g)
        G.assert_bits_equal(eng.population(), st.population, 'gen %d population' % g)
        G.assert_bits_equal(best_x, st.bestSolution, 'gen %d bestX' % g)
    eng.close()


def test_full_size_properties_config2():
    """BASELINE config 2 at full size (Rosenbrock 1024-D, NP=65536, Best1Bin, clamp bounds):
    size-independent properties instead of an oracle run."""
    from mystic_b200 import models
    from mystic_b200.engine import DeviceEngine
    import torch
    NP, D = 65536, 1024
    lo, hi = np.full(D, -5.), np.full(D, 5.)
    runs = []
    for rep in range(2):
        eng = DeviceEngine(NP, D)
        eng.set_cost(models.rosen.cost_id, [])
        eng.set_bounds(orc.BOUNDS_CLAMP, lo, hi)
        eng.set_strategy(orc.STRATEGIES['Best1Bin'][0], 0.8, 0.9)
        eng.set_rng_philox(0x5EED)
        eng.init_population(0x5EED, lo, hi)
        eng.eval_initial()
        hist = [eng.read_result()]
        e_prev = eng.energy_tensor().clone()
        for g in range(4):
            eng.step()
            r = eng.read_result()
            e_now = eng.energy_tensor()
            assert bool((e_now <= e_prev).all())                       # greedy selection never worsens a member
            assert int((e_now < e_prev).sum()) == r[3]                  # n_accepted
            assert r[0] == float(e_now.min())                           # bestEnergy = min(popEnergy)
            assert r[0] <= hist[-1][0]
            idx = int(torch.argmin(e_now))                              # torch.argmin: first minimal index
            pop = eng.population_tensor()
            assert bool((pop >= -5.).all()) and bool((pop <= 5.).all())  # clamp respected
            if r[1] >= 0 and r[0] < hist[-1][0]:
                assert r[1] == idx
                assert np.array_equal(pop[idx].cpu().numpy(), r[5])     # bestSolution is that member
            hist.append(r)
            e_prev = e_now.clone()
        # stored energies are the cost of the stored members (idempotence of evaluation)
        sample = torch.arange(0, NP, 997, device=pop.device)
        re = eng.eval_cost(pop[sample].cpu().numpy())
        assert np.array_equal(re, e_now[sample].cpu().numpy())
        # spot check against the oracle
        sub = pop[sample[:16]].cpu().numpy()
        energy_close(re[:16], orc.cost_rosen(sub), sub, 'rosen', 'oracle spot check')
        runs.append((eng.population_tensor().clone(), [h[0] for h in hist]))
        eng.close()
    assert torch.equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]   # deterministic


def test_griewangk_cosine_absolute_error():
    """the device's griewangk factor uses cos_bounded (one polynomial on [-pi/2, pi/2] after a
    reduction modulo pi): absolute error against CUDA's cos (<= 1 ulp) stays below 4e-16 over the
    whole argument range it is used for"""
    from mystic_b200.engine import DeviceEngine
    eng = DeviceEngine(8, 16)
    for span in (1e-3, 1.0, 3.2, 400.0, 1e5):
        worst = eng.selftest_cosine(1 << 24, seed=int(span * 11) + 5, span=span)
        assert 0.0 <= worst <= 4e-16, (span, worst)
    eng.close()


def test_table_divide_is_ieee_exact():
    """griewangk's x / sqrt(i+1) is computed from a {sqrt, reciprocal} table with two Markstein
    refinements; it must equal the IEEE divide bit for bit"""
    from mystic_b200 import models
    from mystic_b200.engine import DeviceEngine
    eng = DeviceEngine(8, 1024)
    eng.set_cost(models.griewangk.cost_id, [])
    for span in (1e-3, 1.0, 400.0, 1e6, 1e150):
        assert eng.selftest_division(1 << 24, seed=int(span * 7) + 3, span=span) == 0, span
    eng.close()


@pytest.mark.parametrize('name,extra,D,span', [('chebyshev8cost', (61,), 9, 100.), ('chebyshev8cost', (4096,), 9, 3.),
                                               ('chebyshev16cost', (333,), 17, 2.), ('chebyshev4cost', (100,), 5, 10.),
                                               ('chebyshev2cost', (61,), 3, 5.), ('chebyshev6cost', (77,), 7, 5.)])
def test_chebyshev_dmma_cost_matches_oracle(name, extra, D, span):
    """FP64 tensor-core (DMMA) evaluation of the Chebyshev cost against the oracle's Horner evaluation
    on random populations; reports how many abscissae sit within 1e-9 of the |px| = 1 threshold"""
    from mystic_b200 import models
    rng = np.random.default_rng(11)
    X = rng.uniform(-span, span, size=(1000, D))
    X[0] = models.CHEBYSHEV_TARGETS[D - 1]   # the optimum itself: px equioscillates at +-1
    got = getattr(models, name)(X, *extra, tensor_cores=True)
    want = orc.make_cost(name, extra)(X)
    xs = orc.chebyshev_points(extra[0])
    px = np.stack([orc._polyeval(X, x) for x in xs], axis=1)
    near = int((np.abs(np.abs(px[1:]) - 1.0) < 1e-9).sum())
    print('%s M=%d: %d of %d points within 1e-9 of the threshold (random rows)' % (name, extra[0], near, px[1:].size))
    np.testing.assert_allclose(got[1:], want[1:], rtol=1e-12, atol=0)
    assert abs(got[0] - want[0]) <= 4.0 * extra[0]   # at the optimum every point may flip by (1-(-1))^2 = 4


def test_chebyshev_dmma_trajectory_matches_oracle():
    """whole generations through de_step_cheb_mma_kernel (Philox draws, all crossover kinds)"""
    from mystic_b200 import models
    from mystic_b200.engine import DeviceEngine
    for strat, CR, F, M, NP, bounds in (('Best1Exp', 1.0, 0.9, 128, 200, None), ('Best1Bin', 0.9, 0.8, 61, 130, (-50., 50.)),
                                        ('Rand2Exp', 0.8, 0.7, 100, 77, None), ('RandToBest1Bin', 0.9, 0.8, 64, 64, None)):
        D, seed = 9, 0xCEB + NP
        lo, hi = np.full(D, -100.), np.full(D, 100.)
        eng = DeviceEngine(NP, D, keep_trials=True)
        eng.set_cost(models.COST_CHEBYSHEV_MMA, models.chebyshev8cost.params((M,)))
        mode = orc.BOUNDS_NONE
        if bounds:
            mode = orc.BOUNDS_CLAMP
            eng.set_bounds(mode, np.full(D, bounds[0]), np.full(D, bounds[1]))
        eng.set_strategy(orc.STRATEGIES[strat][0], F, CR)
        eng.set_rng_philox(seed)
        eng.init_population(seed, lo, hi)
        st = orc.DEState(orc.philox_init_population(seed, NP, lo, hi), None if not bounds else np.full(D, bounds[0]),
                         None if not bounds else np.full(D, bounds[1]), mode)
        cost = orc.make_cost('chebyshev8cost', (M,))
        for g in range(6):
            if g == 0:
                eng.eval_initial()
                orc.step0(st, cost)
            else:
                eng.step()
                orc.step(st, strat, CR, F, cost, philox=dict(seed=seed, offset=0))
            best_e, best_idx, n_inf, n_acc, gen, best_x = eng.read_result()
            trial, trial_e = eng.trials()
            G.assert_bits_equal(trial, st.trialSolution, '%s gen %d trial' % (strat, g))
            np.testing.assert_allclose(trial_e, st.trialEnergy, rtol=1e-12)
            G.assert_bits_equal(eng.population(), st.population, '%s gen %d population' % (strat, g))
            G.assert_bits_equal(best_x, st.bestSolution, '%s gen %d best' % (strat, g))
            assert n_acc == int(st.accepted.sum())
        eng.close()


def test_save_and_load_solver_on_device(tmp_path):
    """restart files (abstract_solver.py:975-1018): device state -> host arrays -> device state"""
    from mystic_b200 import DifferentialEvolutionSolver2, LoadSolver, models, termination

    def make():
        s = DifferentialEvolutionSolver2(256, 300, rng='philox', seed=91, strategy='Best1Bin')
        s.SetRandomInitialPoints([-400.] * 256, [400.] * 256)
        s.SetStrictRanges([-400.] * 256, [400.] * 256, tight=True)
        s.SetTermination(termination.VTR(-1.0))
        s.SetEvaluationLimits(generations=10 ** 6)
        return s
    a = make()
    a.Step(models.griewangk)
    for _ in range(4):
        a.Step()
    path = str(tmp_path / 'restart.pkl')
    a.SaveSolver(path)
    b = LoadSolver(path)
    for _ in range(3):
        a.Step()
        b.Step(models.griewangk)
    assert list(a.energy_history) == list(b.energy_history)
    assert np.array_equal(a.population, b.population)
    assert np.array_equal(a.popEnergy, b.popEnergy)
    assert a.evaluations == b.evaluations and a.generations == b.generations == 7


FULL_SIZE = {
    # BASELINE.json configs 3-5 at full (per-GPU) size: (dim, NP, strategy, cost, extra, tensor cores, CR, F, span, bounds mode, penalty)
    'c3_chebyshev_dmma': (9, 1 << 20, 'Best1Exp', 'chebyshev8cost', (4096,), True, 1.0, 0.9, 100., orc.BOUNDS_NONE, None),
    'c4_corana_penalty': (64, 1 << 20, 'Rand1Exp', 'corana', (), False, 0.9, 0.8, 1000., orc.BOUNDS_CLAMP,
                          dict(G=1.0, h=100., k=100, hpow=5)),
    'c5_griewangk': (256, 1 << 21, 'Best1Bin', 'griewangk', (), False, 0.9, 0.8, 400., orc.BOUNDS_CLAMP, None),
}


@pytest.mark.parametrize('name', sorted(FULL_SIZE))
def test_full_size_properties(name):
    """size-independent properties at the full BASELINE sizes: greedy selection never worsens a
    member, bestEnergy = min(popEnergy) at the lowest index, bounds respected, stored energies are
    the cost of the stored rows (re-evaluation is idempotent), a sample agrees with the oracle,
    accepted-count bookkeeping, determinism of a second run"""
    import torch
    from mystic_b200 import models, penalty
    from mystic_b200.engine import DeviceEngine
    D, NP, strat, cost, extra, tc, CR, F, span, mode, pen = FULL_SIZE[name]
    lo, hi = np.full(D, -span), np.full(D, span)
    c = getattr(models, cost)
    terms = ()
    finals = []
    for rep in range(2):
        eng = DeviceEngine(NP, D)
        eng.set_cost(c.device_id(tc, D), c.params(extra))
        if mode != orc.BOUNDS_NONE:
            eng.set_bounds(mode, lo, hi)
        if pen:
            @penalty.quadratic_inequality(penalty.linear([pen['G']] * D, pen['h']), k=pen['k'], h=pen['hpow'])
            def p(x):
                return 0.0
            eng.set_penalty(*p.device_terms(D))
            terms = [dict(ptype='quadratic_inequality', G=[pen['G']] * D, h=pen['h'], k=pen['k'], hpow=pen['hpow'])]
        eng.set_strategy(orc.STRATEGIES[strat][0], F, CR)
        eng.set_rng_philox(0xBA5E)
        eng.init_population(0xBA5E, lo, hi)
        eng.eval_initial()
        prev = eng.read_result()
        e_prev = eng.energy_tensor().clone()
        for g in range(3):
            eng.step()
            r = eng.read_result()
            e_now = eng.energy_tensor()
            assert bool((e_now <= e_prev).all())
            assert int((e_now < e_prev).sum()) == r[3]
            assert r[0] == float(e_now.min()) and r[0] <= prev[0]
            if r[0] < prev[0]:
                assert r[1] == int(torch.argmin(e_now))
            prev, e_prev = r, e_now.clone()
        pop = eng.population_tensor()
        if mode != orc.BOUNDS_NONE:
            assert bool((pop >= -span).all()) and bool((pop <= span).all())
        sample = torch.arange(0, NP, 4099, device=pop.device)
        rows = pop[sample].cpu().numpy()
        re = eng.eval_cost(rows)
        assert np.array_equal(re, e_now[sample].cpu().numpy())
        st = orc.DEState(rows[:24], lo if mode else None, hi if mode else None, mode)
        want = orc.evaluate(rows[:24], orc.make_cost(cost, extra), st, terms)
        energy_close(re[:24], want, rows[:24], cost, name + ' oracle sample')
        finals.append((pop.clone(), prev[0]))
        eng.close()
    assert torch.equal(finals[0][0], finals[1][0]) and finals[0][1] == finals[1][1]


@pytest.mark.parametrize('term', [('VTR', 1e-2), ('ChangeOverGeneration', 1e-3, 5), ('VTRChangeOverGeneration', 1e-3, 1e-4, 6),
                                  ('NormalizedChangeOverGeneration', 1e-2, 4), ('EvaluationLimits', 37, None)])
@pytest.mark.parametrize('shape', [(3, 20, 'Best1Exp', 'rosen'), (256, 64, 'Best1Bin', 'griewangk')])
def test_batched_solve_on_device_stops_exactly_like_step_by_step(term, shape):
    """device-side termination test (de_finalize_kernel, batch mode) against step-by-step Solve()"""
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    D, NP, strat, cost = shape

    def make(batch):
        s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=4242, strategy=strat)
        s._batch_generations = batch
        s.SetRandomInitialPoints([0.] * D, [2.] * D)
        s.SetEvaluationLimits(generations=150, evaluations=150 * NP)
        s.Solve(getattr(models, cost), getattr(termination, term[0])(*term[1:]))
        return s
    a, b, c = make(0), make(7), make(64)
    for s in (b, c):
        assert s.generations == a.generations and s.evaluations == a.evaluations
        assert list(s.energy_history) == list(a.energy_history)
        assert s.Terminated(info=True) == a.Terminated(info=True)
        assert np.array_equal(s.population, a.population)
        assert np.array_equal(s.popEnergy, a.popEnergy)
        assert [list(x) for x in s.solution_history] == [list(x) for x in a.solution_history]


def test_diffev2_one_call_interface_on_device():
    """mystic_b200.diffev2 (differential_evolution.py:721-933 interface) on the GPU: converges on the
    3-D Rosenbrock problem, is deterministic for a seed, and honours maxiter with warnflag 2"""
    from mystic_b200 import diffev2, models
    bounds = [(-3., 3.)] * 3
    a = diffev2(models.rosen, [0.] * 3, npop=60, bounds=bounds, ftol=1e-8, gtol=30, seed=11, full_output=1, disp=0)
    b = diffev2(models.rosen, [0.] * 3, npop=60, bounds=bounds, ftol=1e-8, gtol=30, seed=11, full_output=1, disp=0)
    assert list(a[0]) == list(b[0]) and a[1:] == b[1:]
    assert a[1] < 1e-6 and np.allclose(a[0], 1.0, atol=1e-2) and a[4] == 0
    c = diffev2(models.rosen, [0.] * 3, npop=60, bounds=bounds, ftol=1e-30, maxiter=9, seed=11, full_output=1, retall=1, disp=0)
    assert c[2] == 9 and c[4] == 2 and len(c[5]) == 10


def test_chebyshev_dmma_huge_coefficients_take_the_horner_fallback():
    """rows whose coefficients could overflow p(x) are evaluated by the reference's own loop inside the
    DMMA kernel (NaN-safe comparisons): same inf / NaN pattern and values as the oracle"""
    from mystic_b200 import models
    rng = np.random.default_rng(3)
    X = rng.uniform(-100., 100., size=(128, 9))
    X[5, 2] = 1e200
    X[17, 0] = -1e308
    X[17, 1] = 1e308
    X[40, 8] = 3e151
    X[77, 4] = np.inf
    got = models.chebyshev8cost(X, 257, tensor_cores=True)
    with np.errstate(all='ignore'):
        want = orc.make_cost('chebyshev8cost', (257,))(X)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
    fin = np.isfinite(want)
    assert fin.sum() >= 124 and np.allclose(got[fin], want[fin], rtol=1e-12, atol=0)


def test_abi_rejects_malformed_requests():
    """error behaviour of the C ABI: negative return code + message, no exception across the boundary, the
    context stays usable"""
    from mystic_b200 import models
    from mystic_b200._native import NativeError
    from mystic_b200.engine import DeviceEngine
    eng = DeviceEngine(16, 3)
    for cost in (models.zimmermann, models.branins, models.shekel):
        with pytest.raises(NativeError, match='two variables|dim == 2'):
            eng.set_cost(cost.cost_id, ())
    with pytest.raises(NativeError, match='unknown cost id'):
        eng.set_cost(99, ())
    with pytest.raises(NativeError, match='mb2_eval_initial must run before'):
        eng.set_cost(models.rosen.cost_id, ())
        eng.step()
    with pytest.raises(NativeError, match='unknown penalty type'):
        eng.set_penalty([9], [[1., 1., 1.]], [0.], [1.])
    with pytest.raises(NativeError, match='do not split into groups'):
        eng.set_islands(4, share_best_within=3)
    eng.set_islands(1)
    with pytest.raises(NativeError, match='no batch of solvers'):
        import ctypes
        from mystic_b200._native import check, lib
        a = (ctypes.c_int32 * 4)()
        check(lib.mb2_run_result_islands(eng._ctx, a, a, None))
    eng.set_cost(models.ellipsoid.cost_id, ())       # any number of variables
    eng.set_rng_philox(1)
    eng.init_population(1, np.full(3, -1.), np.full(3, 1.))
    eng.eval_initial()
    eng.step()
    assert np.isfinite(eng.read_result()[0])
    eng.close()


def test_gradient_norm_tolerance_on_device():
    """termination.py:363-384 with the forward-difference gradient of _scipy060optimize.py:611-619: the D + 1
    cost evaluations go through the device cost (one batch), the verdict equals the oracle's arithmetic"""
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    D = 5
    s = DifferentialEvolutionSolver2(D, 40, rng='philox', seed=21)
    s.SetRandomInitialPoints([-2.] * D, [2.] * D)
    s.SetTermination(termination.VTR(-1.0))
    s.SetEvaluationLimits(generations=10 ** 6)
    s.Step(models.rosen)
    for _ in range(5):
        s.Step()
    x = np.asarray(s.bestSolution, dtype=np.float64)
    eps = float(np.sqrt(np.finfo(float).eps))
    f = orc.cost_rosen(np.vstack([x] + [x + eps * np.eye(D)[k] for k in range(D)]))
    grad = (f[1:] - f[0]) / eps
    for norm, gnorm in ((termination.inf, np.max(np.abs(grad))), (2, np.sum(np.abs(grad ** 2)) ** 0.5), (1, np.sum(np.abs(grad)))):
        assert gnorm > 0
        assert termination.GradientNormTolerance(gnorm * 1.001, norm)(s) is True
        assert termination.GradientNormTolerance(gnorm * 0.999, norm)(s) is False
        assert termination.GradientNormTolerance(gnorm * 1.001, norm)(s, info=True).startswith('GradientNormTolerance with')
    # as a stop condition of Solve(): a huge tolerance stops after the first generation
    s2 = DifferentialEvolutionSolver2(D, 40, rng='philox', seed=21)
    s2.SetRandomInitialPoints([-2.] * D, [2.] * D)
    s2.Solve(models.rosen, termination.GradientNormTolerance(1e30))
    assert s2.generations <= 1 and str(s2.Terminated(info=True)).startswith('GradientNormTolerance with')
