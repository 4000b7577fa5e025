"""Drop-in claim: UNMODIFIED mystic objects (termination conditions, monitors, strategy functions,
model functions, penalty decorators) are accepted by the B200 solver.  Runs only where the
reference is importable (the build container: /root/reference + the klepto shim); the generation
step is supplied by the test-only OracleEngine, the same code path runs on the GPU in
tests/test_gpu_parity.py with mystic_b200's own twins of these objects."""
import os
import random
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if not os.path.isdir('/root/reference/mystic'):
    pytest.skip('reference not available here', allow_module_level=True)
sys.path.insert(0, os.path.join(REPO, 'oracle', '_shim'))
sys.path.insert(1, '/root/reference')

import mystic.models  # noqa: E402
import mystic.monitors  # noqa: E402
import mystic.penalty  # noqa: E402
import mystic.strategy  # noqa: E402
import mystic.termination  # noqa: E402

from mystic_b200 import DifferentialEvolutionSolver2, penalty  # noqa: E402
from tests import _golden as G  # noqa: E402
from tests._oracle_engine import OracleEngine  # noqa: E402


def test_reference_objects_drive_the_solver_to_the_reference_result():
    """BASELINE config 1 with mystic's own rosen / Best1Exp / VTR / Monitor objects"""
    meta, z = G.load('c1_rosen3_best1exp')
    random.seed(123)
    s = DifferentialEvolutionSolver2(3, 40, rng='reference', engine=OracleEngine)
    s.SetRandomInitialPoints([0.] * 3, [2.] * 3)
    s.SetEvaluationLimits(generations=99999)
    mon = mystic.monitors.Monitor()
    s.SetGenerationMonitor(mon)
    s.Solve(mystic.models.rosen, mystic.termination.VTR(1e-4), strategy=mystic.strategy.Best1Exp,
            CrossProbability=0.5, ScalingFactor=0.6)
    assert s.generations == 39 and s.evaluations == 1600
    assert s.bestEnergy == 7.710567500336223e-05
    assert len(mon) == 40 and mon.y[-1] == s.bestEnergy
    assert s.Terminated(info=True).startswith('VTR with')


def test_reference_compound_termination_and_verbose_monitor(capsys):
    T = mystic.termination
    random.seed(5)
    s = DifferentialEvolutionSolver2(4, 24, rng='reference', engine=OracleEngine)
    s.SetRandomInitialPoints([-2.] * 4, [2.] * 4)
    s.SetStrictRanges([-2.] * 4, [2.] * 4)
    s.SetGenerationMonitor(mystic.monitors.VerboseMonitor(5))
    s.SetEvaluationLimits(generations=40)
    s.Solve(mystic.models.rosen, T.Or(T.VTR(1e-8), T.ChangeOverGeneration(1e-12, 500), T.NormalizedChangeOverGeneration(1e-12, 500)),
            strategy=mystic.strategy.Rand1Bin)
    out = capsys.readouterr().out
    assert 'Generation 0 has' in out and 'Generation 5 has' in out
    assert s.generations == 40
    assert s.Terminated(info=True).startswith('EvaluationLimits with')
    for cond in (T.CandidateRelativeTolerance(), T.PopulationSpread(), T.EvaluationLimits(10, 10), T.VTRChangeOverGeneration()):
        assert isinstance(cond(s), (bool, np.bool_))
        assert isinstance(cond(s, info=True), str)


def test_reference_penalty_decorator_with_linear_condition():
    """mystic.penalty.quadratic_inequality over a mystic_b200.penalty.linear condition"""
    @mystic.penalty.quadratic_inequality(penalty.linear([1.] * 8, 100.), k=100, h=5)
    def pen(x):
        return 0.0
    meta, z = G.load('corana8_rand1exp_pen')
    random.seed(meta['seed'])
    s = DifferentialEvolutionSolver2(8, 32, rng='replay', engine=OracleEngine, CrossProbability=meta['CR'],
                                     ScalingFactor=meta['F'])
    s.SetPopulation(z['pop0'])
    s.SetStrictRanges(meta['bounds'][0], meta['bounds'][1], tight=True)
    s.SetPenalty(pen)
    s.SetReplayDraws((z['donors'][g], z['nstart'][g], z['uni'][g]) for g in range(z['donors'].shape[0]))
    s.SetEvaluationLimits(generations=meta['maxgen'])
    s.Solve(mystic.models.corana, mystic.termination.VTR(-1.0), strategy=mystic.strategy.Rand1Exp)
    G.assert_bits_equal(np.array(s.energy_history), z['energy_history'], 'energy history')
    G.assert_bits_equal(s.population, z['pop'][-1], 'population')


def test_reference_lagrange_decorators_with_stored_iterations():
    """mystic.penalty's own lagrange decorators over linear conditions, after two penalty iterations made with the
    reference's store(x, i) / iter(n): the solver reads k, h, n and the stored condition values out of them"""
    meta, z = G.load('pen_lagrange_rosen5')
    pen = lambda x: 0.0   # noqa: E731
    for t in meta['penalty']:
        pen = getattr(mystic.penalty, t['ptype'])(penalty.linear(t['G'], t['h']), k=t['k'], h=t['hpow'])(pen)
    g = pen
    for t in reversed(meta['penalty']):
        for i, x in enumerate(t['store_points']):
            g.store(x, i)
        g = dict(zip(g.__code__.co_freevars, (c.cell_contents for c in g.__closure__)))['f']
    pen.iter(2)
    dev = penalty.resolve(pen)
    assert [t['y'] for t in dev.terms] == [t['stored'] for t in meta['penalty']] and [t['n'] for t in dev.terms] == [2, 2]
    random.seed(meta['seed'])
    s = DifferentialEvolutionSolver2(5, 25, rng='replay', engine=OracleEngine, CrossProbability=meta['CR'], ScalingFactor=meta['F'])
    s.SetPopulation(z['pop0'])
    s.SetPenalty(pen)
    s.SetReplayDraws((z['donors'][k], z['nstart'][k], z['uni'][k]) for k in range(z['donors'].shape[0]))
    s.SetEvaluationLimits(generations=meta['maxgen'])
    s.Solve(mystic.models.rosen, mystic.termination.VTR(-1.0), strategy=getattr(mystic.strategy, meta['strategy']))
    G.assert_bits_equal(np.array(s.energy_history), z['energy_history'], 'energy history')
    G.assert_bits_equal(s.population, z['pop'][-1], 'population')


def test_host_callables_are_rejected():
    s = DifferentialEvolutionSolver2(3, 8, engine=OracleEngine)
    with pytest.raises(TypeError):
        s.SetObjective(mystic.models.wavy1)           # a mystic model without a device twin (vector-valued)
    from mystic_b200 import models
    for name in ('sphere', 'rastrigin', 'schwefel', 'ackley', 'step', 'michal', 'powers', 'rosen', 'corana', 'griewangk', 'zimmermann',
                 'branins', 'easom', 'goldstein', 'peaks', 'venkat91', 'fosc3d', 'nmin51', 'shekel', 'ellipsoid', 'paviani'):
        assert models.resolve(getattr(mystic.models, name)) is getattr(models, name)
    with pytest.raises(TypeError):
        @mystic.penalty.quadratic_inequality(lambda x: sum(x) - 1.0)
        def pen(x):
            return 0.0
        s.SetPenalty(pen)                             # python condition: not device-expressible


def orc_cost(name, extra):
    from oracle import de_oracle as orc
    return orc.make_cost(name, extra)


def test_reference_costfactory_closures_are_recognised():
    """mystic.models.poly.CostFactory(target, pts) / lorentzian.CostFactory2(...) return closures of
    mystic.forward_model.CostFactory.getCostFunction; resolve() turns them into the device's data-fit
    costs with the same points, observations and sigma, and the reference-seeded run reproduces the
    recording made with that very closure"""
    from mystic_b200 import models
    pts = np.arange(-50., 51., 1.)
    ref_cost = mystic.models.poly.CostFactory([1., 2., 1.], pts)
    dev = models.resolve(ref_cost)
    assert dev.cost_id == models.COST_POLYFIT and dev.nparams == 3 and dev.sigma == 1.0
    assert np.array_equal(dev.pts, pts) and np.array_equal(dev.observations, mystic.models.poly.evaluate([1., 2., 1.], pts))
    bins = np.arange(0.05, 3.0, 0.1)
    lor = models.resolve(mystic.models.lorentzian.CostFactory2(bins, np.ones_like(bins), 7))
    assert lor.cost_id == models.COST_LORENTZIAN and lor.nparams == 7
    # mystic.models.br8: the module's own cost (Bevington's counts, sigma = sqrt(counts) per point)
    import importlib
    br8 = importlib.import_module('mystic.models.br8')
    dec = models.resolve(br8.cost)
    assert dec.cost_id == models.COST_DECAY and dec.nparams == 5
    assert np.array_equal(dec.sigma, np.sqrt(br8.data[:, 1])) and np.array_equal(dec.observations, br8.data[:, 1])
    p = [10., 900., 80., 27., 225.]
    assert orc_cost('decay', (dec.pts, dec.observations, dec.sigma))(np.array([p]))[0] == br8.cost(p)
    # other metrics cannot run on the device
    from mystic.forward_model import CostFactory
    cf = CostFactory()
    cf.addModel(mystic.models.poly.ForwardFactory, 3, 'poly')
    with pytest.raises(TypeError):
        models.resolve(cf.getCostFunction(pts, pts, metric=lambda x: np.sum(np.abs(x))))
    # end to end against the recording: the reference closure handed to the B200 solver
    meta, z = G.load('fit_poly3_rand1exp')
    random.seed(meta['seed'])
    s = DifferentialEvolutionSolver2(3, 20, rng='reference', engine=OracleEngine, CrossProbability=0.5, ScalingFactor=0.5)
    s.SetRandomInitialPoints([-100.] * 3, [100.] * 3)
    s.SetStrictRanges([-100.] * 3, [100.] * 3)
    s.SetEvaluationLimits(generations=meta['maxgen'])
    G.restore_mt_state(z)
    s.Solve(ref_cost, mystic.termination.VTR(-1.0), strategy=mystic.strategy.Rand1Exp)
    assert s.generations == meta['generations'] and s.evaluations == meta['evaluations']
    G.assert_bits_equal(np.array(s.bestSolution), z['bestX'][-1], 'bestSolution')
    G.assert_bits_equal(np.array(s.energy_history), z['energy_history'], 'energy_history')


def test_monitor_twins_print_and_log_like_the_reference(tmp_path, capsys):
    """mystic_b200.monitors.{VerboseMonitor, LoggingMonitor, VerboseLoggingMonitor} against the
    reference's classes on the same call sequence: same screen text, same log file (but its date line)"""
    from mystic_b200 import monitors as ours
    calls = [([0.5 * i, -1.0], 2.0 - 0.25 * i, (9 if i == 3 else None)) for i in range(7)]

    def drive(mod, tag):
        capsys.readouterr()
        v = mod.VerboseMonitor(3, 2)
        lg = mod.LoggingMonitor(2, str(tmp_path / (tag + '_l.txt')), new=True, info='hello')
        vl = mod.VerboseLoggingMonitor(1, 3, 6, str(tmp_path / (tag + '_v.txt')), new=True, label='energy')
        for x, y, i in calls:
            for m in (v, lg, vl):
                m(x, y, id=i)
        lg.info('bye')
        screen = capsys.readouterr().out
        files = [(tmp_path / (tag + s)).read_text().splitlines()[1:] for s in ('_l.txt', '_v.txt')]
        return screen, files, (v.y, lg.y, vl.x)

    assert drive(ours, 'ours') == drive(mystic.monitors, 'ref')


def test_diffev2_one_call_interface_matches_the_reference(capsys):
    """mystic_b200.diffev2 against mystic.differential_evolution.diffev2 under the same seed
    (rng='reference', generation step by the test-only OracleEngine): same xopt, fopt, iterations,
    function calls, warnflag, allvecs and the same convergence messages"""
    import mystic.differential_evolution as mde
    from mystic_b200 import diffev2, models
    bounds = [(-3., 3.)] * 3
    for kw in (dict(ftol=1e-3), dict(ftol=1e-6, gtol=12), dict(ftol=1e-9, maxiter=7), dict(ftol=1e-9, maxfun=150)):
        random.seed(77)
        capsys.readouterr()
        want = mde.diffev2(mystic.models.rosen, [0.] * 3, npop=20, bounds=bounds, full_output=1, retall=1, disp=1, **kw)
        want_out = capsys.readouterr().out
        random.seed(77)
        got = diffev2(models.rosen, [0.] * 3, npop=20, bounds=bounds, full_output=1, retall=1, disp=1,
                      rng='reference', engine=OracleEngine, **kw)
        got_out = capsys.readouterr().out
        assert list(got[0]) == list(want[0]) and got[1:5] == want[1:5], (kw, got[1:5], want[1:5])
        assert [list(v) for v in got[5]] == [list(v) for v in want[5]]
        assert got_out == want_out
    # no bounds: the population starts around x0 (SetInitialPoints)
    random.seed(5)
    want = mde.diffev2(mystic.models.rosen, [0.8, 1.2, 0.7], npop=16, ftol=1e-2, full_output=1, disp=0)
    random.seed(5)
    got = diffev2(mystic.models.rosen, [0.8, 1.2, 0.7], npop=16, ftol=1e-2, full_output=1, disp=0, rng='reference',
                  engine=OracleEngine)
    assert list(got[0]) == list(want[0]) and got[1:] == want[1:]


def test_population_reading_terminations_equal_the_reference_ones():
    """SolutionImprovement / CandidateRelativeTolerance / PopulationSpread (termination.py:242-288, 344-361) read
    the population and the trial vectors of the solver: the twins of mystic_b200.termination and the
    unmodified mystic.termination objects must give the same verdict, generation by generation"""
    from mystic_b200 import termination as T
    s = DifferentialEvolutionSolver2(4, 20, rng='philox', seed=11, engine=OracleEngine, keep_trials=True)
    s.SetRandomInitialPoints([-1.] * 4, [1.] * 4)
    s.SetTermination(T.VTR(-1.0))
    s.SetEvaluationLimits(generations=10 ** 6)
    from mystic_b200 import models
    s.Step(models.sphere)
    pairs = [(T.SolutionImprovement(t), mystic.termination.SolutionImprovement(t)) for t in (1e-3, 0.5, 2.0, 50.)]
    pairs += [(T.CandidateRelativeTolerance(x, f), mystic.termination.CandidateRelativeTolerance(x, f))
              for x, f in ((1e-4, 1e-4), (0.5, 0.5), (3.0, 10.0))]
    pairs += [(T.PopulationSpread(t), mystic.termination.PopulationSpread(t)) for t in (1e-6, 0.5, 1e3)]
    verdicts = set()
    for g in range(60):
        s.Step()
        for ours, theirs in pairs:
            assert ours(s) == theirs(s), (g, ours.__doc__)
            assert ours(s, info=True) == theirs(s, info=True), (g, ours.__doc__)
            verdicts.add((ours.__doc__, ours(s)))
    assert len(verdicts) > len(pairs)          # some conditions changed their verdict along the run
    assert T.type(pairs[0][0]) is T.SolutionImprovement and T.type(3) is int
    assert T.state(pairs[0][0]) == {'SolutionImprovement with %s' % {'tolerance': 1e-3}: {'tolerance': 1e-3}}
