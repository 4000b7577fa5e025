"""Host logic of mystic_b200.solver (CPU, no GPU): Step/Solve/termination/keyword handling and
the reference's bookkeeping, checked against recordings of the unmodified reference.  The
generation step itself is supplied by the test-only OracleEngine (tests/_oracle_engine.py); the
CUDA path has the same tests in tests/test_gpu_parity.py."""
import random

import numpy as np
import pytest

import mystic_b200
from mystic_b200 import models, penalty, strategy, termination
from mystic_b200.solver import DifferentialEvolutionSolver2
from tests import _golden as G
from tests._oracle_engine import OracleEngine


def make_solver(meta, z, rng, engine=OracleEngine, **extra):
    kw = dict(rng=rng, engine=engine)
    kw.update(extra)
    if meta['via'] == 'ctor':
        kw.update(CrossProbability=meta['CR'], ScalingFactor=meta['F'])
    s = DifferentialEvolutionSolver2(meta['dim'], meta['NP'], **kw)
    assert s.nPop == z['pop0'].shape[0]
    if meta.get('bounds'):
        lo, hi, tight = meta['bounds']
        if tight:
            s.SetStrictRanges(lo, hi, tight=True)
        else:
            s.SetStrictRanges(lo, hi)
    if meta.get('penalty'):
        s.SetPenalty(G.device_penalty_of(meta))
    s.SetEvaluationLimits(generations=meta['maxgen'])
    return s


def solve(s, meta, z=None):
    term = getattr(termination, meta['termination'][0])(*meta['termination'][1:]) if meta.get('termination') \
        else termination.VTR(-1.0)
    extra = tuple(meta['ExtraArgs']) if meta.get('ExtraArgs') else None
    kw = dict(strategy=getattr(strategy, meta['strategy']))
    if meta['via'] == 'solve':
        kw.update(CrossProbability=meta['CR'], ScalingFactor=meta['F'])
    s.Solve(G.device_cost_of(meta, z), term, ExtraArgs=extra, **kw)


def replay_draws(z):
    for g in range(z['donors'].shape[0]):
        yield z['donors'][g], z['nstart'][g], z['uni'][g]


def check_against_reference(s, meta, z, exact_energy=True):
    assert s.generations == meta['generations']
    assert s.evaluations == meta['evaluations']
    if exact_energy:
        G.assert_bits_equal(np.array(s.energy_history), z['energy_history'], 'energy_history')
    else:
        np.testing.assert_allclose(np.array(s.energy_history), z['energy_history'], rtol=1e-14)
    G.assert_bits_equal(np.array(s.bestSolution), z['bestX'][-1], 'bestSolution')
    G.assert_bits_equal(s.population, z['pop'][-1], 'population')
    assert str(s.Terminated(info=True)) == meta['stop']
    # bookkeeping invariants of mystic/tests/test_solver_sanity.py:77-90
    assert s.generations == len(s._stepmon._y) - 1
    assert list(s.bestSolution) == list(s._stepmon.x[-1])
    assert s.bestEnergy == s._stepmon.y[-1]


@pytest.mark.parametrize('name', G.names())
def test_solve_replay_matches_reference(name):
    """recorded draws in, the reference's whole solve out (counters, histories, stop message)"""
    meta, z = G.load(name)
    s = make_solver(meta, z, 'replay')
    s.SetPopulation(z['pop0'])
    s.SetReplayDraws(replay_draws(z))
    solve(s, meta, z)
    check_against_reference(s, meta, z, exact_energy=(meta['cost'] not in G.COMPENSATED_SUM_COSTS))


@pytest.mark.parametrize('name', G.names())
def test_solve_reference_rng_reproduces_seeded_reference_run(name):
    """rng='reference': the host makes the reference's stdlib calls in the reference's order (no
    recording involved): same seed => same initial population; same Mersenne-Twister state at
    Solve() => the reference's trajectory.  (The state is restored from the recording because
    mystic's SetStrictRanges(tight=True) burns stdlib draws inside sympy simplification.)"""
    meta, z = G.load(name)
    random.seed(meta['seed'])
    s = make_solver(meta, z, 'reference')
    if meta['init'][0] == 'random':
        s.SetRandomInitialPoints(list(meta['init'][1]), list(meta['init'][2]))
    else:
        s.SetInitialPoints(list(meta['init'][1]), meta['init'][2])
    G.assert_bits_equal(s.population, z['pop0'], 'initial population')
    G.restore_mt_state(z)
    solve(s, meta, z)
    check_against_reference(s, meta, z, exact_energy=(meta['cost'] not in G.COMPENSATED_SUM_COSTS))


def test_baseline_config1_golden_values():
    """BASELINE config 1 / SURVEY 8(c): 39 generations, 1600 evaluations, bestEnergy 7.710567500336223e-05"""
    meta, z = G.load('c1_rosen3_best1exp')
    random.seed(123)
    s = DifferentialEvolutionSolver2(3, 40, rng='reference', engine=OracleEngine)
    s.SetRandomInitialPoints([0.] * 3, [2.] * 3)
    s.SetEvaluationLimits(generations=99999)
    s.Solve(models.rosen, termination.VTR(1e-4), strategy=strategy.Best1Exp, CrossProbability=0.5, ScalingFactor=0.6)
    assert s.generations == 39 and s.evaluations == 1600
    assert s.bestEnergy == 7.710567500336223e-05
    assert list(s.bestSolution) == [1.0012335280875173, 1.0029551101232417, 1.0052618377557658]
    assert list(s.energy_history[:5]) == [17.784260845848216, 17.784260845848216, 2.4416751888867902,
                                          2.4416751888867902, 1.5646391444680794]
    # the reference quirk: Solve(CrossProbability=...) does not survive the first Step
    assert (s.probability, s.scale) == (0.9, 0.8)


def test_keyword_handling():
    s = DifferentialEvolutionSolver2(3, 8, engine=OracleEngine, cross=0.3, scale=0.4, strategy='Rand1Exp')
    assert (s.probability, s.scale, s.strategy) == (0.3, 0.4, 'Rand1Exp')
    with pytest.raises(KeyError):
        DifferentialEvolutionSolver2(3, 8, cross=0.3, CrossProbability=0.2)
    with pytest.raises(KeyError):
        DifferentialEvolutionSolver2(3, 8, scale=0.3, ScalingFactor=0.2)
    assert DifferentialEvolutionSolver2(10, 2).nPop == 10      # NP = max(NP, dim, 4)
    assert DifferentialEvolutionSolver2(2, 1).nPop == 4
    # Step() keywords are honoured and sticky
    s.SetRandomInitialPoints([-1.] * 3, [1.] * 3)
    s.Step(models.rosen, CrossProbability=0.25)
    assert s.probability == 0.25
    s.Step()
    assert s.probability == 0.25
    # unknown strategy names fall back to Best1Bin (differential_evolution.py:682)
    s2 = DifferentialEvolutionSolver2(3, 8, engine=OracleEngine, strategy='NoSuchStrategy')
    s2.SetRandomInitialPoints([-1.] * 3, [1.] * 3)
    s2.Step(models.rosen)
    assert s2.strategy == 'Best1Bin'


def test_setter_errors_match_reference():
    s = DifferentialEvolutionSolver2(3, 8, engine=OracleEngine)
    with pytest.raises(ValueError):
        s.SetStrictRanges([0, 0, 0], [1, -1, 1])
    with pytest.raises(ValueError):
        s.SetStrictRanges([0, 0], [1, 1])
    with pytest.raises(ValueError):
        s.SetStrictRanges([0, 0, 0], [1, 1, 1], tight=False, clip=True)
    with pytest.raises(TypeError):
        s.SetConstraints(5)
    with pytest.raises(TypeError):
        s.SetPenalty(5)
    with pytest.raises(TypeError):
        s.SetConstraints(lambda x: x)      # host callables cannot run on the device
    with pytest.raises(TypeError):
        s.SetPenalty(lambda x: 0.0)
    with pytest.raises(TypeError):
        s.SetObjective(lambda x: sum(x))   # no CPU cost path
    with pytest.raises(ValueError):
        s.SetInitialPoints([1., 2.])
    with pytest.raises(ValueError):
        s.SetRandomInitialPoints([0.], [1.])
    s.SetConstraints(None)
    s.SetPenalty(None)


def test_maxiter_limits():
    """mystic/tests/test_solver_sanity.py:258-276: maxiter in {0,1,2}"""
    for maxiter in (0, 1, 2):
        random.seed(111)
        s = DifferentialEvolutionSolver2(2, 40, rng='reference', engine=OracleEngine)
        s.SetRandomInitialPoints([-5.] * 2, [5.] * 2)
        s.SetStrictRanges([-5.] * 2, [5.] * 2)
        s.SetEvaluationLimits(generations=maxiter)
        s.Solve(models.rosen, termination.VTR(1e-10))
        assert s.generations == maxiter
        assert 0 < s.evaluations <= (maxiter + 1) * 40
        assert s.generations == len(s._stepmon._y) - 1


def test_generation_monitor_prepend_and_null():
    from mystic_b200.monitors import Monitor, Null, VerboseMonitor
    s = DifferentialEvolutionSolver2(3, 8, engine=OracleEngine)
    s.SetRandomInitialPoints([-1.] * 3, [1.] * 3)
    s.Step(models.rosen)
    s.Step()
    m = Monitor()
    s.SetGenerationMonitor(m)
    assert len(m) == 2 and s.generations == 1
    s.SetGenerationMonitor(Null())     # Null is not allowed as a generation monitor: replaced by Monitor
    assert len(s._stepmon) == 2
    s.SetGenerationMonitor(VerboseMonitor(1000), new=True)
    assert len(s._stepmon) == 0


def test_save_and_load_solver_resumes_identically(tmp_path):
    """SetSaveFrequency / SaveSolver / LoadSolver (abstract_solver.py:975-1018, solvers.py:90-125):
    a solver reloaded from its restart file continues exactly like the one that kept running"""
    from mystic_b200 import LoadSolver

    def make():
        s = DifferentialEvolutionSolver2(4, 20, rng='philox', seed=77, engine=OracleEngine, strategy='Best1Exp')
        s.SetRandomInitialPoints([-2.] * 4, [2.] * 4)
        s.SetStrictRanges([-2.] * 4, [2.] * 4)
        s.SetTermination(termination.VTR(-1.0))
        s.SetEvaluationLimits(generations=10 ** 6)
        return s
    a = make()
    a.Step(models.rosen)
    for _ in range(5):
        a.Step()
    path = str(tmp_path / 'restart.pkl')
    a.SaveSolver(path)
    b = LoadSolver(path)
    b._engine_factory = OracleEngine
    assert b.generations == a.generations and b.evaluations == a.evaluations
    for _ in range(4):
        a.Step()
        b.Step(models.rosen)
    assert list(a.energy_history) == list(b.energy_history)
    G.assert_bits_equal(a.population, b.population, 'population after resume')
    G.assert_bits_equal(np.array(a.bestSolution), np.array(b.bestSolution), 'bestSolution after resume')
    assert a.evaluations == b.evaluations


@pytest.mark.parametrize('term', [('VTR', 1e-2), ('ChangeOverGeneration', 1e-3, 5), ('NormalizedChangeOverGeneration', 1e-2, 4),
                                  ('VTRChangeOverGeneration', 1e-3, 1e-4, 6), ('EvaluationLimits', 37, None),
                                  ('EvaluationLimits', None, 900), ('ChangeOverGeneration', 1e-30, 0)])
def test_batched_solve_stops_exactly_like_step_by_step(term):
    """Solve() with the device-side termination test (batches of generations) against the
    same solve executed one Step() at a time: same stop generation, message, histories, counters"""
    def make(batch):
        s = DifferentialEvolutionSolver2(3, 20, rng='philox', seed=4242, engine=OracleEngine, strategy='Best1Exp')
        s._batch_generations = batch
        s.SetRandomInitialPoints([0.] * 3, [2.] * 3)
        s.SetEvaluationLimits(generations=150, evaluations=2500)
        s.Solve(models.rosen, getattr(termination, term[0])(*term[1:]))
        return s
    a, b, c = make(0), make(7), make(64)
    for s in (b, c):
        assert s.generations == a.generations and s.evaluations == a.evaluations
        assert list(s.energy_history) == list(a.energy_history)
        assert s.Terminated(info=True) == a.Terminated(info=True)
        G.assert_bits_equal(s.population, a.population, 'population')
        assert s._stepmon._info == a._stepmon._info
        assert [list(x) for x in s.solution_history] == [list(x) for x in a.solution_history]


def test_logging_monitors_write_the_reference_layout(tmp_path, capsys):
    """LoggingMonitor / VerboseLoggingMonitor / VerboseMonitor: header, line layout and screen text of
    mystic/monitors.py:373-585 (checked against literal expectations here, against the reference's
    own classes in tests/test_reference_objects.py)"""
    import pickle
    from mystic_b200.monitors import LoggingMonitor, VerboseLoggingMonitor, VerboseMonitor
    log = tmp_path / 'log.txt'
    m = LoggingMonitor(2, str(log), new=True, info='run A')
    for i in range(5):
        m([1.0 * i, 2.0], 0.5 * i, id=(7 if i == 4 else None))
    m.info('done')
    lines = log.read_text().splitlines()
    assert lines[0].startswith('# ') and lines[1] == '# run A' and lines[2] == '# ___#___  __ChiSquare__  __params__'
    assert lines[3:] == ['  (0,)     0.0   [0.0, 2.0]', '  (2,)     1.0   [2.0, 2.0]', '  (4, 7)     2.0   [4.0, 2.0]', '# done']
    assert len(m) == 5 and m.y == [0.0, 0.5, 1.0, 1.5, 2.0]
    m2 = pickle.loads(pickle.dumps(m))                    # restart files carry the monitor
    assert m2.y == m.y and m2._filename == str(log)
    capsys.readouterr()
    v = VerboseMonitor(2, 4)
    for i in range(5):
        v([1.0 * i], 0.25 * i, id=(3 if i == 2 else None))
    out = capsys.readouterr().out.splitlines()
    assert out == ['Generation 0 has ChiSquare: 0.0', 'Generation 0 has fit parameters:', ' [0.0]',
                   '[id: 3] Generation 2 has ChiSquare: 0.5', 'Generation 4 has ChiSquare: 1.0',
                   'Generation 4 has fit parameters:', ' [4.0]']
    vl = VerboseLoggingMonitor(1, 2, filename=str(tmp_path / 'v.txt'), new=True, label='cost')
    for i in range(3):
        vl([0.5], 1.0 + i)
    assert capsys.readouterr().out.splitlines() == ['Generation 0 has cost: 1.0', 'Generation 2 has cost: 3.0']
    assert (tmp_path / 'v.txt').read_text().splitlines()[1:] == ['# ___#___  __cost__  __params__', '  (0,)     1.0   [0.5]',
                                                                '  (1,)     2.0   [0.5]', '  (2,)     3.0   [0.5]']


def test_collapse_interface_and_deepcopy():
    """abstract_solver.py:783-892, 1255-1257: a device solver never collapses (no collapse conditions on the
    data-parallel path), `_is_new` follows the counters, copy.deepcopy goes through the restart state"""
    import copy
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    s = DifferentialEvolutionSolver2(3, 12, rng='philox', seed=3, engine=OracleEngine)
    s.SetRandomInitialPoints([-2.] * 3, [2.] * 3)
    s.SetTermination(termination.VTR(1e-9))
    s.SetEvaluationLimits(generations=5)
    assert s._is_new() is False and s.Collapsed() is False and s.Collapsed(info=True) == {} and s.Collapse() == {}
    s.Solve(models.rosen)
    assert s._is_new() is True and s.generations == 5
    t = copy.deepcopy(s)
    assert t.generations == 5 and t.bestEnergy == s.bestEnergy and np.array_equal(t.population, s.population)
    t.SetEvaluationLimits(generations=8)
    s.SetEvaluationLimits(generations=8)
    t.Solve()
    s.Solve()
    assert t.generations == s.generations == 8 and t.bestEnergy == s.bestEnergy
    G.assert_bits_equal(t.population, s.population, 'a copy continues like the original')
