"""Batch of independent solvers in one launch (mb2_set_islands): island i of a batched
engine must be, bit for bit, what a stand-alone engine computes for shard i of the same rows
(same Philox counters: they are keyed by the global row index; donors among the shard's own rows)."""
import numpy as np
import pytest

from oracle import de_oracle as orc

pytestmark = pytest.mark.gpu

CASES = [
    # dim, rows per island, islands, strategy, cost, bounds (lo, hi, mode)
    (8, 24, 5, 'Best1Exp', 'rosen', None),
    (64, 40, 7, 'Best1Bin', 'griewangk', (-400., 400., orc.BOUNDS_CLAMP)),
    (12, 30, 4, 'RandToBest1Exp', 'rastrigin', (-5.12, 5.12, orc.BOUNDS_REJECT)),
    (7, 28, 3, 'Rand1Exp', 'lorentzian', (0.05, 200., orc.BOUNDS_CLAMP)),
    (2, 9, 16, 'Best2Exp', 'zimmermann', (0., 5., orc.BOUNDS_REJECT)),
    (100, 33, 2, 'Rand2Exp', 'sphere', None),
]


def _cost(name, dim):
    from mystic_b200 import models
    if name == 'lorentzian':
        bins = np.arange(0.05, 3.0, 0.1)
        return models.lorentzian.CostFactory([1., 45., -10., 20., 1., 0.1, 120.], bins)
    return getattr(models, name)


def _configure(eng, cost, strat, bounds, dim, seed):
    eng.set_cost(cost.cost_id, cost.params(()))
    if bounds is None:
        eng.set_bounds(orc.BOUNDS_NONE)
    else:
        eng.set_bounds(bounds[2], np.full(dim, bounds[0]), np.full(dim, bounds[1]))
    eng.set_strategy(orc.STRATEGIES[strat][0], 0.8, 0.9)
    eng.set_rng_philox(seed)


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s-%s-D%d-x%d' % (c[3], c[4], c[0], c[2]))
def test_islands_equal_independent_shards(case, monkeypatch):
    from mystic_b200.engine import DeviceEngine
    monkeypatch.setenv('MB2_SMALL_KERNEL', '1')   # the stand-alone shards run the same kernel family: identical energies
    dim, inp, S, strat, cname, bounds = case
    total, seed = inp * S, 0x151A + dim
    cost = _cost(cname, dim)
    ilo = np.full(dim, -3.0 if bounds is None else bounds[0] * 1.2 if bounds[0] < 0 else bounds[0])
    ihi = np.full(dim, 3.0 if bounds is None else bounds[1] * 1.2)
    big = DeviceEngine(total, dim, keep_trials=True)
    _configure(big, cost, strat, bounds, dim, seed)
    big.set_islands(S)
    big.init_population(seed, ilo, ihi)
    shards = []
    for i in range(S):
        e = DeviceEngine(inp, dim, total, i * inp, keep_trials=True)
        _configure(e, cost, strat, bounds, dim, seed)
        e.init_population(seed, ilo, ihi)
        shards.append(e)
    for g in range(6):
        if g == 0:
            big.eval_initial()
            [e.eval_initial() for e in shards]
        else:
            big.step()
            [e.step() for e in shards]
        be, bi, ninf, nacc, gen, bx = big.read_islands()
        pop, en = big.population(), big.energies()
        trial, trial_e = big.trials()
        assert gen == g
        for i, e in enumerate(shards):
            r = e.read_result()
            sl = slice(i * inp, (i + 1) * inp)
            tag = 'generation %d island %d' % (g, i)
            assert np.array_equal(pop[sl], e.population()), tag
            assert np.array_equal(en[sl], e.energies()), tag
            assert np.array_equal(trial[sl], e.trials()[0]) and np.array_equal(trial_e[sl], e.trials()[1], equal_nan=True), tag
            assert (be[i], bi[i], ninf[i], nacc[i]) == (r[0], r[1], r[2], r[3]), (tag, be[i], bi[i], r[:4])
            assert np.array_equal(bx[i], r[5]), tag
    big.close()
    [e.close() for e in shards]


def test_islands_reject_what_they_cannot_do():
    from mystic_b200 import models
    from mystic_b200._native import NativeError
    from mystic_b200.engine import DeviceEngine
    eng = DeviceEngine(30, 4)
    eng.set_cost(models.rosen.cost_id, [])
    eng.set_strategy(orc.STRATEGIES['Rand2Exp'][0], 0.8, 0.9)
    with pytest.raises(NativeError):
        eng.set_islands(7)            # 30 rows do not split into 7 populations
    with pytest.raises(NativeError):
        eng.set_islands(6)            # 5 members cannot supply 5 distinct donors besides the candidate
    eng.set_islands(3)
    eng.close()
    wide = DeviceEngine(40, 4000)
    wide.set_cost(models.rosen.cost_id, [])
    wide.set_islands(2)
    wide.init_population(1, np.full(4000, -1.), np.full(4000, 1.))
    with pytest.raises(NativeError):
        wide.eval_initial()           # the best member and a warp's rows no longer fit in shared memory
    wide.close()


@pytest.mark.parametrize('D,strat,cost', [(200, 'Best1Exp', 'rosen'), (1024, 'Best1Bin', 'rosen'), (300, 'Rand1Exp', 'griewangk'),
                                          (130, 'RandToBest1Bin', 'corana')])
def test_islands_of_long_rows_equal_the_oracle_shards(D, strat, cost):
    """batched solvers of more than 128 dimensions (a warp per row): every island equals the oracle's stand-alone shard"""
    from mystic_b200 import models
    from mystic_b200.engine import DeviceEngine
    from tests._oracle_engine import OracleEngine
    S, inp, seed = 3, 12, 0x1512 + D
    lo, hi = np.full(D, -4.), np.full(D, 4.)
    engs = []
    for E in (DeviceEngine, OracleEngine):
        e = E(S * inp, D)
        e.set_cost(getattr(models, cost).cost_id)
        e.set_bounds(orc.BOUNDS_CLAMP, lo, hi)
        e.set_rng_philox(seed)
        e.init_population(seed, lo, hi)
        e.set_islands(S)
        e.eval_initial()
        engs.append(e)
    for g in range(4):
        for e in engs:
            e.set_strategy(orc.STRATEGIES[strat][0], 0.8, 0.9)
            e.step()
        rd, ro = engs[0].read_islands(), engs[1].read_islands()
        assert np.array_equal(engs[0].population(), engs[1].population()), g
        assert np.array_equal(rd[5], ro[5]) and np.array_equal(rd[2], ro[2]) and np.array_equal(rd[3], ro[3])
        scale = np.maximum(np.abs(ro[0]), (engs[1].population() ** 2).sum(axis=1).max() / 4000. + 2.0 if cost == 'griewangk' else 0.0)
        assert (np.abs(rd[0] - ro[0]) <= 1e-12 * scale).all()
    engs[0].close()


@pytest.mark.parametrize('term', [('VTR', 1e-4), ('ChangeOverGeneration', 1e-4, 8), ('EvaluationLimits', None, 1500)])
def test_device_side_termination_of_nested_solvers(term):
    """BuckshotSolver with the condition tested inside the argmin epilogue (64 generations per host round trip) against
    the per-generation host loop on the device and against the oracle engine: same stop generations, messages, members"""
    import random
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    from mystic_b200.ensemble import BuckshotSolver
    from tests._oracle_engine import OracleEngine
    runs = []
    for engine, batch in ((None, 64), (None, 0), (OracleEngine, 64)):
        random.seed(21)
        s = BuckshotSolver(4, 12, engine=engine, seed=99)
        s._batch_generations = batch
        s.SetNestedSolver(DifferentialEvolutionSolver2(4, 24, strategy='Best1Exp'))
        s.SetStrictRanges([-2.] * 4, [2.] * 4)
        s.SetEvaluationLimits(generations=150)
        s.Solve(models.rosen, getattr(termination, term[0])(*term[1:]))
        runs.append(s)
    batched, stepped, ref = runs
    for other in (stepped, ref):
        assert batched._all_iters == other._all_iters and batched._all_evals == other._all_evals
        assert batched.Terminated(info=True, all=True) == other.Terminated(info=True, all=True)
        for a, b in zip(batched._all_bestSolution, other._all_bestSolution):
            assert np.array_equal(a, b)
        assert np.allclose(batched._all_bestEnergy, other._all_bestEnergy, rtol=1e-12, atol=0)
    assert len(set(batched._all_iters)) > 1 or term[0] == 'EvaluationLimits'
    # the batched run synchronises once per 64 generations: far fewer host round trips than generations
    assert batched._engine.launch_count() < stepped._engine.launch_count() + 2 * 64 + 16


def test_buckshot_on_device_equals_the_oracle_ensemble():
    """mystic_b200.ensemble.BuckshotSolver with its nested DE2 runs on the device against the same
    class over the test-only oracle engine: same members, same stop generations, best members bit-exact"""
    import random
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    from mystic_b200.ensemble import BuckshotSolver
    from tests._oracle_engine import OracleEngine
    runs = []
    for engine in (None, OracleEngine):
        random.seed(21)
        s = BuckshotSolver(4, 12, engine=engine, seed=99)
        s.SetNestedSolver(DifferentialEvolutionSolver2(4, 24, strategy='Best1Exp'))
        s.SetStrictRanges([-2.] * 4, [2.] * 4)
        s.SetEvaluationLimits(generations=80)
        s.Solve(models.rosen, termination.VTR(1e-4))
        runs.append(s)
    dev, ref = runs
    assert dev._all_iters == ref._all_iters and dev._all_evals == ref._all_evals
    assert len(set(dev._all_iters)) > 1
    for a, b in zip(dev._all_bestSolution, ref._all_bestSolution):
        assert np.array_equal(a, b)
    assert np.allclose(dev._all_bestEnergy, ref._all_bestEnergy, rtol=1e-12, atol=0)
    assert np.array_equal(dev.bestSolution, ref.bestSolution) and dev.Terminated(info=True) == ref.Terminated(info=True)
    # the energy fill of the upload, then ONE fused launch + one epilogue per generation for all members together, in
    # batches of 64 generations (arm + status kernels per batch; the launches after the last stop copy rows through)
    nb = -(-(max(dev._all_iters) + 1) // 64)
    assert dev._engine.launch_count() == 1 + nb * (2 * 64 + 2)


def test_islands_sharing_their_best_are_a_sharded_population():
    """mb2_set_island_groups: groups of w islands whose best member is merged every generation equal the oracle's
    emulation of a population sharded over w ranks (the emulation tests/dist_gpu_check.py holds real ranks to)"""
    import numpy as np
    from mystic_b200 import models
    from mystic_b200.engine import DeviceEngine
    from oracle import de_oracle as orc
    from tests._oracle_engine import OracleEngine
    for strat, cost, D, inp, w, groups, mode in (('Best1Exp', 'rosen', 6, 10, 4, 3, orc.BOUNDS_CLAMP),
                                                 ('RandToBest1Bin', 'griewangk', 30, 12, 2, 5, orc.BOUNDS_REJECT),
                                                 ('Best1Bin', 'rosen', 64, 16, 8, 2, orc.BOUNDS_CLAMP)):
        NP, seed = inp * w * groups, 0x15A + D
        lo, hi = np.full(D, -3.), np.full(D, 3.)
        engs = []
        for E in (DeviceEngine, OracleEngine):
            e = E(NP, D)
            e.set_cost(getattr(models, cost).cost_id)
            e.set_bounds(mode, lo, hi)
            e.set_rng_philox(seed)
            e.init_population(seed, lo, hi)
            e.set_islands(w * groups, share_best_within=w)
            e.eval_initial()
            engs.append(e)
        for g in range(6):
            for e in engs:
                e.set_strategy(orc.STRATEGIES[strat][0], 0.8, 0.9)
                e.step()
            rd, ro = engs[0].read_islands(), engs[1].read_islands()
            assert np.array_equal(engs[0].population(), engs[1].population()), (strat, g)
            assert np.array_equal(rd[5], ro[5]) and np.array_equal(rd[1], ro[1]), (strat, g)
            assert np.allclose(rd[0], ro[0], rtol=1e-12, atol=0)
            assert (rd[0].reshape(groups, w) == rd[0].reshape(groups, w)[:, :1]).all()   # one best per group
        engs[0].close()
