"""north_star: "The Philox mode is validated by matching VTR success rate and generations-to-converge
over 100 seeds."  Reference statistics: tests/golden/stats_reference.json, recorded from the
unmodified reference by oracle/record_statistics.py (seeds 0..99, stdlib Mersenne Twister).  Here the
drop-in solver runs the same problems with the in-kernel Philox RNG for 100 seeds."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(p, seed):
    from mystic_b200 import DifferentialEvolutionSolver2, models, strategy, termination
    s = DifferentialEvolutionSolver2(p['dim'], p['NP'], rng='philox', seed=1000003 * seed + 17,
                                     CrossProbability=p['CR'], ScalingFactor=p['F'])
    s.SetRandomInitialPoints(list(p['init'][0]), list(p['init'][1]))
    if p['bounds']:
        lo, hi, tight = p['bounds']
        s.SetStrictRanges(list(lo), list(hi), tight=True) if tight else s.SetStrictRanges(list(lo), list(hi))
    s.SetEvaluationLimits(generations=p['maxgen'])
    s.Solve(getattr(models, p['cost']), termination.VTR(p['vtr']), strategy=getattr(strategy, p['strategy']))
    return bool(abs(s.bestEnergy) <= p['vtr']), int(s.generations)


@pytest.mark.parametrize('name', ['rosen3_best1exp', 'rosen5_best1bin_reject', 'zimmermann_rand1bin',
                                  'griewangk4_rand1exp_reject'])
def test_philox_convergence_statistics_match_reference(name):
    ref = json.load(open(os.path.join(HERE, 'golden', 'stats_reference.json')))[name]
    p = ref['problem']
    rows = [_run(p, seed) for seed in range(100)]
    ok = np.array([r[0] for r in rows])
    gens = np.array([r[1] for r in rows], dtype=float)
    rok = np.array(ref['success'])
    rgens = np.array(ref['generations'], dtype=float)
    # success rate: two binomial samples of n=100 -> allow 3 sigma of the difference
    p_pool = (ok.mean() + rok.mean()) / 2.0
    sigma = np.sqrt(max(2.0 * p_pool * (1.0 - p_pool) / 100.0, 1e-4))
    assert abs(ok.mean() - rok.mean()) <= 3.0 * sigma + 0.02, (ok.mean(), rok.mean())
    # generations to converge (successful runs): medians within 15 %, and a two-sample
    # Kolmogorov-Smirnov distance below the 0.1 % critical value
    a, b = np.sort(gens[ok]), np.sort(rgens[rok])
    assert abs(np.median(a) - np.median(b)) <= 0.15 * np.median(b) + 2, (np.median(a), np.median(b))
    grid = np.concatenate([a, b])
    ks = np.max(np.abs(np.searchsorted(a, grid, side='right') / len(a) - np.searchsorted(b, grid, side='right') / len(b)))
    crit = 1.95 * np.sqrt((len(a) + len(b)) / (len(a) * len(b)))
    assert ks <= crit, (ks, crit)
    print('%s: success %d/100 (reference %d/100), median generations %g (reference %g), KS %.3f (crit %.3f)'
          % (name, ok.sum(), rok.sum(), np.median(a), np.median(b), ks, crit))


# ---- the sharded solver's deviation from the reference: donors are drawn from the candidate's own shard ----------
def _sharded_batch(p, world, nseeds=100):
    """100 runs of a population sharded over `world` ranks -- shard-local donors, best member shared every generation
    (mb2_set_island_groups: the same kernels and the same merge rule as the multi-GPU path, which
    tests/dist_gpu_check.py compares with this very emulation on real ranks) -- advanced together by one launch per
    generation; -> (success[nseeds], generations[nseeds])"""
    from mystic_b200 import models, strategy as st
    from mystic_b200.engine import DeviceEngine, BOUNDS_NONE, BOUNDS_REJECT, BOUNDS_CLAMP
    NP, D = p['NP'], p['dim']
    assert NP % world == 0
    eng = DeviceEngine(nseeds * NP, D)
    cost = getattr(models, p['cost'])
    eng.set_cost(cost.device_id(False, D), cost.params(()))
    if p['bounds']:
        lo, hi, tight = p['bounds']
        eng.set_bounds(BOUNDS_CLAMP if tight else BOUNDS_REJECT, np.asarray(lo, float), np.asarray(hi, float))
    else:
        eng.set_bounds(BOUNDS_NONE)
    (sid, k, _), _name = st.strategy_id(getattr(st, p['strategy']))
    eng.set_strategy(sid, p['F'], p['CR'])
    eng.set_rng_philox(0xD15C0 + world)
    eng.init_population(0xD15C0 + world, np.asarray(p['init'][0], float), np.asarray(p['init'][1], float))
    eng.set_islands(nseeds * world, share_best_within=world)
    done = np.zeros(nseeds, dtype=bool)
    gens = np.full(nseeds, p['maxgen'])
    for g in range(p['maxgen'] + 1):
        if g == 0:
            eng.eval_initial()
        else:
            eng.step()
        be = eng.read_islands()[0].reshape(nseeds, world)
        assert (be == be[:, :1]).all()                     # every shard of a run holds the run's best
        hit = (np.abs(be[:, 0]) <= p['vtr']) & ~done
        gens[hit] = g
        done |= hit
        if done.all():
            break
    eng.close()
    return done, gens.astype(float)


def _compare(name, ok, gens, ref, med_tol, ks_factor, rate_slack):
    rok = np.array(ref['success'])
    rgens = np.array(ref['generations'], dtype=float)
    p_pool = (ok.mean() + rok.mean()) / 2.0
    sigma = np.sqrt(max(2.0 * p_pool * (1.0 - p_pool) / 100.0, 1e-4))
    a, b = np.sort(gens[ok]), np.sort(rgens[rok])
    grid = np.concatenate([a, b])
    ks = np.max(np.abs(np.searchsorted(a, grid, side='right') / max(len(a), 1) - np.searchsorted(b, grid, side='right') / len(b))) \
        if len(a) else 1.0
    crit = 1.95 * np.sqrt((len(a) + len(b)) / (max(len(a), 1) * len(b)))
    print('%s: success %d/100 (reference %d/100), median generations %g (reference %g), KS %.3f (crit %.3f)'
          % (name, ok.sum(), rok.sum(), np.median(a) if len(a) else float('nan'), np.median(b), ks, crit))
    assert ok.mean() >= rok.mean() - 3.0 * sigma - rate_slack, (ok.mean(), rok.mean())
    assert abs(np.median(a) - np.median(b)) <= med_tol * np.median(b) + 2, (np.median(a), np.median(b))
    assert ks <= ks_factor * crit, (ks, crit)


@pytest.mark.parametrize('name', ['rosen3_best1exp', 'rosen5_best1bin_reject', 'zimmermann_rand1bin',
                                  'griewangk4_rand1exp_reject'])
def test_one_shard_batch_reproduces_the_single_solver_statistics(name):
    """world = 1 through the batched path: the bars of the single-solver test above"""
    ref = json.load(open(os.path.join(HERE, 'golden', 'stats_reference.json')))[name]
    ok, gens = _sharded_batch(ref['problem'], 1)
    _compare(name + ' x1', ok, gens, ref, 0.15, 1.0, 0.02)


@pytest.mark.parametrize('name,world', [('rosen3_best1exp', 2), ('rosen5_best1bin_reject', 2)])
def test_sharded_best_strategies_match_the_reference_statistics(name, world):
    """Best1Exp / Best1Bin: every shard follows the GLOBAL best member, so a population split over two ranks (20-25
    members each) converges like the reference's unsharded run: the single-solver bars hold (B200: medians 36 vs 39
    and 97 vs 102 generations, KS 0.25 / 0.28 against critical values 0.28 / 0.29)"""
    ref = json.load(open(os.path.join(HERE, 'golden', 'stats_reference.json')))[name]
    ok, gens = _sharded_batch(ref['problem'], world)
    _compare('%s x%d' % (name, world), ok, gens, ref, 0.15, 1.0, 0.02)


@pytest.mark.parametrize('name,world', [('zimmermann_rand1bin', 2), ('griewangk4_rand1exp_reject', 2), ('rosen3_best1exp', 4),
                                        ('rosen5_best1bin_reject', 5), ('zimmermann_rand1bin', 8),
                                        ('griewangk4_rand1exp_reject', 10), ('rosen3_best1exp', 8), ('rosen5_best1bin_reject', 10)])
def test_sharded_deviation_is_reported_and_bounded(name, world):
    """Where the sharded path departs from the reference.  Rand1* never reads the best member, so
    with shard-local donors its shards are INDEPENDENT smaller populations whose best is reported: fewer generations to
    the target (a population of NP/world members converges faster) at a success rate that falls as the shards shrink.
    Best* over 4-5 shards converges ~10 % faster than the reference (the donor differences come from 10 rows).
    Measured on B200 (100 runs each; reference in brackets): zimmermann Rand1Bin x2 100/100, median 52 (54); griewangk
    Rand1Exp x2 100/100, 321 (372); rosen3 Best1Exp x4 100/100, 35 (39); rosen5 Best1Bin x5 82/100 (85), 90 (102);
    and with shards of 5-6 members: zimmermann x8 89/100, griewangk x10 0/100 (six members cannot carry Rand1Exp),
    rosen3 x8 100/100, rosen5 x10 74/100 (85).  Asserted is what a user can rely on: the run is never slower than the
    reference's median by more than 15 %, and with shards of >= 10 members it succeeds as often as the reference does
    (3 sigma).  BASELINE config 5 has 2 M members per shard."""
    ref = json.load(open(os.path.join(HERE, 'golden', 'stats_reference.json')))[name]
    p = ref['problem']
    ok, gens = _sharded_batch(p, world)
    rok = np.array(ref['success'])
    rmed = np.median(np.array(ref['generations'], dtype=float)[rok])
    a = gens[ok]
    print('%s x%d shards of %d: success %d/100 (reference %d/100), median generations %s (reference %g)'
          % (name, world, p['NP'] // world, ok.sum(), rok.sum(), np.median(a) if len(a) else None, rmed))
    if len(a):
        assert np.median(a) <= 1.15 * rmed + 2
    if p['NP'] // world >= 10:
        p_pool = (ok.mean() + rok.mean()) / 2.0
        sigma = np.sqrt(max(2.0 * p_pool * (1.0 - p_pool) / 100.0, 1e-4))
        assert ok.mean() >= rok.mean() - 3.0 * sigma - 0.02, (ok.mean(), rok.mean())
