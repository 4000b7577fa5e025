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
