#!/usr/bin/env python
"""Convergence statistics of the UNMODIFIED reference over 100 seeds (TEST INFRASTRUCTURE ONLY).

north_star: "The Philox mode is validated by matching VTR success rate and generations-to-converge
over 100 seeds."  This script runs mystic.DifferentialEvolutionSolver2 (from /root/reference) for
seeds 0..99 on a few small problems and stores, per seed, whether VTR was reached within the
generation limit and after how many generations -> tests/golden/stats_reference.json.
"""
import json
import multiprocessing
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, '_shim'))
sys.path.insert(1, '/root/reference')

PROBLEMS = {
    'rosen3_best1exp': dict(dim=3, NP=40, cost='rosen', strategy='Best1Exp', CR=0.9, F=0.8,
                            init=([0.] * 3, [2.] * 3), bounds=None, vtr=1e-4, maxgen=1000),
    'rosen5_best1bin_reject': dict(dim=5, NP=50, cost='rosen', strategy='Best1Bin', CR=0.9, F=0.8,
                                   init=([-5.] * 5, [5.] * 5), bounds=([-5.] * 5, [5.] * 5, False), vtr=1e-4,
                                   maxgen=1500),
    'zimmermann_rand1bin': dict(dim=2, NP=40, cost='zimmermann', strategy='Rand1Bin', CR=0.9, F=0.8,
                                init=([0.] * 2, [10.] * 2), bounds=([0.] * 2, [10.] * 2, False), vtr=5e-3,
                                maxgen=1000),
    # (tight=True bounds are avoided here: the reference's symbolic clamp runs ~176 evals/s)
    'griewangk4_rand1exp_reject': dict(dim=4, NP=60, cost='griewangk', strategy='Rand1Exp', CR=0.9, F=0.5,
                                       init=([-50.] * 4, [50.] * 4), bounds=([-50.] * 4, [50.] * 4, False), vtr=1e-3,
                                       maxgen=1500),
}


def run(p, seed):
    import mystic.models as mm
    import mystic.strategy as ms
    from mystic.solvers import DifferentialEvolutionSolver2
    from mystic.termination import VTR
    from mystic.tools import random_seed
    random_seed(seed)
    s = DifferentialEvolutionSolver2(p['dim'], p['NP'], CrossProbability=p['CR'], ScalingFactor=p['F'])
    s.SetRandomInitialPoints(list(p['init'][0]), list(p['init'][1]))
    if p['bounds']:
        lo, hi, tight = p['bounds']
        s.SetStrictRanges(list(lo), list(hi), tight=True) if tight else s.SetStrictRanges(list(lo), list(hi))
    s.SetEvaluationLimits(generations=p['maxgen'])
    s.Solve(getattr(mm, p['cost']), VTR(p['vtr']), strategy=getattr(ms, p['strategy']))
    ok = bool(abs(s.bestEnergy) <= p['vtr'])
    return ok, int(s.generations), float(s.bestEnergy)


def main():
    out = {}
    for name, p in PROBLEMS.items():
        with multiprocessing.Pool(8) as pool:
            rows = pool.starmap(run, [(p, seed) for seed in range(100)])
        out[name] = dict(problem=p, success=[r[0] for r in rows], generations=[r[1] for r in rows],
                         best=[r[2] for r in rows])
        gens = sorted(r[1] for r in rows if r[0])
        print('%-28s success %3d/100  median gens %s' % (name, sum(r[0] for r in rows),
                                                         gens[len(gens) // 2] if gens else None), flush=True)
    with open(os.path.join(REPO, 'tests', 'golden', 'stats_reference.json'), 'w') as f:
        json.dump(out, f)


if __name__ == '__main__':
    main()
