#!/usr/bin/env python
"""Time the UNMODIFIED reference (mystic @ /root/reference) on the BASELINE problem shapes, in the
build container (TEST INFRASTRUCTURE ONLY; the reference cannot travel to the GPU box).
SURVEY.md 8(d): serial python_map, 3 timed generations after generation 0, NP in {max(D,1024), 4096};
cand-evals/s = NP * generations / wall.  Writes profiles/r1_reference_python_timings.json."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, '_shim'))
sys.path.insert(1, '/root/reference')
sys.path.insert(2, REPO)


def run(name, dim, NP, cost, strategy, CR, F, lo, hi, tight, gens, penalty=None, extra=None):
    import mystic.models as mm
    import mystic.strategy as ms
    from mystic.solvers import DifferentialEvolutionSolver2
    from mystic.termination import VTR
    from mystic.tools import random_seed
    from oracle.host_callables import build_reference_penalty
    if not hasattr(mm, cost):
        import importlib
        importlib.import_module('mystic.models.poly')
        fn = getattr(sys.modules['mystic.models.poly'], cost)
    else:
        fn = getattr(mm, cost)
    random_seed(123)
    s = DifferentialEvolutionSolver2(dim, NP, CrossProbability=CR, ScalingFactor=F)
    s.SetRandomInitialPoints([lo] * dim, [hi] * dim)
    if tight is not None:
        s.SetStrictRanges([lo] * dim, [hi] * dim, tight=True) if tight else s.SetStrictRanges([lo] * dim, [hi] * dim)
    if penalty:
        s.SetPenalty(build_reference_penalty(penalty))
    s.SetTermination(VTR(-1.0))
    s.SetEvaluationLimits(generations=10 ** 9, evaluations=10 ** 12)
    s.Step(fn, ExtraArgs=extra, strategy=getattr(ms, strategy))   # generation 0
    t0 = time.perf_counter()
    for _ in range(gens):
        s.Step()
    dt = time.perf_counter() - t0
    out = dict(problem=name, dim=dim, NP=NP, generations=gens, seconds=dt, cand_evals_per_s=NP * gens / dt)
    print(json.dumps(out), flush=True)
    return out


def main():
    rows = []
    global run
    _run = run

    def run(*a, **k):
        try:
            return _run(*a, **k)
        except Exception as e:   # e.g. RecursionError: 2*D nested symbolic clamp functions at D=1024
            out = dict(problem=a[0], dim=a[1], NP=a[2], error='%s: %s' % (type(e).__name__, str(e)[:80]))
            print(json.dumps(out), flush=True)
            return out
    rows.append(run('C2 rosen 1024-D Best1Bin clamp', 1024, 1024, 'rosen', 'Best1Bin', 0.9, 0.8, -5., 5., True, 2))
    rows.append(run('C2 rosen 1024-D Best1Bin reject', 1024, 1024, 'rosen', 'Best1Bin', 0.9, 0.8, -5., 5., False, 3))
    rows.append(run('C2 rosen 1024-D Best1Bin reject', 1024, 4096, 'rosen', 'Best1Bin', 0.9, 0.8, -5., 5., False, 2))
    rows.append(run('C5 griewangk 256-D Best1Bin reject', 256, 1024, 'griewangk', 'Best1Bin', 0.9, 0.8, -400., 400., False, 3))
    rows.append(run('C5 griewangk 256-D Best1Bin reject', 256, 4096, 'griewangk', 'Best1Bin', 0.9, 0.8, -400., 400., False, 2))
    rows.append(run('C4 corana 64-D Rand1Exp clamp + penalty', 64, 1024, 'corana', 'Rand1Exp', 0.9, 0.8, -1000., 1000., True, 2,
                    penalty=[dict(ptype='quadratic_inequality', G=[1.] * 64, h=100., k=100, hpow=5)]))
    rows.append(run('C3 chebyshev8cost M=4096 Best1Exp', 9, 64, 'chebyshev8cost', 'Best1Exp', 1.0, 0.9, -100., 100., None, 2,
                    extra=(4096,)))
    rows.append(run('C1 rosen 3-D NP=40 Best1Exp', 3, 40, 'rosen', 'Best1Exp', 0.9, 0.8, 0., 2., None, 200))
    import platform
    meta = dict(host='build container, %s' % platform.processor(), cores_used=1, python=sys.version.split()[0],
                note='serial python_map; trial generation is a per-dimension Python loop (strategy.py) and '
                     'get_random_candidates is O(NP) per candidate, so throughput falls with NP')
    with open(os.path.join(REPO, 'profiles', 'r1_reference_python_timings.json'), 'w') as f:
        json.dump(dict(meta=meta, rows=rows), f, indent=1)


if __name__ == '__main__':
    main()
