#!/usr/bin/env python
"""Times the UNMODIFIED reference on the host cores: mystic's DifferentialEvolutionSolver2 imported
from baseline/_ref (installed by baseline/install_reference.sh; klepto comes from oracle/_shim), run
serially and with a `multiprocess.Pool(nproc).map` installed through SetMapper, the way
/root/reference/examples/mpmap_desolve.py:74-81 does (pathos' ProcessPool wraps the same
multiprocess.Pool; pathos itself is not in this image).  TEST INFRASTRUCTURE: bench.py's CPU legs run
this file in a subprocess and copy its one JSON line into `cpu_baseline.reference_python`.

    python oracle/reference_python.py --workload c2 --np 1024 4096 --generations 2
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.path.join(REPO, 'baseline', '_ref')


def run(w, NP, gens, pool):
    import mystic.models as mm
    import mystic.strategy as ms
    from mystic.solvers import DifferentialEvolutionSolver2
    from mystic.termination import VTR
    from mystic.tools import random_seed
    from oracle.host_callables import build_reference_penalty
    dim = w['dim']
    if hasattr(mm, w['cost']):
        fn = getattr(mm, w['cost'])
    else:
        import importlib
        fn = getattr(importlib.import_module('mystic.models.poly'), w['cost'])
    random_seed(123)
    s = DifferentialEvolutionSolver2(dim, NP)
    if pool is not None:
        s.SetMapper(pool.map)
    s.SetRandomInitialPoints([w['lo']] * dim, [w['hi']] * dim)
    note = ''
    if w['bounds'] == 'clamp' and dim <= 256:
        s.SetStrictRanges([w['lo']] * dim, [w['hi']] * dim, tight=True)
    elif w['bounds'] is not None:
        # tight=True nests 2*D symbolic clamp functions: RecursionError in the reference at D = 1024
        s.SetStrictRanges([w['lo']] * dim, [w['hi']] * dim)
        if w['bounds'] == 'clamp':
            note = 'reject-mode bounds (the reference raises RecursionError with tight=True at this dim)'
    if w.get('penalty'):
        p = w['penalty']
        s.SetPenalty(build_reference_penalty([dict(ptype='quadratic_inequality', G=[p['G']] * dim, h=p['h'], k=p['k'],
                                                   hpow=p['hpow'])]))
    s.SetTermination(VTR(-1.0))
    s.SetEvaluationLimits(generations=10 ** 9, evaluations=10 ** 12)
    extra = tuple(w['extra']) or None
    s.Step(fn, ExtraArgs=extra, strategy=getattr(ms, w['strategy']), CrossProbability=w['CR'], ScalingFactor=w['F'])
    t0 = time.perf_counter()
    for _ in range(gens):
        s.Step()
    dt = time.perf_counter() - t0
    return dict(np=NP, generations=gens, seconds=dt, cand_evals_per_s=NP * gens / dt, note=note)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--np', type=int, nargs='+', default=[1024, 4096])
    ap.add_argument('--generations', type=int, default=2)
    ap.add_argument('--procs', type=int, default=0)
    ap.add_argument('--pool-np', type=int, nargs='*', default=[64],
                    help='population sizes of the Pool.map runs (the rows handed to map are lists of numpy.float64 '
                         'scalars, which multiprocess/dill pickles one object at a time: ~0.1 s per 1024-D candidate)')
    ap.add_argument('--pool-generations', type=int, default=1)
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, 'mystic')):
        print(json.dumps({'unavailable': 'baseline/_ref/mystic is missing (run baseline/install_reference.sh)'}))
        return
    sys.path[:0] = [os.path.join(HERE, '_shim'), REF, REPO]
    sys.setrecursionlimit(20000)
    import bench
    w = dict(bench.WORKLOADS[args.workload])
    nproc = args.procs or len(os.sched_getaffinity(0))
    out = {'what': "unmodified mystic DifferentialEvolutionSolver2 (baseline/_ref), %s" % w['desc'],
           'cores_available': len(os.sched_getaffinity(0)), 'pool_processes': nproc, 'serial': [], 'pool_map': []}
    try:
        import mystic
        out['mystic_version'] = getattr(mystic, '__version__', '?')
        for NP in args.np:
            out['serial'].append(run(w, NP, args.generations, None))
        from multiprocess import Pool
        pool = Pool(nproc)
        try:
            for NP in args.pool_np:
                out['pool_map'].append(run(w, NP, args.pool_generations, pool))
        finally:
            pool.close()
            pool.join()
        out['best_cand_evals_per_s'] = max(r['cand_evals_per_s'] for r in out['serial'] + out['pool_map'])
    except Exception as e:  # reported, not fatal: the port's number still stands beside it
        out['error'] = '%s: %s' % (type(e).__name__, str(e)[:200])
    print(json.dumps(out))


if __name__ == '__main__':
    main()
