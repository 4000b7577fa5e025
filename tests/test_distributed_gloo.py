"""N > 1 host path on CPU: two gloo ranks, population sharded by candidate, per-generation
all-gather of the shard bests.  The generation step is the test-only OracleEngine; what is under
test is mystic_b200.solver's sharding, exchange and bookkeeping (the CUDA engine implements the
same pack/merge rule in pack_best_kernel / merge_best_kernel)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NP, D, GENS, SEED = 50, 6, 8, 0xABCDEF


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, strategy, out):
    sys.path.insert(0, REPO)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    from tests._oracle_engine import OracleEngine
    s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=SEED, distributed=True, engine=OracleEngine,
                                     strategy=strategy)
    s.SetRandomInitialPoints([-3.] * D, [3.] * D)
    s.SetStrictRanges([-3.] * D, [3.] * D)
    s.SetTermination(termination.VTR(-1.0))
    s.SetEvaluationLimits(generations=10 ** 6)
    hist = []
    s.Step(models.rosen)
    for _ in range(GENS):
        s.Step()
    out[rank] = dict(energy_history=list(s.energy_history), best=np.array(s.bestSolution), evals=s.evaluations,
                     pop=s.population, popE=s.popEnergy, offset=s._offset, np_local=s._np_local,
                     last=s._last)
    dist.destroy_process_group()


def _run(strategy):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), strategy, out), nprocs=2, join=True)
    return out[0], out[1]


def _emulate(strategy):
    """the same two shards in one process: step each shard, merge by hand"""
    sys.path.insert(0, REPO)
    from oracle import de_oracle as orc
    from tests._oracle_engine import OracleEngine
    sizes = [NP // 2 + (1 if r < NP % 2 else 0) for r in range(2)]
    offs = [0, sizes[0]]
    engs = []
    for r in range(2):
        e = OracleEngine(sizes[r], D, NP, offs[r])
        e.set_cost(0)
        e.set_bounds(orc.BOUNDS_REJECT, [-3.] * D, [3.] * D)
        e.set_rng_philox(SEED)
        e.init_population(SEED, [-3.] * D, [3.] * D)
        engs.append(e)
    hist, evals = [], 0
    for g in range(GENS + 1):
        for e in engs:
            if g == 0:
                e.eval_initial()
            else:
                e.set_strategy(orc.STRATEGIES[strategy][0], 0.8, 0.9)
                e.step()
        recs = torch.stack([e.pack_best().clone() for e in engs])
        for e in engs:
            e.merge_best(recs)
        r0 = engs[0].read_result()
        r1 = engs[1].read_result()
        assert r0[0] == r1[0] and r0[1] == r1[1]
        hist.append(r0[0])
        evals += NP - r0[2]
    return hist, evals, engs


def test_two_ranks_agree_and_match_single_process_emulation():
    for strategy in ('Best1Bin', 'Rand1Exp'):
        a, b = _run(strategy)
        assert a['energy_history'] == b['energy_history']
        assert np.array_equal(a['best'], b['best'])
        assert a['evals'] == b['evals']
        assert (a['offset'], a['np_local'], b['offset'], b['np_local']) == (0, 25, 25, 25)
        # global best = best member over both shards
        allE = np.concatenate([a['popE'], b['popE']])
        allX = np.concatenate([a['pop'], b['pop']])
        assert a['energy_history'][-1] == allE.min()
        assert np.array_equal(a['best'], allX[int(np.argmin(allE))])
        # monotone, and identical to stepping the two shards in one process
        assert all(x >= y for x, y in zip(a['energy_history'], a['energy_history'][1:]))
        hist, evals, engs = _emulate(strategy)
        assert hist == a['energy_history']
        assert evals == a['evals']
        assert np.array_equal(engs[0].population(), a['pop']) and np.array_equal(engs[1].population(), b['pop'])


def _worker_unseeded(rank, world, port, out):
    sys.path.insert(0, REPO)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import random
    from mystic_b200 import DifferentialEvolutionSolver2, models, termination
    from tests._oracle_engine import OracleEngine
    random.seed(1000 + rank)        # every rank's private generator differs
    s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=None, distributed=True, engine=OracleEngine, strategy='Best1Bin')
    s.SetRandomInitialPoints([-3.] * D, [3.] * D)
    s.SetStrictRanges([-3.] * D, [3.] * D, tight=True)
    s.SetTermination(termination.VTR(-1.0))
    s.SetEvaluationLimits(generations=10 ** 6)
    s.Step(models.rosen)
    for _ in range(3):
        s.Step()
    # mid-run SetStrictRanges re-decorates: the member kept (clipped) is the GLOBAL best, which one shard holds
    gbest = s._last[2]
    s.SetStrictRanges([-1.5] * D, [1.5] * D, tight=True)
    before = s.population.copy()
    s.Step()
    out[rank] = dict(seed=s._seed, gbest=gbest, offset=s._offset, np_local=s._np_local, before=before,
                     engine_seed=s._engine.seed)
    dist.destroy_process_group()


def test_unseeded_sharded_run_shares_one_key():
    """with seed=None every rank drew its own Philox key; rank 0's is now broadcast"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_unseeded, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a['seed'] == b['seed'] == a['engine_seed'] == b['engine_seed'] and a['gbest'] == b['gbest'] >= 0
