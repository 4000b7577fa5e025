#!/usr/bin/env python
"""Multi-GPU parity check, launched by torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py

Every rank runs the drop-in solver with distributed=True on its GPU and, on the CPU, the oracle's
emulation of ALL shards (step each shard, merge the bests by the same rule).  The rank's shard
must match the emulation bit for bit (populations, best solution), energies to 1e-12."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from mystic_b200 import DifferentialEvolutionSolver2, models, termination  # noqa: E402
from oracle import de_oracle as orc  # noqa: E402
from tests._oracle_engine import OracleEngine  # noqa: E402


def emulate(world, NP, D, seed, strategy, cost_id, bounds, gens):
    base, rem = divmod(NP, world)
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    offs = [r * base + min(r, rem) for r in range(world)]
    engs = []
    for r in range(world):
        e = OracleEngine(sizes[r], D, NP, offs[r])
        e.set_cost(cost_id)
        e.set_bounds(*bounds)
        e.set_rng_philox(seed)
        e.init_population(seed, bounds[1], bounds[2])
        engs.append(e)
    hist = []
    for g in range(gens + 1):
        for e in engs:
            if g == 0:
                e.eval_initial()
            else:
                e.set_strategy(orc.STRATEGIES[strategy][0], 0.8, 0.9)
                e.step()
        recs = torch.stack([e.pack_best().clone() for e in engs])
        for e in engs:
            e.merge_best(recs)
        hist.append(engs[0].read_result()[0])
    return engs, hist


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    # MB2_DIST_ONE_GPU=1: every rank on cuda:0 with a gloo group (NCCL refuses two ranks on one
    # device) -- lets a single-GPU box exercise the peer-memory exchange between processes
    one_gpu = os.environ.get('MB2_DIST_ONE_GPU') == '1'
    if one_gpu:
        local = 0
    torch.cuda.set_device(local)
    if one_gpu:
        dist.init_process_group('gloo')
    else:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok = True
    for (strategy, cost, D, NP, mode, lo, hi) in [('Best1Bin', 'rosen', 64, 1000, orc.BOUNDS_CLAMP, -5., 5.),
                                                  ('RandToBest1Exp', 'griewangk', 30, 301, orc.BOUNDS_REJECT, -400., 400.),
                                                  ('Rand1Exp', 'rosen', 9, 77, orc.BOUNDS_CLAMP, -2., 2.),
                                                  # sharded runs of the bulk-copy kernels: short rows (de_rows_bulk.cuh) and
                                                  # long rows with the exchanged best member (de_tma_exp.cuh)
                                                  ('Rand1Exp', 'corana', 64, 2003, orc.BOUNDS_CLAMP, -1000., 1000.),
                                                  ('Best1Exp', 'rosen', 200, 303, orc.BOUNDS_CLAMP, -5., 5.),
                                                  # the headline kernel (de_tma.cuh, one-warp groups with a two-stage ring) on shards
                                                  ('Best1Bin', 'rosen', 256, 64 * world + 259, orc.BOUNDS_CLAMP, -5., 5.)]:
        seed, gens = 0xD15C0 + D, 6
        s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=seed, distributed=True, strategy=strategy)
        if os.environ.get('MB2_EXCHANGE', 'peer') == 'peer':
            assert hasattr(s, '_engine')
        s.SetRandomInitialPoints([lo] * D, [hi] * D)
        s.SetStrictRanges([lo] * D, [hi] * D, tight=(mode == orc.BOUNDS_CLAMP))
        s.SetTermination(termination.VTR(-1.0))
        s.SetEvaluationLimits(generations=10 ** 6)
        s.Step(getattr(models, cost))
        for _ in range(gens):
            s.Step()
        cid = getattr(models, cost).cost_id
        engs, hist = emulate(world, NP, D, seed, strategy, cid, (mode, np.full(D, lo), np.full(D, hi)), gens)
        mine = engs[rank]
        if os.environ.get('MB2_EXCHANGE', 'peer') == 'peer':
            assert s._engine.peer_exchange, 'the peer-memory exchange is not in use'
        same_pop = np.array_equal(s.population, mine.population())
        same_best = np.array_equal(np.array(s.bestSolution), mine.state.bestSolution)
        e_ok = np.allclose(np.array(s.energy_history), np.array(hist), rtol=1e-12, atol=0)
        print('rank %d %s/%s D=%d NP=%d: population %s, bestSolution %s, energy history %s, evaluations %d'
              % (rank, strategy, cost, D, NP, same_pop, same_best, e_ok, s.evaluations), flush=True)
        ok = ok and same_pop and same_best and e_ok
    # Solve(): with the peer exchange inside the epilogue kernel, sharded runs take the device-resident
    # multi-generation loop too (termination test on the merged best, the same on every rank); it must
    # stop at the same generation, with the same history, as stepping generation by generation
    runs = []
    for batch in (64, 0):
        D, NP = 12, 500
        s = DifferentialEvolutionSolver2(D, NP, rng='philox', seed=0xBA7C4, distributed=True, strategy='Best1Exp')
        s._batch_generations = batch
        s.SetRandomInitialPoints([-3.] * D, [3.] * D)
        s.SetStrictRanges([-3.] * D, [3.] * D, tight=True)
        s.SetEvaluationLimits(generations=400)
        s.Solve(models.rosen, termination.ChangeOverGeneration(1e-2, 7))
        runs.append((s.generations, s.evaluations, list(s.energy_history), list(s.bestSolution), str(s.Terminated(info=True))))
    same = runs[0] == runs[1] and 7 < runs[0][0] < 400
    print('rank %d batched Solve == stepwise Solve: %s (%d generations, %s)' % (rank, same, runs[0][0], runs[0][4]), flush=True)
    ok = ok and same
    t = torch.tensor([1 if ok else 0], device='cpu' if one_gpu else 'cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(t.item()) != 1:
        print('MULTI-GPU PARITY FAILED', flush=True)
        sys.exit(1)
    if rank == 0:
        print('multi-gpu parity ok (world %d)' % world, flush=True)


if __name__ == '__main__':
    main()
